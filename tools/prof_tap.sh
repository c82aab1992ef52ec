timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tapmul_kernel -c 2 -f -o gpurun_out/r02E_tapmul python tools/prof_step.py --batch 8 --first 5 --last 7 > gpurun_out/r02E_tapmul.log 2>&1
ncu -i gpurun_out/r02E_tapmul.ncu-rep --page raw --csv > gpurun_out/r02E_tapmul_raw.csv 2>/dev/null
ncu -i gpurun_out/r02E_tapmul.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02E_tapmul_source.csv.gz
rm -f gpurun_out/r02E_tapmul.ncu-rep
