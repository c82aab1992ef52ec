mkdir -p gpurun_out
for L in 0 3; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01d_L${L}_launches.csv python tools/prof_step.py --batch 8 --first $L --last $((L+1)) > gpurun_out/r01d_L${L}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tcn2_mac -c 1 -f -o gpurun_out/r01d_L${L}_tcn python tools/prof_step.py --batch 8 --first $L --last $((L+1)) >> gpurun_out/r01d_L${L}.log 2>&1
  ncu -i gpurun_out/r01d_L${L}_tcn.ncu-rep --page raw --csv > gpurun_out/r01d_L${L}_tcn_raw.csv
  ncu -i gpurun_out/r01d_L${L}_tcn.ncu-rep --page source --csv | gzip > gpurun_out/r01d_L${L}_tcn_source.csv.gz
  rm -f gpurun_out/r01d_L${L}_tcn.ncu-rep
done
ls -la gpurun_out | tail -8
