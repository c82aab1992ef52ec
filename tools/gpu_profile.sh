#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one forward + `ncu --set full` of the top kernels.
# usage: tools/gpu_profile.sh <tag> [kernel-regex[:launches] ...]
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/prof_step.py --batch 8 > gpurun_out/${tag}_launches.log 2>&1
for spec in "$@"; do
  k=${spec%%:*}; cnt=${spec##*:}; [ "$cnt" = "$spec" ] && cnt=1
  name=$(echo $k | tr -c 'A-Za-z0-9_\n' '_')
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c $cnt -f \
      -o gpurun_out/${tag}_${name} python tools/prof_step.py --batch 8 > gpurun_out/${tag}_${name}.log 2>&1
  echo "$k rc=$?"
  # gpurun brings back at most 64 MiB: keep the CSV pages, drop big report files
  rep=gpurun_out/${tag}_${name}.ncu-rep
  if [ -f $rep ]; then
    ncu -i $rep --page raw --csv > gpurun_out/${tag}_${name}_raw.csv 2>/dev/null
    ncu -i $rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_${name}_source.csv.gz
    [ $(stat -c %s $rep) -gt 6000000 ] && rm -f $rep
  fi
done
du -sh gpurun_out
