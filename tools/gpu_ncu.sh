#!/bin/bash
# ncu --set full captures of named kernels, exported to CSV on the box (the .ncu-rep files exceed gpurun's 64 MiB return limit)
# usage: KERNELS="k_a_miller9:a:227328 k_f_miller_s:f:151552" bash tools/gpu_ncu.sh
mkdir -p gpurun_out
for spec in $KERNELS; do
  k=${spec%%:*}; rest=${spec#*:}; w=${rest%%:*}; n=${rest#*:}
  rep=/tmp/prof_$k
  PBC_B200_UNDER_PROFILER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $rep \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $w --configs none --n $n > gpurun_out/ncu_$k.out 2>&1; echo "ncu $k rc=$?"
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${k}_raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page details > gpurun_out/r2_ncu_${k}_details.txt 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_${k}_source.csv.gz
  ls -la $rep.ncu-rep gpurun_out/r2_ncu_${k}_*; rm -f $rep.ncu-rep
done
