#!/bin/bash
# `ncu --set full` capture of ONE launch of one kernel of the bench workload.  Usage: gpurun -- bash tools/gpu_ncu_one.sh <tag> <kernel-regex> [skip]
tag=$1; pat=$2; skip=${3:-3}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s $skip -c 1 -f -o gpurun_out/${tag} \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_bench.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/${tag}.ncu-rep --page details --csv > gpurun_out/${tag}_details.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
