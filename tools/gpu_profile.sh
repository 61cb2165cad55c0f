#!/bin/bash
# one GPU-box visit: launch list of a short bench run + one `ncu --set full` capture of the dominant kernel(s).
# Usage (from the repo root): gpurun -- bash tools/gpu_profile.sh <tag> [kernel-regex]
tag=${1:-prof}
pat=${2:-k_build}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 4 -c 2 -f -o gpurun_out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_full_bench.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/${tag}_full.ncu-rep --page details --csv > gpurun_out/${tag}_full_details.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page source --csv > gpurun_out/${tag}_full_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
