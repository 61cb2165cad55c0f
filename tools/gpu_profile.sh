#!/bin/bash
# one GPU-box visit: bench line (not under a profiler), launch list of a short bench run, one `ncu --set full` capture of the dominant kernel.
# Usage (from the repo root): gpurun -- bash tools/gpu_profile.sh <tag> [kernel-regex]
tag=${1:-prof}
pat=${2:-k_build_tma}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 3 -c 1 -f -o gpurun_out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_full_bench.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/${tag}_full.ncu-rep --page details --csv > gpurun_out/${tag}_ncu_full_${pat}_details.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_${pat}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_full_${pat}_source.csv 2>/dev/null
cat gpurun_out/${tag}_bench_n1.json
