#!/bin/bash
# one GPU-box visit: compute-sanitizer memcheck / racecheck / synccheck over smoke() (text ingestion, TMA-staged sweep, single-pass scans, fused
# extension, file output on a small case).  Usage (from the repo root): gpurun -- bash tools/gpu_sanitize.sh <tag>
tag=${1:-san}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_${tool}_smoke.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|smoke ok|RACECHECK SUMMARY|hazard" gpurun_out/${tag}_${tool}_smoke.log | head -5
done
