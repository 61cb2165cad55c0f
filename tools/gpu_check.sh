#!/bin/bash
# one GPU-box visit: parity tests, bench line, tuning variants.  Usage (from the repo root): gpurun -- bash tools/gpu_check.sh [tag]
tag=${1:-check}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
timeout 600 bash tools/variants.sh > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
