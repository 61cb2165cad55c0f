mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err; echo "bench rc=$?"; cat gpurun_out/r03d_bench.json; tail -5 gpurun_out/r03d_bench.err
