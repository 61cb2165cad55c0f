mkdir -p gpurun_out
AG_DEBUG_INGEST=1 timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03e_bench.json 2> gpurun_out/r03e_bench.err; echo "bench rc=$?"; cat gpurun_out/r03e_bench.json; tail -5 gpurun_out/r03e_bench.err
