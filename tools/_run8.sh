mkdir -p gpurun_out
AG_DEBUG_INGEST=1 timeout 900 python -m pytest tests/test_gpu_ingest.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r03g_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['device_ms_per_step'], d['roofline'])"
bash tools/gpu_ncu_one.sh r03g_kbuild_tma k_build_tma 3 2>&1 | tail -2
