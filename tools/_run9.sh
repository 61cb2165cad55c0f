mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ingest.py -m gpu -x -q -k "not full_size_unit" 2>&1 | grep -v "^$" | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03h_bench.json 2> gpurun_out/r03h_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r03h_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['device_ms_per_step'], d['roofline'])"
echo "--- stager: normal / no copy / no read"
AG_POST_TIMING=1 timeout 200 python tools/file_level_time.py --reps 3 2>&1 | grep -E "ingest (reads|sam)\] staged" | tail -2
AG_STAGE_NOCOPY=1 AG_POST_TIMING=1 timeout 200 python tools/file_level_time.py --reps 3 2>&1 | grep -E "ingest (reads|sam)\] staged" | tail -2
AG_STAGE_NOREAD=1 AG_POST_TIMING=1 timeout 200 python tools/file_level_time.py --reps 3 2>&1 | grep -E "ingest (reads|sam)\] staged" | tail -2
bash tools/gpu_ncu_one.sh r03h_kbuild_tma k_build_tma 3 2>&1 | tail -2
