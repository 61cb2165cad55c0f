mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ingest.py -m gpu -x -q -k "not full_size_unit" 2>&1 | grep -v "^$" | tail -25
AG_POST_TIMING=1 timeout 300 python tools/file_level_time.py --reps 3 2>&1 | grep -E "ingest" | head -16
for t in 1 0; do
AG_TMA=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03f_bench_tma$t.json 2> gpurun_out/r03f_bench_tma$t.err; echo "bench tma=$t rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r03f_bench_tma$t.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['device_ms_per_step'], d['roofline'])"
done
