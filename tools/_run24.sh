mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -x -q -k "not full_size" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r03o_bench.json 2> gpurun_out/r03o_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03o_bench.json')); print(d['value'], d['ms_per_step'], d['host_s_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'], d['device_ms_per_step'], d['gpu_launches'])"
