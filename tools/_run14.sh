mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size_unit" 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03i_bench.json 2> gpurun_out/r03i_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r03i_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'], d['device_ms_per_step'], d['host_s_per_step'])"
tail -3 gpurun_out/r03i_bench.err
