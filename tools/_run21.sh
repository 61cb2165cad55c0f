mkdir -p gpurun_out
taskset -c 0-3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03m_4cores.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r03m_4cores.json')); print('4 cores', d['value'], d['ms_per_step'], d['host_s_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"
AG_POST_TIMING=1 AG_JOB_TIMING=1 taskset -c 0-3 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 >/dev/null | grep -E "^\s+\[" | tail -42
