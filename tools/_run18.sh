mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r03j_bench.json 2> gpurun_out/r03j_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03j_bench.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:900], d['device_ms_per_step'], d['roofline'])"
AG_SCAN_TWOPASS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r03j_bench_2p.json 2> gpurun_out/r03j_bench_2p.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03j_bench_2p.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['device_ms_per_step'])"
for kb in 256 512 2048; do AG_STAGE_PIECE_KB=$kb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03j_bench_p$kb.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r03j_bench_p$kb.json')); print($kb, d['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"; done
AG_JOB_TIMING=1 AG_POST_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 >/dev/null | grep -E "^\s+\[" | tail -60
