mkdir -p gpurun_out
timeout 2400 python bench.py --gpus 1 --config c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r03_bench_c4_n1.json 2> gpurun_out/r03_bench_c4_n1.err; echo "bench c4@1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03_bench_c4_n1.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:700], d['device_ms_per_step'], d['counts'])"
grep -v "^$" gpurun_out/r03_bench_c4_n1.err | tail -4
