mkdir -p gpurun_out
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r03_bench_n8.json 2> gpurun_out/r03_bench_n8.err; echo "bench c2@8 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03_bench_n8.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:700], d['device_ms_per_step'])"
grep -v "^$" gpurun_out/r03_bench_n8.err | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --config c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r03_bench_c4_n8.json 2> gpurun_out/r03_bench_c4_n8.err; echo "bench c4@8 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03_bench_c4_n8.json')); print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:700], d['device_ms_per_step'])"
grep -v "^$" gpurun_out/r03_bench_c4_n8.err | tail -3
