mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --config c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r03_bench_c3_n4.json 2> gpurun_out/r03_bench_c3_n4.err; echo "bench c3@4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r03_bench_c3_n4.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['device_ms_per_step'], d['counts'])"
grep -v "^$" gpurun_out/r03_bench_c3_n4.err | tail -4
