mkdir -p gpurun_out
for ring in 8 16 32 64 256; do
  AG_STAGE_RING_MB=$ring timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03l_ring$ring.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r03l_ring$ring.json')); print('ring', $ring, d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"
done
for ring in 16 256; do
  AG_STAGE_RING_MB=$ring AG_STAGE_PIECE_KB=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03l_ring${ring}_p128.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r03l_ring${ring}_p128.json')); print('ring p128', $ring, d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"
done
for thr in 4 8; do
  AG_STAGE_RING_MB=16 AG_STAGE_THREADS=$thr timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r03l_ring16_t$thr.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r03l_ring16_t$thr.json')); print('ring16 threads', $thr, d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"
done
nproc; lscpu | grep -E "L2|L3|Model name"
