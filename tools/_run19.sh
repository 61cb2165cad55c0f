mkdir -p gpurun_out
timeout 120 tools/scratch/hostreg_test /tmp/hostreg.bin
df /tmp | tail -1; mount | grep -E " / | /tmp " | head -3
for mode in fallocate plain; do
  if [ $mode = plain ]; then export AG_WRITE_PLAIN=1; else unset AG_WRITE_PLAIN; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r03k_bench_$mode.json 2> gpurun_out/r03k_bench_$mode.err; echo "bench $mode rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r03k_bench_$mode.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"
  AG_JOB_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 >/dev/null | grep -E "write unit files" | tail -3
done
unset AG_WRITE_PLAIN
for nc in 1 2; do
  timeout 900 python bench.py --config c3 --gpus 1 --steps 3 --warmup 3 --no-cpu --contexts-per-gpu $nc > gpurun_out/r03k_bench_c3_n1_ctx$nc.json 2> gpurun_out/r03k_bench_c3_n1_ctx$nc.err; echo "bench c3@1 ctx=$nc rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r03k_bench_c3_n1_ctx$nc.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'])"
  tail -2 gpurun_out/r03k_bench_c3_n1_ctx$nc.err
done
