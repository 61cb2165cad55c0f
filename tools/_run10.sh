mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_misassembly.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -8
for lib in default aligngraph_b200/_variants/lib_tma_minb5.so aligngraph_b200/_variants/lib_tma_minb4.so; do
  if [ "$lib" = default ]; then unset AG_LIB_PATH; else export AG_LIB_PATH=$PWD/$lib; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$lib', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['rank0_breakdown_ms_per_step'], {k: d['device_ms_per_step'][k] for k in ('nodes','stage','build_kernel')})"
done
