#!/bin/bash
# run bench.py against each tuning variant library (aligngraph_b200/_variants/lib_*.so) and print the kernel times
for lib in default aligngraph_b200/_variants/lib_*.so; do
  if [ "$lib" = default ]; then unset AG_LIB_PATH; else export AG_LIB_PATH=$PWD/$lib; fi
  python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$lib', d['value'], {k: d['device_ms_per_step'][k] for k in ('nodes','edges')})"
done
