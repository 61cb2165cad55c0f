mkdir -p gpurun_out
AG_DEBUG_INGEST=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "full_size_unit or regrow or fused" 2>&1 | grep -v "^$" | grep -E "ag ingest|passed|failed|Error|assert" | tail -30
