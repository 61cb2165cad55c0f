timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_step" 2>&1 | grep -v "^$" | tail -40
