mkdir -p gpurun_out
timeout 600 bash tools/variants.sh > gpurun_out/r03a_variants.log 2>&1
cat gpurun_out/r03a_variants.log
AG_LIB_PATH=$PWD/aligngraph_b200/_variants/lib_code4.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or edge" 2>&1 | tail -3
bash tools/gpu_ncu_one.sh r03a_kbuild k_build 3 2>&1 | tail -3
