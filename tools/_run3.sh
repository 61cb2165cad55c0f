mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
AG_POST_TIMING=1 timeout 300 python tools/file_level_time.py --reps 4 > gpurun_out/r03c_file_level.json 2> gpurun_out/r03c_file_level.err; cat gpurun_out/r03c_file_level.json; grep -E "^\[parse\]|contig" gpurun_out/r03c_file_level.err | tail -6
