for cfg in "4 2 32" "2 4 32" "1 8 32" "2 4 16" "4 2 16" "8 2 32" "2 8 32"; do
  set -- $cfg
  echo "chunk_mb=$1 slots=$2 threads=$3"
  AG_STAGE_CHUNK_MB=$1 AG_STAGE_SLOTS=$2 AG_STAGE_THREADS=$3 AG_POST_TIMING=1 timeout 200 python tools/file_level_time.py --reps 4 2>&1 | grep -E "ingest (reads|sam)\] staged" | tail -2
done
