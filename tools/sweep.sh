for o in "" "eval_mode=1" "chunk=4096" "chunk=16384" "chunk=32768" "points_per_cell=4" "points_per_cell=32" "eval_mode=1 chunk=16384" "fps_barrier=1"; do
  timeout 120 python tools/quick_bench.py 1000000 1000 $o 2>&1 | grep -E "opts|fps |evals/s"
done
