python -m pytest tests/test_gpu_queries.py -x -q -m gpu -k "q1 or static or smoke" 2>&1 | tail -3
for i in 1 2; do
for v in "VDL_RS_GEOMETRY=352,4" "VDL_RS_GEOMETRY=352,2" "VDL_RS_GEOMETRY=480,2" "VDL_RS_ALL_SLOTS=1"; do
  echo "$v: $(env $v python bench.py --query q01 --sf ${SF:-10} --no-e2e --no-cpu-baseline --steps 20 --warmup 5 2>&1 | grep -o "\"kernel_ms\": [0-9.]*\|\"kernel_ms_min\": [0-9.]*" | tr '\n' ' ')"
done; done
