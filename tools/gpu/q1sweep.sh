python -m pytest tests/test_gpu_queries.py -x -q -m gpu -k "q1 or static or smoke" 2>&1 | tail -3
for g in ${GEOMS:-"352,4" "352,3" "352,2"}; do
  VDL_DEBUG_SHAPE=1 VDL_RS_GEOMETRY=$g python bench.py --query q01 --sf ${SF:-10} --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep -o "fused scan: shape.*\|\"kernel_ms\": [0-9.]*\|\"frac\": [0-9.]*" | head -4
done
