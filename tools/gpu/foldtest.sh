python tools/sweep_ops.py --sizes 1e8,1e9 --reps 5 2>/dev/null | grep -E "Fold|Partition \(32|Scatter" | cut -c1-200
python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -2
