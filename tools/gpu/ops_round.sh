# Per-op kernels: GPU tests of the ops, the sweep, Q3, ncu --set full of the op kernels at 1e8 rows.
pre=${1:-r2q}
(time python -m pytest tests/test_gpu_ops.py tests/test_agg_strategies.py tests/test_gpu_queries.py tests/test_gpu_fuzz.py -m gpu -q) > gpurun_out/${pre}_tests.log 2>&1; tail -5 gpurun_out/${pre}_tests.log
python tools/sweep_ops.py --sizes ${2:-1e7,1e8} --reps 5 > gpurun_out/${pre}_op_sweep.jsonl 2> gpurun_out/${pre}_op_sweep.err; tail -3 gpurun_out/${pre}_op_sweep.err
grep -E '"rows": 100000000' gpurun_out/${pre}_op_sweep.jsonl | cut -c1-200
python bench.py --query q03 --sf 10 --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${pre}_q03.json 2> gpurun_out/${pre}_q03.err; cut -c1-200 gpurun_out/${pre}_q03.json
ncu --set full --clock-control none --import-source on -k regex:'^scatter_kernel|bucket_place|bucket_count|fold_lookback|select_lookback' -c 14 -f -o gpurun_out/${pre}_ops_full \
  python tools/sweep_ops.py --sizes 1e8 --reps 1 > gpurun_out/${pre}_ops_ncu.log 2>&1
tail -2 gpurun_out/${pre}_ops_ncu.log
