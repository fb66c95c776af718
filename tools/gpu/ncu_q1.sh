# usage: tools/gpu/ncu_q1.sh <tag> [env assignments...]   -> gpurun_out/<tag>.ncu-rep
tag=$1; shift
env "$@" python bench.py --query q01 --sf 10 --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/$tag.plain.log 2>&1 || exit 1
env "$@" ncu --set full --clock-control none --import-source on -k regex:fused_scan_fold -s 2 -c 1 -f -o gpurun_out/$tag \
  python bench.py --query q01 --sf 10 --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/$tag.ncu.log 2>&1
tail -3 gpurun_out/$tag.ncu.log
