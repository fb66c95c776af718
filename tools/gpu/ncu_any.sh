# usage: tools/gpu/ncu_any.sh <tag> <kernel regex> <skip> -- bench args...   -> gpurun_out/<tag>.ncu-rep
tag=$1; kre=$2; skip=$3; shift 4
python bench.py "$@" --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/$tag.plain.log 2>&1 || { tail -5 gpurun_out/$tag.plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/$tag \
  python bench.py "$@" --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/$tag.ncu.log 2>&1
tail -2 gpurun_out/$tag.ncu.log
