# Round-2 multi-GPU measurements on one box: N ranks (default 8).  JSON lines under gpurun_out/${pre}_*.json(l)
pre=${1:-r02s}; n=${2:-8}
run() {  # tag, bench args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" > gpurun_out/${pre}_${tag}_n$n.json 2> gpurun_out/${pre}_${tag}_n$n.err
  tail -1 gpurun_out/${pre}_${tag}_n$n.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d.get('e2e') or {}
    print('$tag N=$n', 'ms/step', round(d['ms_per_step'],4), 'rows/s %.3g' % d['value'], 'kernel_ms', round(d['roofline'].get('kernel_ms',0),4), 'parity', (d.get('parity') or {}).get('status'), 'e2e_ms', e.get('ms_per_step'), d['combine'][:50])
except Exception as ex: print('$tag N=$n FAILED', ex)
"
}
run q06_sf100 --steps 20 --warmup 5 --no-cpu-baseline
run q01_sf100 --query q01 --sf 100 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
run q05_sf100 --query q05 --sf 100 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity
run q12_sf100 --query q12 --sf 100 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity
run q03_sf100 --query q03 --sf 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity
run q03_sf10 --query q03 --sf 10 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
run q19_sf100 --query q19 --sf 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29950 tools/suite_bench.py --sf 100 --steps 5 --warmup 2 > gpurun_out/${pre}_suite_sf100_n$n.jsonl 2> gpurun_out/${pre}_suite_sf100_n$n.err
cut -c1-260 gpurun_out/${pre}_suite_sf100_n$n.jsonl
tail -3 gpurun_out/${pre}_suite_sf100_n$n.err
