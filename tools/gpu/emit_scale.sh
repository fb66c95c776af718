# Emit plans (Q3, Q19) at N ranks of one box: sharded tail vs one GPU.  Usage: emit_scale.sh PREFIX N
pre=${1:-r2t}; n=${2:-2}
for q in q03 q19; do
  python bench.py --query $q --sf 10 --no-e2e --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${pre}_${q}_n1.json 2> gpurun_out/${pre}_${q}_n1.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --query $q --sf 10 --no-e2e --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${pre}_${q}_n$n.json 2> gpurun_out/${pre}_${q}_n$n.err
  VDL_NO_TAIL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $n --query $q --sf 10 --no-e2e --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${pre}_${q}_n${n}_gather.json 2> gpurun_out/${pre}_${q}_n${n}_gather.err
  for f in gpurun_out/${pre}_${q}_n1.json gpurun_out/${pre}_${q}_n$n.json gpurun_out/${pre}_${q}_n${n}_gather.json; do
    tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$f', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'parity', (d.get('parity') or {}).get('status'), d['combine'][:60])
except Exception as e: print('$f', 'FAILED', e)
"
  done
done
tail -5 gpurun_out/${pre}_q03_n$n.err
