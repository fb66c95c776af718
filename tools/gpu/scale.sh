# usage: tools/gpu/scale.sh "<N list>" [bench args...]   -- prints value / ms_per_step / kernel_ms per N
ns=$1; shift
port=29600
for n in $ns; do
  port=$((port+1))
  if [ $n = 1 ]; then out=$(python bench.py --gpus 1 "$@" 2>&1 | tail -1)
  else out=$(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>&1 | tail -1); fi
  echo "N=$n ${VDL_NO_PEER:+(nccl)} $(echo "$out" | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|combine: [a-zA-Z-]*' | tr '\n' ' ')"
  echo "$out" >> gpurun_out/scale_raw.jsonl
done
