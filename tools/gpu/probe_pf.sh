# A/B of the probe kernel's next-tile L2 prefetch distance (VDL_PROBE_PREFETCH: off, or N tiles ahead; 0 = own tile)
for dist in ${1:-0 64 256 1024}; do
  for q in q12 q05 q03 q19; do
    echo -n "dist $dist $q: "
    VDL_PROBE_PREFETCH=$dist python bench.py --query $q --sf 10 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
  done
done
