# Round profile capture: for each workload, (1) plain run must exit 0, (2) launch list with gpu__time_duration,
# (3) one ncu --set full capture of the dominant kernel.  Outputs under gpurun_out/ with prefix $1.
pre=${1:-r01b}
only=${2:-.}          # regex over the workload tags
run() {  # tag kernel-regex skip bench-args...
  tag=$1; kre=$2; skip=$3; shift 3
  echo "$tag" | grep -Eq "$only" || return 0
  python bench.py "$@" --no-e2e --no-cpu-baseline --no-parity --steps 3 --warmup 3 > gpurun_out/${pre}_$tag.plain.json 2> gpurun_out/${pre}_$tag.plain.err || { echo "$tag: plain run failed"; tail -3 gpurun_out/${pre}_$tag.plain.err; return; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${pre}_launches_$tag.csv \
    python bench.py "$@" --no-e2e --no-cpu-baseline --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/${pre}_${tag}_full \
    python bench.py "$@" --no-e2e --no-cpu-baseline --no-parity --steps 3 --warmup 3 > gpurun_out/${pre}_$tag.ncu.log 2>&1
  echo "$tag: done ($(tail -1 gpurun_out/${pre}_$tag.ncu.log))"
}
# kernel names: the run-time compiled instantiations (vdl_scan_jit / vdl_probe_jit: NVRTC, round 2) or the precompiled ones
run q06_sf100 "vdl_scan_jit|fused_scan_fold" 4 --query q06 --sf 100
run q01_sf10 "vdl_scan_jit|fused_scan_fold" 4 --query q01 --sf 10
run q01_sf100 "vdl_scan_jit|fused_scan_fold" 4 --query q01 --sf 100
run q05_sf10 "vdl_probe_jit|probe_kernel" 4 --query q05 --sf 10
run q03_sf10 "vdl_probe_jit|probe_kernel" 4 --query q03 --sf 10
run q12_sf10 "vdl_probe_jit|probe_kernel" 4 --query q12 --sf 10
run q19_sf10 "vdl_probe_jit|probe_kernel" 4 --query q19 --sf 10
