#!/bin/bash
TAG=${1:-r2i}
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest_gpu.log
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms episodes %s"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"], d["episode_stats"]["episodes"]))
PY
}
for g in amidar space_invaders breakout; do
  timeout 300 python bench.py --game $g --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_$g.log 2>&1; show gpurun_out/${TAG}_bench_$g.log "$g gray84 steady"
  timeout 300 python bench.py --game $g --steps 100 --warmup 10 --presteps 0 $B > gpurun_out/${TAG}_bench_${g}_fresh.log 2>&1; show gpurun_out/${TAG}_bench_${g}_fresh.log "$g gray84 fresh"
done
timeout 300 python bench.py --game amidar --steps 100 --warmup 10 --presteps 6000 $B > gpurun_out/${TAG}_bench_amidar_6000.log 2>&1; show gpurun_out/${TAG}_bench_amidar_6000.log "amidar gray84 presteps 6000"
TBX_AREA_KERNEL=tile timeout 300 python bench.py --game amidar --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_amidar_tile.log 2>&1; show gpurun_out/${TAG}_bench_amidar_tile.log "amidar gray84 steady (tile kernel)"
timeout 900 python bench.py --mixed 1048576 --steps 512 --warmup 10 > gpurun_out/${TAG}_bench_cfg5_mixed_1gpu.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg5_mixed_1gpu.log | cut -c1-200
cap() { # name kernel-regex skip bench-args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o /tmp/${TAG}_prof_$name python bench.py "$@" --steps 4 --warmup 3 $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py /tmp/${TAG}_prof_$name.ncu-rep 45 > gpurun_out/${TAG}_lines_$name.txt 2>&1
  rm -f /tmp/${TAG}_prof_$name.ncu-rep
}
cap direct_ami ami_direct 6 --game amidar
cap direct_si si_direct 6 --game space_invaders
head -12 gpurun_out/${TAG}_ncu_direct_ami.txt
