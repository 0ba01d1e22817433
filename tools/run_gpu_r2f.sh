#!/bin/bash
# round 2: GPU test tier, SI direct kernel bench, step-kernel captures at steady state
TAG=${1:-r2f}
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
for g in space_invaders amidar breakout; do
  timeout 300 python bench.py --game $g --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_$g.log 2>&1; show gpurun_out/${TAG}_bench_$g.log "$g gray84 steady"
  timeout 300 python bench.py --game $g --steps 100 --warmup 10 --presteps 0 $B > gpurun_out/${TAG}_bench_${g}_fresh.log 2>&1; show gpurun_out/${TAG}_bench_${g}_fresh.log "$g gray84 fresh"
done
TBX_AREA_KERNEL=tile timeout 300 python bench.py --game space_invaders --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_si_tile.log 2>&1; show gpurun_out/${TAG}_bench_si_tile.log "space_invaders gray84 steady (tile kernel)"
cap() { # name kernel-regex skip bench-args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o /tmp/${TAG}_prof_$name python bench.py "$@" --steps 4 --warmup 3 $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py /tmp/${TAG}_prof_$name.ncu-rep 45 > gpurun_out/${TAG}_lines_$name.txt 2>&1
  rm -f /tmp/${TAG}_prof_$name.ncu-rep
}
cap step_brk step_kernel 2004 --game breakout
cap step_amidar step_kernel 2004 --game amidar
cap step_si step_kernel 2004 --game space_invaders
cap direct_si si_direct 6 --game space_invaders
head -12 gpurun_out/${TAG}_ncu_step_brk.txt; head -12 gpurun_out/${TAG}_ncu_direct_si.txt
