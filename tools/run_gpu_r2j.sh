#!/bin/bash
TAG=${1:-r2j}
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms episodes %s"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"], d["episode_stats"]["episodes"]))
PY
}
for g in breakout amidar space_invaders; do
  timeout 300 python bench.py --game $g --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_$g.log 2>&1; show gpurun_out/${TAG}_bench_$g.log "$g gray84 steady"
done
TBX_STEP_STAGED=1 timeout 300 python bench.py --game breakout --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_breakout_staged.log 2>&1; show gpurun_out/${TAG}_bench_breakout_staged.log "breakout gray84 steady STEP_STAGED=1"
timeout 300 python bench.py --policy track --presteps 3000 --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_track.log 2>&1; show gpurun_out/${TAG}_bench_track.log "breakout gray84 track 3000"
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_full.log 2>&1; tail -1 gpurun_out/${TAG}_bench_full.log | cut -c1-200
cap() { # name kernel-regex skip bench-args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o /tmp/${TAG}_prof_$name python bench.py "$@" --steps 4 --warmup 3 $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py /tmp/${TAG}_prof_$name.ncu-rep 45 > gpurun_out/${TAG}_lines_$name.txt 2>&1
  rm -f /tmp/${TAG}_prof_$name.ncu-rep
}
cap step_brk "step_.*kernel" 2004 --game breakout
head -12 gpurun_out/${TAG}_ncu_step_brk.txt
