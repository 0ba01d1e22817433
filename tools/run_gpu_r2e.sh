#!/bin/bash
# round 2: GPU test tier + ncu captures (summaries and per-line tables are produced on the box; the .ncu-rep files stay there)
TAG=${1:-r2e}
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest_gpu.log
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
cap() { # name kernel-regex traffic-key bench-args...
  local name=$1 rx=$2 key=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s 6 -c 1 -o /tmp/${TAG}_prof_$name python bench.py "$@" --steps 4 --warmup 3 $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py /tmp/${TAG}_prof_$name.ncu-rep 45 > gpurun_out/${TAG}_lines_$name.txt 2>&1
  python tools/ncu_traffic.py /tmp/${TAG}_prof_$name.ncu-rep $key gpurun_out/${TAG}_traffic.json > /dev/null 2>&1
  rm -f /tmp/${TAG}_prof_$name.ncu-rep
}
cap direct_steady brk_direct breakout/gray84/65536 --game breakout
cap direct_fresh brk_direct breakout/gray84-fresh/65536 --game breakout --presteps 0
cap direct_track brk_direct breakout/gray84-track3000/65536 --game breakout --policy track --presteps 3000
cap step_brk step_kernel breakout/step/65536 --game breakout
cap step_amidar step_kernel amidar/step/65536 --game amidar
cap step_si step_kernel space_invaders/step/65536 --game space_invaders
cap area_amidar area_tile amidar/gray84/65536 --game amidar
cap area_si area_tile space_invaders/gray84/65536 --game space_invaders
head -12 gpurun_out/${TAG}_ncu_step_brk.txt
