#!/bin/bash
# round 2, full pass: GPU test tier, smoke, the headline bench line + reference arm, BASELINE configs 3/4/5, ncu captures
TAG=${1:-r2d}
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_1gpu.log 2>&1; tail -1 gpurun_out/${TAG}_bench_1gpu.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1gpu_driver.log 2>&1; tail -1 gpurun_out/${TAG}_bench_1gpu_driver.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.log 2>&1; tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-200
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
timeout 900 python bench.py --game amidar --envs 262144 --obs rgb --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_cfg3_amidar_rgb.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg3_amidar_rgb.log | cut -c1-300
timeout 900 python bench.py --interventions --steps 256 --warmup 10 > gpurun_out/${TAG}_bench_cfg4_interventions.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg4_interventions.log | cut -c1-1200
timeout 900 python bench.py --mixed 1048576 --steps 512 --warmup 10 > gpurun_out/${TAG}_bench_cfg5_mixed_1gpu.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg5_mixed_1gpu.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 5 $B > gpurun_out/${TAG}_launches.log 2>&1
cap() { # name kernel-regex traffic-key bench-args...
  local name=$1 rx=$2 key=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s 6 -c 1 -o gpurun_out/${TAG}_prof_$name python bench.py "$@" --steps 4 --warmup 3 $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_traffic.py gpurun_out/${TAG}_prof_$name.ncu-rep $key gpurun_out/${TAG}_traffic.json > /dev/null 2>&1
}
cap direct_steady brk_direct breakout/gray84/65536 --game breakout
cap direct_fresh brk_direct breakout/gray84-fresh/65536 --game breakout --presteps 0
cap direct_track brk_direct breakout/gray84-track3000/65536 --game breakout --policy track --presteps 3000
cap step_brk step_kernel breakout/step/65536 --game breakout
cap step_amidar step_kernel amidar/step/65536 --game amidar
cap step_si step_kernel space_invaders/step/65536 --game space_invaders
cap area_amidar area_tile amidar/gray84/65536 --game amidar
cap area_si area_tile space_invaders/gray84/65536 --game space_invaders
ls -la gpurun_out/${TAG}_prof_*.ncu-rep | head -20
