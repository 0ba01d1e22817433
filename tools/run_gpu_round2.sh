#!/bin/bash
# One GPU-box pass of round 2 (single GPU): GPU test tier, smoke, the headline bench line + reference arm, BASELINE configs 3 / 4 / 5
# at one GPU, the ncu launch list and the full-set captures whose summaries are committed under profiles/ (the .ncu-rep files are
# summarised on the box and left there).  Run through gpurun from the repo root:
#   gpurun --timeout 2700 -- 'bash tools/run_gpu_round2.sh r2'
TAG=${1:-r2}
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_1gpu.log 2>&1; tail -1 gpurun_out/${TAG}_bench_1gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1gpu_driver_flags.log 2>&1; tail -1 gpurun_out/${TAG}_bench_1gpu_driver_flags.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference_arm.log 2>&1; tail -1 gpurun_out/${TAG}_bench_reference_arm.log | cut -c1-200
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
timeout 900 python bench.py --game amidar --envs 262144 --obs rgb --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_cfg3_amidar_rgb_262144.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg3_amidar_rgb_262144.log | cut -c1-200
timeout 900 python bench.py --interventions --steps 256 --warmup 10 > gpurun_out/${TAG}_bench_cfg4_si_interventions_262144.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg4_si_interventions_262144.log | cut -c1-200
timeout 900 python bench.py --mixed 1048576 --steps 512 --warmup 10 > gpurun_out/${TAG}_bench_cfg5_mixed_1gpu.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg5_mixed_1gpu.log | cut -c1-200
for g in breakout amidar space_invaders; do
  timeout 300 python bench.py --wrapped --game $g --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_wrapped_$g.log 2>&1; tail -1 gpurun_out/${TAG}_bench_wrapped_$g.log | cut -c1-200
done
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms episodes %s"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"], d["episode_stats"]["episodes"]))
PY
}
for cfg in "breakout rgb" "breakout rgba" "breakout gray" "amidar rgb" "amidar gray" "amidar gray84" "space_invaders rgb" "space_invaders rgba" "space_invaders gray" "space_invaders gray84"; do
  set -- $cfg
  timeout 300 python bench.py --game $1 --obs $2 --steps 50 --warmup 5 $B > gpurun_out/${TAG}_bench_$1_$2.log 2>&1; show gpurun_out/${TAG}_bench_$1_$2.log "$1 $2 (steady state)"
done
timeout 300 python bench.py --envs 131072 --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_brk_131072.log 2>&1; show gpurun_out/${TAG}_bench_brk_131072.log "breakout gray84 131072 envs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file /tmp/${TAG}_launches_all.csv python bench.py --steps 20 --warmup 5 $B > gpurun_out/${TAG}_launches.log 2>&1
tail -120 /tmp/${TAG}_launches_all.csv > gpurun_out/${TAG}_launches.csv   # the last launches = warm-up + the timed steps (the 2,000 untimed presteps come first)
cap() { # name kernel-regex skip traffic-key bench-args...
  local name=$1 rx=$2 skip=$3 key=$4; shift 4
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o /tmp/${TAG}_prof_$name python bench.py "$@" --steps 4 --warmup 3 $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_lines.py /tmp/${TAG}_prof_$name.ncu-rep 40 > gpurun_out/${TAG}_lines_$name.txt 2>&1
  python tools/ncu_traffic.py /tmp/${TAG}_prof_$name.ncu-rep $key gpurun_out/${TAG}_traffic.json > /dev/null 2>&1
  rm -f /tmp/${TAG}_prof_$name.ncu-rep
}
cap direct_brk_steady brk_direct 6 breakout/gray84/65536 --game breakout
cap direct_brk_fresh brk_direct 6 breakout/gray84-fresh/65536 --game breakout --presteps 0
cap direct_brk_midgame brk_direct 6 breakout/gray84-track3000/65536 --game breakout --policy track --presteps 3000
cap direct_si_steady si_direct 6 space_invaders/gray84/65536 --game space_invaders
cap direct_si_fresh si_direct 6 space_invaders/gray84-fresh/65536 --game space_invaders --presteps 0
cap area_amidar_steady "area_tile|ami_direct" 6 amidar/gray84/65536 --game amidar
cap step_brk_steady "step_.*kernel" 2004 breakout/step/65536 --game breakout
cap step_amidar_steady "step_.*kernel" 2004 amidar/step/65536 --game amidar
cap step_si_steady "step_.*kernel" 2004 space_invaders/step/65536 --game space_invaders
cap fill_si base_fill 6 space_invaders/rgb-fill/65536 --game space_invaders --obs rgb
cap patch_si native_patch 6 space_invaders/rgb-patch/65536 --game space_invaders --obs rgb
cap fill_amidar base_fill 6 amidar/rgb-fill/65536 --game amidar --obs rgb
cap patch_amidar native_patch 6 amidar/rgb-patch/65536 --game amidar --obs rgb
cap fill_brk base_fill 6 breakout/rgba-fill/65536 --game breakout --obs rgba
cap patch_brk native_patch 6 breakout/rgba-patch/65536 --game breakout --obs rgba
