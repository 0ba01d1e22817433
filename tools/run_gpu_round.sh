#!/bin/bash
# One GPU-box pass of the round: GPU test tier, the headline bench line + reference arm, the ncu launch list and the
# full-set captures whose summaries are committed under profiles/.  Run through gpurun from the repo root:
#   gpurun --timeout 2400 -- 'bash tools/run_gpu_round.sh r1'
TAG=${1:-r1}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_1gpu.log 2>&1; tail -1 gpurun_out/${TAG}_bench_1gpu.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.log 2>&1; tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-200
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
cap() { # name kernel-regex traffic-key bench-args...
  local name=$1 rx=$2 key=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$rx -s 6 -c 1 -o gpurun_out/${TAG}_prof_$name python bench.py "$@" $B > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_$name.ncu-rep > gpurun_out/${TAG}_ncu_$name.txt 2>&1
  python tools/ncu_traffic.py gpurun_out/${TAG}_prof_$name.ncu-rep $key gpurun_out/${TAG}_traffic.json > /dev/null 2>&1
}
cap area_brk area_tile breakout/gray84/65536 --game breakout
cap step_brk step_kernel breakout/step/65536 --game breakout
cap fill_brk base_fill breakout/rgba-fill/65536 --game breakout --obs rgba
cap patch_brk native_patch breakout/rgba-patch/65536 --game breakout --obs rgba
cap fill_amidar base_fill amidar/rgb-fill/65536 --game amidar --obs rgb
cap patch_amidar native_patch amidar/rgb-patch/65536 --game amidar --obs rgb
cap fill_si base_fill space_invaders/rgb-fill/65536 --game space_invaders --obs rgb
cap patch_si native_patch space_invaders/rgb-patch/65536 --game space_invaders --obs rgb
cap area_amidar area_tile amidar/gray84/65536 --game amidar
cap area_si area_tile space_invaders/gray84/65536 --game space_invaders
rm -f gpurun_out/${TAG}_prof_*.ncu-rep
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"]))
PY
}
for cfg in "breakout rgb" "breakout rgba" "breakout gray" "amidar rgb" "amidar gray" "space_invaders rgb" "space_invaders rgba" "space_invaders gray" "amidar gray84" "space_invaders gray84"; do
  set -- $cfg
  timeout 300 python bench.py --game $1 --obs $2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_$1_$2.log 2>&1; show gpurun_out/${TAG}_bench_$1_$2.log "$1 $2"
done
timeout 300 python bench.py --envs 131072 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_brk_131072.log 2>&1; show gpurun_out/${TAG}_bench_brk_131072.log "breakout gray84 131072 envs"
timeout 300 python bench.py --policy track --presteps 3000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_track.log 2>&1; show gpurun_out/${TAG}_bench_track.log "breakout gray84 mid-game"
timeout 300 python bench.py --obs rgb --policy track --presteps 3000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_track_rgb.log 2>&1; show gpurun_out/${TAG}_bench_track_rgb.log "breakout rgb mid-game"
for g in breakout amidar space_invaders; do
  timeout 300 python bench.py --wrapped --game $g --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_wrapped_$g.log 2>&1; tail -1 gpurun_out/${TAG}_bench_wrapped_$g.log | cut -c1-200
done
