set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu18.log
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"]))
PY
}
for v in "8 3" "8 2" "8 6" "4 3" "4 6"; do
  set -- $v
  export TBX_AREA_TILE_H=$1 TBX_AREA_MAX_RUN=$2
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench18_brk_h$1_r$2.log 2>&1; show gpurun_out/bench18_brk_h$1_r$2.log "breakout gray84 h$1 r$2"
  timeout 300 python bench.py --policy track --presteps 3000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench18_track_h$1_r$2.log 2>&1; show gpurun_out/bench18_track_h$1_r$2.log "breakout track h$1 r$2"
  for g in amidar space_invaders; do
    timeout 300 python bench.py --game $g --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench18_${g}_h$1_r$2.log 2>&1; show gpurun_out/bench18_${g}_h$1_r$2.log "$g gray84 h$1 r$2"
  done
done
unset TBX_AREA_TILE_H TBX_AREA_MAX_RUN
timeout 600 ncu --set full --clock-control none --import-source on -k regex:area_tile -s 6 -c 1 -o gpurun_out/prof_render_v18_brk_gray84 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_render18.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:area_tile -s 6 -c 1 -o gpurun_out/prof_render_v18_track python bench.py --policy track --presteps 3000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_render18t.log 2>&1
