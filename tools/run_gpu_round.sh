set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu25.log
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"]))
PY
}
for h in 8 16; do
export TBX_AREA_TILE_H=$h
for g in breakout amidar space_invaders; do
  timeout 300 python bench.py --game $g --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench25_${g}_h$h.log 2>&1; show gpurun_out/bench25_${g}_h$h.log "$g gray84 h$h"
done
timeout 300 python bench.py --policy track --presteps 3000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench25_track_h$h.log 2>&1; show gpurun_out/bench25_track_h$h.log "breakout track h$h"
timeout 300 python bench.py --wrapped --steps 50 --warmup 5 > gpurun_out/bench25_wrapped_h$h.log 2>&1; tail -1 gpurun_out/bench25_wrapped_h$h.log | cut -c1-200
done
