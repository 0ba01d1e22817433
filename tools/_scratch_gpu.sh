set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu30.log
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"]))
PY
}
for k in patch canvas; do
export TBX_NATIVE_KERNEL=$k
for cfg in "breakout rgb" "breakout rgba" "breakout gray" "amidar rgb" "space_invaders rgb" "amidar gray" "space_invaders rgba"; do
  set -- $cfg
  timeout 300 python bench.py --game $1 --obs $2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench30_$1_$2_$k.log 2>&1; show gpurun_out/bench30_$1_$2_$k.log "$1 $2 kernel=$k"
done
done
unset TBX_NATIVE_KERNEL
timeout 300 python bench.py --obs rgb --policy track --presteps 3000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench30_track_rgb.log 2>&1; show gpurun_out/bench30_track_rgb.log "breakout rgb mid-game patch"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches30.csv python bench.py --game amidar --obs rgb --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -i "base_fill\|native_patch" gpurun_out/launches30.csv | tail -4
