set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu14.log
for cfg in "breakout gray84" "amidar gray84" "space_invaders gray84" "breakout rgb" "amidar rgb" "space_invaders rgb" "breakout gray" "breakout rgba"; do
  set -- $cfg
  timeout 300 python bench.py --game $1 --obs $2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench14_$1_$2.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/bench14_$1_$2.log"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("$1 $2: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms"%(d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"]))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 1 -o gpurun_out/prof_render_v14_brk_gray84 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_render14.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 1 -o gpurun_out/prof_render_v14_si_rgb python bench.py --game space_invaders --obs rgb --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_render14_si.log 2>&1
