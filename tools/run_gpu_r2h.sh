#!/bin/bash
TAG=${1:-r2h}
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
for st in 1 0; do
for g in breakout space_invaders amidar; do
  TBX_STEP_STAGED=$st timeout 300 python bench.py --game $g --steps 100 --warmup 10 $B > gpurun_out/${TAG}_bench_${g}_staged$st.log 2>&1; show gpurun_out/${TAG}_bench_${g}_staged$st.log "$g gray84 steady STEP_STAGED=$st"
done; done
for pb in 1 0; do
for cfg in "space_invaders rgb" "amidar rgb" "breakout rgba" "space_invaders gray"; do
  set -- $cfg
  TBX_NATIVE_PATCH_BLOCKS=$pb timeout 300 python bench.py --game $1 --obs $2 --steps 50 --warmup 5 $B > gpurun_out/${TAG}_bench_$1_$2_pb$pb.log 2>&1; show gpurun_out/${TAG}_bench_$1_$2_pb$pb.log "$1 $2 steady PATCH_BLOCKS=$pb"
done; done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_full.log 2>&1; tail -1 gpurun_out/${TAG}_bench_full.log | cut -c1-3000
