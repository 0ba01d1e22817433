#!/bin/bash
# quick single-GPU check
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms frac %.3f step %.3f ms"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["frac"], r["step_kernel_ms"]))
PY
}
for v in _head "" _head ""; do
  TBX_LIB_PATH=$PWD/toybox_b200/libtoybox_b200$v.so timeout 300 python bench.py --game space_invaders --steps 50 --warmup 5 $B > gpurun_out/q_bench_si$v.log 2>&1; show gpurun_out/q_bench_si$v.log "SI lib$v"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:_direct -s 6 -c 1 -o gpurun_out/q_prof_si3 python bench.py --game space_invaders --steps 4 --warmup 3 $B > gpurun_out/q_ncu_si3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:_direct -s 6 -c 1 -o gpurun_out/q_prof_brk python bench.py --game breakout --steps 4 --warmup 3 $B > gpurun_out/q_ncu_brk.log 2>&1
