#!/bin/bash
# tuning: variants of the library (TBX_LIB_PATH) on the headline workload
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms frac %.3f step %.3f ms"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["frac"], r["step_kernel_ms"]))
PY
}
for v in "" v3 v5 v6; do
  for g in breakout amidar; do
    if [ -z "$v" ]; then unset TBX_LIB_PATH; else export TBX_LIB_PATH=$PWD/toybox_b200/libtbx_$v.so; fi
    timeout 300 python bench.py --game $g --steps 200 --warmup 20 $B > gpurun_out/var_${g}_$v.log 2>&1; show gpurun_out/var_${g}_$v.log "$g variant '$v'"
  done
done
unset TBX_LIB_PATH
for grid in 444 592 740 888 1184 2368; do
  TBX_DIRECT_GRID=$grid timeout 300 python bench.py --steps 200 --warmup 20 $B > gpurun_out/var_grid_$grid.log 2>&1; show gpurun_out/var_grid_$grid.log "breakout grid=$grid"
done
