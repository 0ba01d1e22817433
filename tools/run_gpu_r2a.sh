#!/bin/bash
# round 2, pass a: the direct INTER_AREA kernel (Breakout) -- GPU parity tests, bench at fresh / steady / mid-game states, ncu
TAG=${1:-r2a}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_area_kernels.py -x -q > gpurun_out/${TAG}_pytest_area.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_area.log
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("%s: %.2f M/s render %.3f ms %.0f GB/s frac %.3f step %.3f ms episodes %s"%(sys.argv[2], d["value"]/1e6, r["launch_ms"], r["achieved"], r["frac"], r["step_kernel_ms"], d["episode_stats"]["episodes"]))
PY
}
B="--no-cpu-baseline --no-e2e"
timeout 300 python bench.py --steps 50 --warmup 5 $B > gpurun_out/${TAG}_bench_fresh.log 2>&1; show gpurun_out/${TAG}_bench_fresh.log "fresh 50 steps"
timeout 300 python bench.py --steps 200 --warmup 20 --presteps 2000 $B > gpurun_out/${TAG}_bench_steady.log 2>&1; show gpurun_out/${TAG}_bench_steady.log "steady presteps 2000"
timeout 300 python bench.py --policy track --presteps 3000 --steps 50 --warmup 5 $B > gpurun_out/${TAG}_bench_track.log 2>&1; show gpurun_out/${TAG}_bench_track.log "track presteps 3000"
TBX_AREA_KERNEL=tile timeout 300 python bench.py --steps 200 --warmup 20 --presteps 2000 $B > gpurun_out/${TAG}_bench_steady_tile.log 2>&1; show gpurun_out/${TAG}_bench_steady_tile.log "steady presteps 2000 (tile kernel)"
timeout 300 python bench.py --envs 131072 --steps 100 --warmup 10 --presteps 2000 $B > gpurun_out/${TAG}_bench_131072.log 2>&1; show gpurun_out/${TAG}_bench_131072.log "steady 131072 envs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --presteps 500 $B > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:brk_direct -s 6 -c 1 -o gpurun_out/${TAG}_prof_direct python bench.py --steps 4 --warmup 3 --presteps 2000 $B > gpurun_out/${TAG}_ncu_direct.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_prof_direct.ncu-rep > gpurun_out/${TAG}_ncu_direct.txt 2>&1; head -30 gpurun_out/${TAG}_ncu_direct.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:brk_direct -s 6 -c 1 -o gpurun_out/${TAG}_prof_direct_track python bench.py --steps 4 --warmup 3 --policy track --presteps 3000 $B > gpurun_out/${TAG}_ncu_direct_track.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_prof_direct_track.ncu-rep > gpurun_out/${TAG}_ncu_direct_track.txt 2>&1; head -12 gpurun_out/${TAG}_ncu_direct_track.txt
