#!/bin/bash
# quick GPU pass: area-kernel parity tests + device-timed render of Breakout gray84 at three states (no bench.py dependencies)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_area_kernels.py -x -q > gpurun_out/${TAG}_pytest_area.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_area.log
timeout 600 python - <<'PY'
import torch, numpy as np, toybox_b200, os
dev = torch.device("cuda", 0)
def run(name, policy, pre, n=65536, game="breakout"):
    pool = toybox_b200.BatchedToybox(game, n, device=dev, obs="gray84", seeds=(1234 + np.arange(n)) & 0xFFFFFFFF)
    acts = torch.empty(n, dtype=torch.int32, device=dev)
    obs = torch.empty((n,) + pool.obs_shape, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev)
    t = 0
    def fill():
        nonlocal t
        if policy == "track": pool.fill_policy_actions(acts, 1, t)
        else: pool.fill_random_actions(acts, 0xB200, t, 0)
        t += 1
    for _ in range(pre):
        fill(); pool.apply_ale_action(acts, auto_reset=True)
    for _ in range(5):
        fill(); pool.apply_ale_action(acts, auto_reset=True); pool.render(out=obs)
    K = 100
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(s)
    for k in range(K):
        fill(); ev[k][0].record(s); pool.apply_ale_action(acts, auto_reset=True); ev[k][1].record(s); pool.render(out=obs); ev[k][2].record(s)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    st = sum(e[0].elapsed_time(e[1]) for e in ev) / K; rd = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    print("%s: %.1f M steps/s  step %.4f ms  render %.4f ms  (%.0f GB/s, frac %.3f)" % (name, n / ms / 1e3, st, rd, n * 7344 / rd / 1e6, n * 7344 / rd / 1e6 / 6548.2), flush=True)
    pool.close()
run("fresh", "random", 0)
run("steady2000", "random", 2000)
run("track3000", "track", 3000)
run("steady2000 131072", "random", 2000, 131072)
for g in os.environ.get("TBX_DIRECT_GRID_SWEEP", "").split():
    os.environ["TBX_DIRECT_GRID"] = g
    run("steady2000 grid=" + g, "random", 2000)
PY
