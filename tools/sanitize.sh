#!/bin/bash
# compute-sanitizer memcheck + racecheck over every kernel of the library on small pools (all games, all layouts, the
# wrapper stack, property access).  gpurun --timeout 2700 -- 'bash tools/sanitize.sh'
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, ".")
import toybox_b200
from toybox_b200.wrappers import DeepmindToybox
for game in toybox_b200.GAMES:
    pool = toybox_b200.BatchedToybox(game, 21, seeds=5)
    legal = np.asarray(pool.get_legal_action_set(), np.int32)
    rng = np.random.default_rng(0)
    acts = torch.empty(21, dtype=torch.int32, device=pool.device)
    for t in range(60):
        pool.apply_ale_action(legal[rng.integers(0, len(legal), 21)], auto_reset=True)
    for t in range(300):                              # scripted / synthetic streams: dead bricks, painted tiles, shot invaders
        if game == "breakout":
            pool.fill_policy_actions(acts, 1, t)
            pool.apply_ale_action(acts, auto_reset=True)
        else:
            pool.step_random(0xB200, t, 0, auto_reset=True)
    os.environ["TBX_STEP_STAGED"] = "1"
    pool.step_random(0xB200, 1000, 0, auto_reset=True)
    os.environ["TBX_STEP_STAGED"] = "0"
    pool.step_random(0xB200, 1001, 0, auto_reset=True)
    os.environ.pop("TBX_STEP_STAGED")
    js = pool.to_state_json([3, 4])                   # envs the direct kernels hand to the tile kernel / evaluate the slow way
    if game == "breakout":
        js[0]["bricks"][5]["position"]["x"] += 3.0
        js[1]["balls"] = [{"position": {"x": 100.0, "y": 5.0}, "velocity": {"x": 1.0, "y": 1.0}}]
    elif game == "space_invaders":
        js[0]["enemies"][0]["x"], js[0]["enemies"][0]["y"] = -5, -3
        js[1]["enemy_lasers"] = [{"x": 20, "y": 10, "w": 280, "h": 150, "t": 0, "movement": "Down", "speed": 3, "color": {"r": 9, "g": 99, "b": 199, "a": 255}}]
    else:
        js[0]["player"]["position"] = {"x": -40, "y": 2600}
        for b in js[1]["board"]["boxes"][:6]:
            b["painted"] = True
    pool.write_state_json(js, [3, 4])
    for kernel in ("", "tile"):
        if kernel:
            os.environ["TBX_AREA_KERNEL"] = kernel
        for mode in ("gray84", "rgb", "rgba", "gray", ("gray_area", 96, 80), ("gray_area", 48, 60), ("gray_area", 64, 64)):
            pool.render(obs=mode)
        os.environ.pop("TBX_AREA_KERNEL", None)
    pool.get_property("lives"); pool.set_property("lives", 2, [1] * 10 + [0] * 11)
    pool.to_state_json_text([0, 20])
    stats = torch.zeros(4, dtype=torch.int64, device=pool.device)
    pool.episode_stats_into(stats)
    if game == "breakout":
        from toybox_b200 import interventions as IV
        IV.breakout_add_channel(pool, 4); IV.breakout_channel_count(pool)
    pool.close()
    for mode in ("fill", "zero"):
        env = DeepmindToybox(game, 13, seeds=9, stack_reset=mode)
        env.reset()
        for t in range(12):
            env.step(torch.as_tensor(rng.integers(0, env.n_actions, 13).astype(np.int32), device=env.device))
        env.close()
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -12 gpurun_out/sanitizer_racecheck.log
