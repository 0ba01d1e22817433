#!/bin/bash
# compute-sanitizer memcheck + racecheck over every kernel of the library on small pools (all games, all layouts, the
# wrapper stack, property access).  gpurun --timeout 2700 -- 'bash tools/sanitize.sh'
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
import toybox_b200
from toybox_b200.wrappers import DeepmindToybox
for game in toybox_b200.GAMES:
    pool = toybox_b200.BatchedToybox(game, 21, seeds=5)
    legal = np.asarray(pool.get_legal_action_set(), np.int32)
    rng = np.random.default_rng(0)
    for t in range(60):
        pool.apply_ale_action(legal[rng.integers(0, len(legal), 21)], auto_reset=True)
    for mode in ("gray84", "rgb", "rgba", "gray", ("gray_area", 96, 80), ("gray_area", 48, 60)):
        pool.render(obs=mode)
    pool.get_property("lives"); pool.set_property("lives", 2, [1] * 10 + [0] * 11)
    pool.to_state_json([0, 20])
    pool.close()
    env = DeepmindToybox(game, 13, seeds=9)
    env.reset()
    for t in range(12):
        env.step(torch.as_tensor(rng.integers(0, env.n_actions, 13).astype(np.int32), device=env.device))
    env.close()
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -12 gpurun_out/sanitizer_racecheck.log
