"""GPU tier: the CUDA path (through the C ABI) against the CPU oracle on identical seeds and action streams.
Bit-exact: scalars every frame, frames in every layout and full state JSON at intervals."""
import numpy as np
import pytest

from conftest import json_diff

pytestmark = pytest.mark.gpu
GAMES = ["breakout", "amidar", "space_invaders"]
ORACLE_MODE = {"rgba": "rgba", "rgb": "rgb", "gray": "gray", "gray84": "gray84"}


def actions_for(oracle_mod, game, n, t, seed=0xB200, env0=0):
    legal = np.asarray(oracle_mod.LEGAL[game], np.int32)
    return legal[[oracle_mod.action_index(seed, env0 + i, t, len(legal)) for i in range(n)]]


def check_frames(pool, ref, n, modes=("rgba", "rgb", "gray", "gray84")):
    for mode in modes:
        got = pool.render(obs=mode).cpu().numpy().reshape(n, -1)
        want = ref.render(ORACLE_MODE[mode]).reshape(n, -1)
        bad = np.argwhere(got != want)
        assert bad.size == 0, (mode, bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("game", GAMES)
def test_initial_state_matches_fixture_lineage(tbx, oracle_mod, game):
    """A fresh pool is n copies of what a fresh ctoybox.Toybox(game) holds."""
    n = 5
    pool = tbx.BatchedToybox(game, n)
    ref = oracle_mod.OracleBatch(game, n)
    states = pool.to_state_json()
    for i in range(n):
        assert json_diff(states[i], ref.state_json(i)) == []
    # config_to_json shows the simulator as it is now: its rng is the one new_game has advanced (env 0's in a pool)
    want = oracle_mod.CODEC[game][2](ref.cfg)
    want["rand"] = {"state": [int(ref.sim[0].s[0]), int(ref.sim[0].s[1])]}
    assert json_diff(pool.config_to_json(), want) == []
    one = oracle_mod.OracleToybox(game)                      # ... which is what the batch-1 ctoybox look-alike reports
    assert json_diff(pool.config_to_json(), one.config_to_json()) == []
    check_frames(pool, ref, n)
    # write_config_json replaces the simulator, rng included: the next new_game starts the same lineage again
    cfg = pool.config_to_json()
    pool.write_config_json(cfg)
    pool.new_game()
    one.write_config_json(cfg)
    one.new_game()
    assert json_diff(pool.to_state_json([n - 1])[0], one.to_state_json()) == []
    assert json_diff(pool.config_to_json(), one.config_to_json()) == []
    pool.close()


@pytest.mark.parametrize("game", GAMES)
def test_rollout_bit_exact(tbx, oracle_mod, game):
    """Random-action rollout with auto-reset from per-env seeds: every scalar every frame, frames + JSON at intervals."""
    n, steps = 96, 1500
    pool = tbx.BatchedToybox(game, n, seeds=1234)
    ref = oracle_mod.OracleBatch(game, n, seeds=1234 + np.arange(n))
    for t in range(steps):
        acts = actions_for(oracle_mod, game, n, t)
        pool.apply_ale_action(acts, auto_reset=True)
        r, d, s, l = ref.step(acts, auto_reset=True)
        assert np.array_equal(pool.score.cpu().numpy(), s), t
        assert np.array_equal(pool.lives.cpu().numpy(), l), t
        assert np.array_equal(pool.reward.cpu().numpy(), r), t
        assert np.array_equal(pool.done.cpu().numpy().astype(bool), d), t
        if t % 100 == 0 or t == steps - 1:
            check_frames(pool, ref, n)
            for i in (0, n // 2, n - 1):
                assert json_diff(pool.to_state_json([i])[0], ref.state_json(i)) == [], (t, i)
    pool.close()


def test_breakout_tracking_policy_hits_bricks(tbx, oracle_mod):
    """A paddle that tracks the ball clears bricks: exercises paddle bounces, brick hits, speed-up, level refill."""
    n, steps = 32, 6000
    pool = tbx.BatchedToybox("breakout", n, seeds=99)
    ref = oracle_mod.OracleBatch("breakout", n, seeds=99 + np.arange(n))
    for t in range(steps):
        acts = np.zeros(n, np.int32)
        for i in range(n):
            s = ref.states[i]
            if s.is_dead:
                acts[i] = 1
            else:
                bx = s.balls[0].position.x if s.n_balls else 120.0
                off = ((i * 7 + t // 50) % 9) - 4          # deliberate aim error so all paddle segments get used
                acts[i] = 3 if bx + off > s.paddle.position.x + 1 else 4 if bx + off < s.paddle.position.x - 1 else 0
        pool.apply_ale_action(acts, auto_reset=True)
        r, d, s_, l = ref.step(acts, auto_reset=True)
        assert np.array_equal(pool.score.cpu().numpy(), s_), t
        assert np.array_equal(pool.lives.cpu().numpy(), l), t
        if t % 500 == 0 or t == steps - 1:
            check_frames(pool, ref, n, modes=("rgb", "gray84"))
            for i in (0, n - 1):
                assert json_diff(pool.to_state_json([i])[0], ref.state_json(i)) == [], (t, i)
    assert int(pool.score.max()) > 30
    pool.close()


@pytest.mark.parametrize("game", GAMES)
def test_json_round_trip_and_intervention(tbx, oracle_mod, game):
    """write_state_json(to_state_json()) is the identity, and an edited state steps exactly like the oracle's."""
    n = 8
    pool = tbx.BatchedToybox(game, n, seeds=7)
    ref = oracle_mod.OracleBatch(game, n, seeds=7 + np.arange(n))
    for t in range(150):
        acts = actions_for(oracle_mod, game, n, t, seed=5)
        pool.apply_ale_action(acts, auto_reset=True)
        ref.step(acts, auto_reset=True)
    before = pool.to_state_json()
    pool.write_state_json(before)
    after = pool.to_state_json()
    for i in range(n):
        assert json_diff(before[i], after[i]) == []
    # intervention on env 2: lives and a game-specific field
    js = before[2]
    js["lives"] = 1
    if game == "breakout":
        for b in js["bricks"][:12]:
            b["alive"] = False
        js["bricks"][20]["color"] = {"r": 1, "g": 2, "b": 3, "a": 255}
        js["balls"].append({"position": {"x": 100.0, "y": 100.0}, "velocity": {"x": 1.0, "y": -2.0}})
    elif game == "amidar":
        js["jumps"] = 1
        js["enemies"] = js["enemies"][:4]
        js["enemies"][0]["ai"] = {"EnemyRandomMvmt": {"start": {"tx": 0, "ty": 0}, "start_dir": "Right", "dir": "Right"}}
    else:
        js["ufo"]["appearance_counter"] = 3
        js["shields"][0]["data"][0][5]["a"] = 0
    pool.write_state_json([js], [2])
    ref.write_state_json(2, js)
    assert json_diff(pool.to_state_json([2])[0], ref.state_json(2)) == []
    for t in range(400):
        acts = actions_for(oracle_mod, game, n, t, seed=6)
        pool.apply_ale_action(acts, auto_reset=True)
        r, d, s, l = ref.step(acts, auto_reset=True)
        assert np.array_equal(pool.score.cpu().numpy(), s), t
        assert np.array_equal(pool.lives.cpu().numpy(), l), t
    check_frames(pool, ref, n)
    for i in range(n):
        assert json_diff(pool.to_state_json([i])[0], ref.state_json(i)) == [], i
    pool.close()


@pytest.mark.parametrize("game", GAMES)
def test_step_random_equals_fill_then_step(tbx, oracle_mod, game):
    """tbx_step_random (the synthetic stream generated inside the step kernel) == tbx_fill_actions + tbx_step, and both follow
    the oracle driven by the CPU restatement of the stream"""
    import torch
    n = 37
    a = tbx.BatchedToybox(game, n, seeds=11)
    b = tbx.BatchedToybox(game, n, seeds=11)
    ref = oracle_mod.OracleBatch(game, n, seeds=11 + np.arange(n))
    acts = torch.empty(n, dtype=torch.int32, device=a.device)
    legal = np.asarray(oracle_mod.LEGAL[game], np.int32)
    for t in range(300):
        a.step_random(0xB200, t, env0=5, auto_reset=True)
        b.fill_random_actions(acts, 0xB200, t, 5)
        b.apply_ale_action(acts, auto_reset=True)
        r, d, s, l = ref.step(legal[[oracle_mod.action_index(0xB200, 5 + i, t, len(legal)) for i in range(n)]], auto_reset=True)
        for pool in (a, b):
            assert np.array_equal(pool.score.cpu().numpy(), s) and np.array_equal(pool.lives.cpu().numpy(), l), t
            assert np.array_equal(pool.reward.cpu().numpy(), r) and np.array_equal(pool.done.cpu().numpy().astype(bool), d), t
    assert a.to_state_json([0, n - 1]) == b.to_state_json([0, n - 1])
    a.close(); b.close()


def test_ragged_batch_sizes_and_invalid_action(tbx, oracle_mod):
    for n in (1, 7, 33):
        pool = tbx.BatchedToybox("space_invaders", n, seeds=3)
        ref = oracle_mod.OracleBatch("space_invaders", n, seeds=3 + np.arange(n))
        for t in range(200):
            acts = actions_for(oracle_mod, "space_invaders", n, t)
            pool.apply_ale_action(acts)
            ref.step(acts, auto_reset=False)
        check_frames(pool, ref, n, modes=("gray", "gray84"))
        pool.close()
    pool = tbx.BatchedToybox("breakout", 4)
    pool.apply_ale_action([0, 1, 99, 3])
    with pytest.raises(ValueError):
        pool.check()
    pool.close()


def test_step_host_matches_device_path(tbx, oracle_mod):
    n = 64
    pool = tbx.BatchedToybox("breakout", n, seeds=11, obs="gray84")
    ref = oracle_mod.OracleBatch("breakout", n, seeds=11 + np.arange(n))
    for t in range(120):
        acts = actions_for(oracle_mod, "breakout", n, t)
        obs, r, d, info = pool.step_host(acts)
        ro, do, so, lo = ref.step(acts, auto_reset=True)
        assert np.array_equal(r.numpy(), ro) and np.array_equal(info["score"].numpy(), so) and np.array_equal(info["lives"].numpy(), lo)
    assert np.array_equal(obs.numpy().reshape(n, -1), ref.render("gray84").reshape(n, -1))
    pool.close()


def test_episode_stats(tbx, oracle_mod):
    n = 128
    pool = tbx.BatchedToybox("space_invaders", n, seeds=21)
    ref = oracle_mod.OracleBatch("space_invaders", n, seeds=21 + np.arange(n))
    episodes = 0
    for t in range(2500):
        acts = actions_for(oracle_mod, "space_invaders", n, t)
        pool.apply_ale_action(acts, auto_reset=True)
        r, d, s, l = ref.step(acts, auto_reset=True)
        episodes += int(d.sum())
    stats = pool.episode_stats()
    assert stats[0] == episodes
    pool.close()


def test_batched_intervention_context_manager(tbx, oracle_mod):
    from toybox_b200.interventions import BatchedIntervention, parse_property_access
    assert parse_property_access("abc.def[7][8].y[5]") == ["abc", "def", 7, 8, "y", 5]      # test_get_property.py:76-78
    n = 16
    pool = tbx.BatchedToybox("space_invaders", n, seeds=5)
    ref = oracle_mod.OracleBatch("space_invaders", n, seeds=5 + np.arange(n))
    ids = [1, 4, 9]
    with BatchedIntervention(pool, ids) as iv:
        iv.set("lives", 1)
        iv.set("ufo.appearance_counter", [3, 4, 5])
        assert iv.get("ship.x") == [68, 68, 68]
    assert iv.dirty_state and not iv.dirty_config
    for k, i in enumerate(ids):
        js = ref.state_json(i)
        js["lives"] = 1
        js["ufo"]["appearance_counter"] = 3 + k
        ref.write_state_json(i, js)
    for t in range(300):
        acts = actions_for(oracle_mod, "space_invaders", n, t)
        pool.apply_ale_action(acts, auto_reset=True)
        r, d, s, l = ref.step(acts, auto_reset=True)
        assert np.array_equal(pool.lives.cpu().numpy(), l), t
    for i in range(n):
        assert json_diff(pool.to_state_json([i])[0], ref.state_json(i)) == [], i
    with BatchedIntervention(pool) as iv:
        iv.config["start_lives"] = 7
    assert iv.dirty_config
    assert int(pool.get_lives().min()) == 7
    pool.close()


def test_ctoybox_shim_and_env_surface(tbx, oracle_mod):
    """The batch-1 drop-in (`toybox_b200.ctoybox.Toybox`) and the gym-free env classes behave like the reference's."""
    from toybox_b200.ctoybox import Toybox, Input
    from toybox_b200.envs import BreakoutEnv, BatchedToyboxEnv
    with Toybox("breakout") as tb:
        ref = oracle_mod.OracleToybox("breakout")
        fire = Input()
        fire.button1 = True
        tb.apply_action(fire)
        ref.apply_action(fire)
        for a in [3, 3, 4, 0, 1] * 30:
            tb.apply_ale_action(a)
            ref.apply_ale_action(a)
        assert json_diff(tb.to_state_json(), ref.to_state_json()) == []
        assert tb.get_state().shape == (160, 240, 1) and np.array_equal(tb.get_state(), ref.get_state())
        assert np.array_equal(tb.get_rgb_frame(), ref.get_rgb_frame())
        assert tb.get_legal_action_set() == [0, 1, 3, 4] and tb.get_lives() == ref.get_lives() and not tb.game_over()
        with pytest.raises(ValueError):
            tb.apply_ale_action(42)
        assert tb.query_state_json("bricks_remaining") == 108
    env = BreakoutEnv(grayscale=False)
    obs = env.reset()
    assert obs.shape == (160, 240, 3)
    obs, reward, done, info = env.step(1)
    assert reward == 0 and not done and info["lives"] == 5 and env.ale.getScreenRGB().shape == (160, 240, 3)
    env.close()
    benv = BatchedToyboxEnv("amidar", 32, seed=3)
    obs, reward, done, info = benv.step(np.zeros(32, np.int64))
    assert tuple(obs.shape) == (32, 84, 84, 1) and int(info["lives"].min()) == 3
    benv.close()


def test_delta_render_mixed_states_in_one_chunk(tbx, oracle_mod):
    """Consecutive envs of a chunk share one persistent canvas: dead bricks, custom brick tables (base 0 next to
    base 1), painted Amidar tiles and moved shields must never leak from one env's frame into the next."""
    rng = np.random.default_rng(17)
    n = 24
    pool = tbx.BatchedToybox("breakout", n, seeds=100)
    ref = oracle_mod.OracleBatch("breakout", n, seeds=100 + np.arange(n))
    states = pool.to_state_json()
    for i, js in enumerate(states):
        for k in rng.choice(108, size=int(rng.integers(0, 108)), replace=False):
            js["bricks"][int(k)]["alive"] = False
        js["score"], js["lives"] = int(rng.integers(0, 900)), int(rng.integers(1, 6))
        js["paddle"]["position"]["x"] = float(rng.integers(24, 216))
        if i % 5 == 3:
            js["bricks"][int(rng.integers(108))]["color"] = {"r": 250, "g": 250, "b": 250, "a": 255}      # custom table
        if i % 7 == 2:
            js["balls"].append({"position": {"x": float(rng.integers(20, 220)), "y": float(rng.integers(30, 150))},
                                "velocity": {"x": 1.0, "y": 1.0}})
        ref.write_state_json(i, js)
    pool.write_state_json(states)
    check_frames(pool, ref, n)
    for t in range(60):
        acts = actions_for(oracle_mod, "breakout", n, t)
        pool.apply_ale_action(acts, auto_reset=True)
        ref.step(acts, auto_reset=True)
    check_frames(pool, ref, n)
    pool.close()
    pool = tbx.BatchedToybox("amidar", n, seeds=100)
    ref = oracle_mod.OracleBatch("amidar", n, seeds=100 + np.arange(n))
    states = pool.to_state_json()
    for i, js in enumerate(states):
        for _ in range(int(rng.integers(0, 200))):
            ty, tx = int(rng.integers(31)), int(rng.integers(32))
            js["board"]["tiles"][ty][tx] = ["Empty", "Unpainted", "ChaseMarker", "Painted"][int(rng.integers(4))]
        for b in js["board"]["boxes"]:
            b["painted"] = bool(rng.integers(2))
        js["score"] = int(rng.integers(0, 5000))
        ref.write_state_json(i, js)
    pool.write_state_json(states)
    check_frames(pool, ref, n)
    pool.close()


def test_cuda_graph_replay_equals_stepwise_launches(tbx, oracle_mod):
    """fill_actions (frame counter in device memory) + step + render captured in a CUDA graph and replayed == the same
    steps launched one by one, and == the oracle"""
    import torch
    n, steps = 64, 40
    a = tbx.BatchedToybox("breakout", n, seeds=3)
    b = tbx.BatchedToybox("breakout", n, seeds=3)
    ref = oracle_mod.OracleBatch("breakout", n, seeds=3 + np.arange(n))
    dev = a.device
    acts_a = torch.empty(n, dtype=torch.int32, device=dev)
    acts_b = torch.empty(n, dtype=torch.int32, device=dev)
    obs_a = torch.empty((n,) + a.obs_shape, dtype=torch.uint8, device=dev)
    obs_b = torch.empty_like(obs_a)
    a.fill_random_actions(acts_a, 0xB200, 0)
    a.render(out=obs_a)                                   # everything the render path allocates exists before the capture
    t_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            a.fill_random_actions(acts_a, 0xB200, t_dev)
            t_dev += 1
            a.apply_ale_action(acts_a, auto_reset=True)
            a.render(out=obs_a)
    torch.cuda.current_stream(dev).wait_stream(side)
    legal = np.asarray(a.get_legal_action_set(), np.int32)
    for t in range(steps):
        g.replay()
        b.fill_random_actions(acts_b, 0xB200, t)
        b.apply_ale_action(acts_b, auto_reset=True)
        b.render(out=obs_b)
        ref.step(legal[[oracle_mod.action_index(0xB200, i, t, len(legal)) for i in range(n)]], auto_reset=True)
        assert torch.equal(acts_a, acts_b) and torch.equal(obs_a, obs_b), t
    assert int(t_dev.item()) == steps
    assert np.array_equal(obs_a.cpu().numpy().reshape(n, -1), ref.render("gray84").reshape(n, -1))
    a.close(); b.close()
