"""GPU tier: the fused wrapper stack (tbx_wrap_step) against the oracle-side restatement of the reference's wrapper
chain (oracle/wrappers.py: NoopReset, MaxAndSkip, EpisodicLife, FireReset, WarpFrame, ClipReward, FrameStack + the
VecEnv worker's reset-on-done), env by env, bit-exact: stacked observations, rewards, dones, lives, scores."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GAMES = ["breakout", "amidar", "space_invaders"]


def _run(game, n, steps, kw, seed0=500, policy=None):
    import torch
    from oracle import oracle as O
    from oracle import wrappers as OW
    from toybox_b200.wrappers import DeepmindToybox
    env = DeepmindToybox(game, n, seeds=seed0, noop_seed=7, **kw)
    refs = [OW.WrappedEnv(game, seed0 + i, env_id=i, noop_seed=7, **kw) for i in range(n)]
    n_act = env.n_actions
    obs = env.reset()
    want = np.stack([r.reset() for r in refs])
    assert np.array_equal(obs[:, env.order()].cpu().numpy(), want), "reset"
    ndone = 0
    nreal = [0]
    for t in range(steps):
        acts = np.asarray([O.action_index(0xB200, i, t, n_act) for i in range(n)], np.int32)
        if policy is not None:
            acts = policy(refs, acts, t)
        obs, rew, done, info = env.step(torch.as_tensor(acts, device=env.device))
        got = obs[:, env.order()].cpu().numpy()
        rew, done = rew.cpu().numpy(), done.cpu().numpy().astype(bool)
        lives, score, real = info["lives"].cpu().numpy(), info["score"].cpu().numpy(), info["real_done"].cpu().numpy().astype(bool)
        ep_r, ep_l = info["ep_return"].cpu().numpy(), info["ep_length"].cpu().numpy()
        for i, r in enumerate(refs):
            w_obs, w_rew, w_done, w_info = r.step(int(acts[i]))
            assert w_rew == rew[i] and w_done == done[i], (game, t, i, w_rew, rew[i], w_done, done[i])
            assert w_info["lives"] == lives[i] and w_info["score"] == score[i] and w_info["real_done"] == real[i], (game, t, i)
            assert ("episode" in w_info) == bool(real[i]), (game, t, i)          # Monitor's record reaches the agent with the game over
            if real[i]:
                assert (w_info["episode"]["r"], w_info["episode"]["l"]) == (ep_r[i], ep_l[i]), (game, t, i, w_info["episode"], ep_r[i], ep_l[i])
                nreal[0] += 1
            bad = np.argwhere(got[i] != w_obs)
            assert bad.size == 0, (game, t, i, bad[:4], got[i][tuple(bad[0])], w_obs[tuple(bad[0])])
            ndone += int(w_done)
    env.check()
    env.close()
    _run.episodes = nreal[0]
    return ndone


@pytest.mark.parametrize("game", GAMES)
def test_wrapped_rollout_bit_exact(game):
    nd = _run(game, 12, 260 if game == "breakout" else 140, dict())
    if game == "breakout":
        assert nd > 0          # random play loses lives: EpisodicLife resets and full resets are exercised


def test_vec_frame_stack_mode_and_monitor_records():
    """stack_reset="zero": the stack of a finished env is zeroed and holds the reset observation only (VecFrameStack,
    vec_env/vec_frame_stack.py:17-30); whole games are played out so that Monitor's {'r', 'l'} records are compared"""
    nd = _run("breakout", 10, 900, dict(stack_reset="zero"))
    assert nd > 0 and _run.episodes > 0
    _run("space_invaders", 6, 200, dict(stack_reset="zero", frame_stack=3))
    _run("breakout", 6, 700, dict(stack_reset="zero", episode_life=False, fire_reset=False))
    assert _run.episodes > 0


def test_wrapper_switches():
    _run("breakout", 6, 120, dict(frame_skip=2, noop_max=5, episode_life=False, fire_reset=False, clip_rewards=False, frame_stack=2, size=(64, 72)))
    _run("breakout", 6, 120, dict(frame_skip=1, noop_max=0, episode_life=True, fire_reset=True, clip_rewards=True, frame_stack=3))
    _run("space_invaders", 4, 60, dict(frame_skip=3, noop_max=3, frame_stack=1))


def test_breakout_scoring_play():
    """a paddle that follows the ball: rewards > 0, bricks vanish between the two max-ed frames"""
    def policy(refs, acts, t):
        out = acts.copy()
        for i, r in enumerate(refs):
            s = r.base.b.states[0]
            if s.is_dead:
                out[i] = 1
            elif i % 3:
                bx = s.balls[0].position.x if s.n_balls else 120.0
                out[i] = 2 if bx > s.paddle.position.x + 1 else 3 if bx < s.paddle.position.x - 1 else 0
        return out
    _run("breakout", 9, 500, dict(), policy=policy)


def test_vec_env_surface():
    from toybox_b200.wrappers import ToyboxVecEnv
    venv = ToyboxVecEnv("breakout", 16, seeds=3)
    obs = venv.reset()
    assert obs.shape == (16, 84, 84, 4) and obs.dtype == np.uint8 and venv.observation_space.shape == (84, 84, 4)
    assert venv.action_space.n == 4 and venv.num_envs == 16
    episodes = 0
    for t in range(400):
        obs, rew, done, infos = venv.step(np.full(16, t % 4, np.int32))
        assert obs.shape == (16, 84, 84, 4) and rew.dtype == np.float32 and done.dtype == bool and len(infos) == 16
        episodes += sum(1 for i in infos if "episode" in i)
        for i in np.flatnonzero(done):                       # VecFrameStack: a finished env's stack is [0, 0, 0, reset observation]
            assert not obs[i, :, :, :3].any() and obs[i, :, :, 3].any()
    assert episodes > 0
    venv.close()


def test_vec_env_monitor_parity_and_csv(tmp_path):
    """infos[i]['episode'] {r, l} of ToyboxVecEnv against the oracle-side Monitor (bench/monitor.py:58-76) env by env, and the
    Monitor-compatible CSV (a '#' JSON header, then r,l,t rows) read back the way baselines' load_results does"""
    import json
    from oracle import oracle as O
    from oracle import wrappers as OW
    from toybox_b200.wrappers import ToyboxVecEnv
    n = 8
    path = str(tmp_path / "0")
    venv = ToyboxVecEnv("breakout", n, seeds=40, noop_seed=9, monitor_file=path)
    refs = [OW.WrappedEnv("breakout", 40 + i, env_id=i, noop_seed=9, stack_reset="zero") for i in range(n)]
    obs = venv.reset()
    want = np.stack([r.reset() for r in refs])
    assert np.array_equal(obs, want.transpose(0, 2, 3, 1))
    rows = []
    for t in range(1200):
        acts = np.asarray([O.action_index(0xB200, i, t, 4) for i in range(n)], np.int32)
        obs, rew, done, infos = venv.step(acts)
        for i, r in enumerate(refs):
            w_obs, w_rew, w_done, w_info = r.step(int(acts[i]))
            assert np.array_equal(obs[i], w_obs.transpose(1, 2, 0)) and w_rew == rew[i] and w_done == done[i], (t, i)
            assert ("episode" in infos[i]) == ("episode" in w_info), (t, i)
            if "episode" in w_info:
                assert infos[i]["episode"]["r"] == w_info["episode"]["r"] and infos[i]["episode"]["l"] == w_info["episode"]["l"], (t, i)
                rows.append((w_info["episode"]["r"], w_info["episode"]["l"]))
    venv.close()
    assert rows and venv.get_episode_rewards() == [float(r) for r, _ in rows] and venv.get_episode_lengths() == [l for _, l in rows]
    lines = open(path + ".monitor.csv").read().splitlines()
    assert lines[0].startswith("#") and "t_start" in json.loads(lines[0][1:]) and lines[1] == "r,l,t"
    got = [(float(x.split(",")[0]), int(x.split(",")[1])) for x in lines[2:]]
    assert got == [(float(r), l) for r, l in rows]
