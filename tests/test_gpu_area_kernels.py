"""GPU tier: every code path of the INTER_AREA (WarpFrame) render -- the warp-per-env tile kernel with its list,
sweep and per-tile rebuild paths, and the CTA-canvas kernel -- against the oracle's cv2-exact resize, on mid-game
states of all three games and on several output sizes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GAMES = ["breakout", "amidar", "space_invaders"]
VARIANTS = [{}, {"TBX_AREA_LCAP": "24"}, {"TBX_AREA_LCAP": "3"}, {"TBX_AREA_KERNEL": "cta"}, {"TBX_AREA_THREADS": "128"},
            {"TBX_AREA_TILE_H": "4", "TBX_AREA_MAX_RUN": "8"}, {"TBX_AREA_TILE_H": "16"}, {"TBX_AREA_MAX_RUN": "1", "TBX_AREA_LCAP": "40"}]
SIZES = [(84, 84), (96, 80), (64, 64), (48, 60), (100, 37)]


def _advance(tbx, oracle_mod, game, n, steps, seed):
    pool = tbx.BatchedToybox(game, n, seeds=seed)
    ref = oracle_mod.OracleBatch(game, n, seeds=seed + np.arange(n))
    legal = np.asarray(oracle_mod.LEGAL[game], np.int32)
    for t in range(steps):
        acts = legal[[oracle_mod.action_index(0xB200, i, t, len(legal)) for i in range(n)]]
        if game == "breakout":          # track the ball with an aim error so that bricks get broken
            for i in range(n):
                s = ref.states[i]
                if s.is_dead:
                    acts[i] = 1
                elif i % 4:
                    bx = s.balls[0].position.x if s.n_balls else 120.0
                    off = ((i * 7 + t // 50) % 9) - 4
                    acts[i] = 3 if bx + off > s.paddle.position.x + 1 else 4 if bx + off < s.paddle.position.x - 1 else 0
        pool.apply_ale_action(acts, auto_reset=True)
        ref.step(acts, auto_reset=True)
    return pool, ref


def _set_env(v):
    for k in ("TBX_AREA_LCAP", "TBX_AREA_KERNEL", "TBX_AREA_THREADS", "TBX_AREA_TILE_H", "TBX_AREA_MAX_RUN"):
        os.environ.pop(k, None)
    os.environ.update(v)


@pytest.mark.parametrize("game", GAMES)
def test_area_render_paths_bit_exact(tbx, oracle_mod, game):
    n = 45                                   # ragged: the last chunk holds 5 envs
    pool, ref = _advance(tbx, oracle_mod, game, n, 1200 if game == "breakout" else 400, 77)
    try:
        for ow, oh in SIZES:
            want = ref.render("gray84", ow, oh).reshape(n, -1)
            for v in VARIANTS:
                _set_env(v)
                got = pool.render(obs=("gray_area", ow, oh)).cpu().numpy().reshape(n, -1)
                bad = np.argwhere(got != want)
                assert bad.size == 0, (game, ow, oh, v, bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])
    finally:
        _set_env({})
        pool.close()


@pytest.mark.parametrize("game", GAMES)
def test_native_render_paths_bit_exact(tbx, oracle_mod, game):
    """native layouts: broadcast + patch with its dense-env hand-over to the canvas kernel (never / always / mixed), and the
    canvas kernel alone"""
    n = 45
    pool, ref = _advance(tbx, oracle_mod, game, n, 1200 if game == "breakout" else 400, 78)
    keys = ("TBX_NATIVE_KERNEL", "TBX_NATIVE_DENSE")
    try:
        for mode in ("rgb", "rgba", "gray"):
            want = ref.render(mode).reshape(n, -1)
            for v in ({}, {"TBX_NATIVE_DENSE": "-1"}, {"TBX_NATIVE_DENSE": "0"}, {"TBX_NATIVE_DENSE": "3"}, {"TBX_NATIVE_KERNEL": "canvas"}):
                for k in keys:
                    os.environ.pop(k, None)
                os.environ.update(v)
                got = pool.render(obs=mode).cpu().numpy().reshape(n, -1)
                bad = np.argwhere(got != want)
                assert bad.size == 0, (game, mode, v, bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])
    finally:
        for k in keys:
            os.environ.pop(k, None)
        pool.close()
