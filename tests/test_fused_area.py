"""CPU tier: the fused render kernel's INTER_AREA algorithm (static base frame + dirty-rectangle recompute with
zero-padded fixed taps), emulated on the host from the same headers, equals the straightforward full-frame
resize -- over rollouts, interventions that move things around, config colour changes and other output sizes."""
import numpy as np
import pytest

import emu_lib

GAMES = ["breakout", "amidar", "space_invaders"]


@pytest.mark.parametrize("game", GAMES)
def test_fused_equals_full_over_rollout(oracle_mod, game):
    e = emu_lib.Emu(game)
    e.seed(31)
    e.new_game()
    legal = oracle_mod.LEGAL[game]
    for t in range(4000):
        e.step(ale_action=legal[oracle_mod.action_index(11, 0, t, len(legal))], auto_reset=True)
        if t % 40 == 0:
            assert np.array_equal(e.render_fast(), e.render("gray84")), t
    fresh = emu_lib.Emu(game)
    assert np.array_equal(fresh.render_fast(), oracle_mod.OracleBatch(game, 1).render("gray84")[0])


def test_fused_with_interventions_and_sizes(oracle_mod):
    b = emu_lib.Emu("breakout")
    js = b.state_json()
    js["bricks"][3]["position"]["x"] = 15.0       # overlapping bricks: the table stops being disjoint -> in-order group
    js["bricks"][3]["color"] = {"r": 255, "g": 255, "b": 255, "a": 255}
    js["balls"] += [{"position": {"x": 14.0, "y": 30.0}, "velocity": {"x": 1.0, "y": 1.0}},
                    {"position": {"x": 238.0, "y": 158.0}, "velocity": {"x": 1.0, "y": 1.0}}]
    js["paddle_width"] = 300.0
    js["score"] = 123456789
    js["lives"] = 987
    b.write_state_json(js)
    for size in ((84, 84), (100, 60), (42, 42), (128, 128)):
        assert np.array_equal(b.render_fast(*size), b.render("gray84", *size)), size
    cfg = b.config_json()
    cfg["bg_color"] = {"r": 40, "g": 50, "b": 60, "a": 255}
    cfg["frame_color"] = {"r": 200, "g": 10, "b": 10, "a": 255}
    b.write_config_json(cfg)
    b.new_game()
    assert np.array_equal(b.render_fast(), b.render("gray84"))
    s = emu_lib.Emu("space_invaders")
    js = s.state_json()
    js["shields"][0]["x"], js["shields"][0]["y"] = -5, 200        # partly off-screen
    js["enemies"][0]["x"], js["enemies"][0]["y"] = 310, -4
    js["ufo"]["appearance_counter"] = None
    js["ufo"]["x"] = 300
    js["enemy_lasers"] = [{"x": 100, "y": 100, "w": 2, "h": 8, "t": 0, "movement": "Down", "speed": 3, "color": {"r": 9, "g": 9, "b": 9, "a": 255}}]
    s.write_state_json(js)
    assert np.array_equal(s.render_fast(), s.render("gray84"))
    a = emu_lib.Emu("amidar")
    js = a.state_json()
    js["player"]["position"] = {"x": -40, "y": 2600}
    js["score"], js["lives"], js["jumps"] = 99999, 1234, 567
    for b_ in js["board"]["boxes"][:5]:
        b_["painted"] = True
    js["board"]["tiles"][6] = ["Painted"] * 32
    a.write_state_json(js)
    assert np.array_equal(a.render_fast(), a.render("gray84"))
    with pytest.raises(ValueError):
        a.render_fast(16, 16)       # more than 8 taps per axis: outside the fused kernel's limits


def test_delta_bases(oracle_mod):
    """Base frame 1 (fresh-game look) + deltas: dead bricks on the default table, painted / emptied Amidar tiles;
    an env whose brick table differs falls back to base 0 with every alive brick painted."""
    rng = np.random.default_rng(3)
    b = emu_lib.Emu("breakout")
    js = b.state_json()
    for i in rng.choice(108, size=40, replace=False):
        js["bricks"][int(i)]["alive"] = False
    b.write_state_json(js)
    assert b.n_tables() == 1                                 # still the default table -> base 1
    assert np.array_equal(b.render_fast(), b.render("gray84"))
    for br in js["bricks"]:
        br["alive"] = False
    js["bricks"][7]["alive"] = True
    b.write_state_json(js)
    assert np.array_equal(b.render_fast(), b.render("gray84"))
    js["bricks"][50]["points"] = 3                           # a different table -> base 0
    b.write_state_json(js)
    assert b.n_tables() == 2
    assert np.array_equal(b.render_fast(), b.render("gray84"))
    cfg = b.config_json()
    cfg["frame_color"] = cfg["bg_color"]                     # bricks would sit on non-distinguishable walls: still exact
    cfg["row_colors"][0] = {"r": 1, "g": 1, "b": 1, "a": 255}
    b.write_config_json(cfg)
    b.new_game()
    assert np.array_equal(b.render_fast(), b.render("gray84"))
    a = emu_lib.Emu("amidar")
    js = a.state_json()
    for ty in range(31):
        for tx in range(32):
            if rng.random() < 0.3:
                js["board"]["tiles"][ty][tx] = ["Empty", "Unpainted", "ChaseMarker", "Painted"][int(rng.integers(4))]
    a.write_state_json(js)
    assert np.array_equal(a.render_fast(), a.render("gray84"))
