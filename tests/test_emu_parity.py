"""CPU tier: the PRODUCT's engine headers (the functions the CUDA kernels are compiled from), built for the host
by tests/emu, against the oracle: transitions, env-level bookkeeping, draw lists in every layout, INTER_AREA,
JSON codecs, interventions.  Bit-exact."""
import numpy as np
import pytest

import emu_lib
from conftest import json_diff

GAMES = ["breakout", "amidar", "space_invaders"]


def rollout(oracle_mod, game, seed, steps, action_seed, check_every, policy=None):
    o = oracle_mod.OracleBatch(game, 1, seeds=[seed])
    e = emu_lib.Emu(game)
    e.seed(seed)
    e.new_game()
    legal = oracle_mod.LEGAL[game]
    for t in range(steps):
        a = policy(o, t) if policy else legal[oracle_mod.action_index(action_seed, 0, t, len(legal))]
        r, d, s, l = o.step([a], auto_reset=True)
        assert (int(r[0]), bool(d[0]), int(s[0]), int(l[0])) == e.step(ale_action=a, auto_reset=True), t
        if t % check_every == 0 or t == steps - 1:
            assert json_diff(e.state_json(), o.state_json(0)) == [], t
            for mode in ("rgba", "rgb", "gray", "gray84"):
                assert np.array_equal(e.render(mode), o.render(mode)[0]), (t, mode)
    return e, o


@pytest.mark.parametrize("game", GAMES)
@pytest.mark.parametrize("seed", [1234, 7, 4000000000])
def test_random_rollout(oracle_mod, game, seed):
    rollout(oracle_mod, game, seed, 6000, 0xB200 + seed, 250)


def test_breakout_tracking_policy(oracle_mod):
    def policy(o, t):
        s = o.states[0]
        if s.is_dead:
            return 1
        bx = s.balls[0].position.x + ((t // 40) % 9 - 4)
        return 3 if bx > s.paddle.position.x + 1 else 4 if bx < s.paddle.position.x - 1 else 0
    e, o = rollout(oracle_mod, "breakout", 42, 20000, 0, 500, policy)
    assert o.states[0].score > 50 or o.states[0].level > 1


def test_amidar_wandering_policy(oracle_mod):
    """Holds a direction for a while so the player paints segments, fills boxes and meets enemies."""
    dirs = [2, 3, 5, 4]
    e, o = rollout(oracle_mod, "amidar", 5, 20000, 0, 500, lambda o_, t: dirs[(t // 37 + (t // 500)) % 4] if t % 211 else 1)
    assert o.states[0].score > 0


def test_space_invaders_shooting_policy(oracle_mod):
    e, o = rollout(oracle_mod, "space_invaders", 9, 20000, 0, 500, lambda o_, t: [11, 12, 1, 11, 12, 1, 3, 4][(t // 15) % 8])
    assert o.states[0].score > 0


@pytest.mark.parametrize("game", GAMES)
def test_all_18_ale_actions_and_invalid(oracle_mod, game):
    o = oracle_mod.OracleBatch(game, 1)
    e = emu_lib.Emu(game)
    for t in range(360):
        a = t % 18
        o.step([a], auto_reset=True)
        e.step(ale_action=a, auto_reset=True)
    assert json_diff(e.state_json(), o.state_json(0)) == []
    with pytest.raises(ValueError):
        e.step(ale_action=18)
    with pytest.raises(ValueError):
        e.step(ale_action=-1)


def test_interventions_step_like_oracle(oracle_mod):
    edits = {
        "breakout": lambda js: (js["bricks"][5].update(alive=False), js["bricks"][30].update(color={"r": 9, "g": 8, "b": 7, "a": 255}),
                                js["bricks"][31].update(destructible=False), js["bricks"][32]["position"].update(x=100.5),
                                js["balls"].append({"position": {"x": 60.0, "y": 90.0}, "velocity": {"x": -1.5, "y": -1.5}}),
                                js.update(paddle_width=48.0, lives=2)),
        "amidar": lambda js: (js.update(jumps=2, chase_timer=50), js["enemies"].pop(),
                              js["enemies"][0].update(ai={"EnemyTargetPlayer": {"start": {"tx": 0, "ty": 0}, "start_dir": "Down", "vision_distance": 15,
                                                                               "dir": "Down", "player_seen": None}}),
                              js["enemies"][1].update(ai={"EnemyAmidarMvmt": {"vert": "Down", "horiz": "Right", "start_vert": "Down",
                                                                             "start_horiz": "Right", "start": {"tx": 0, "ty": 0}}}),
                              js["enemies"][2].update(ai={"EnemyPerimeterAI": {"start": {"tx": 0, "ty": 0}}}),
                              js["enemies"][3].update(ai={"EnemyRandomMvmt": {"start": {"tx": 31, "ty": 30}, "start_dir": "Up", "dir": "Up"}},
                                                      position={"x": 1984, "y": 2400}),
                              js["player"]["position"].update(x=1000, y=480)),
        "space_invaders": lambda js: (js.update(lives=1, enemy_shot_delay=2), js["ufo"].update(appearance_counter=1),
                                      js["shields"][1]["data"][3][4].update(a=0), js["enemies"][35].update(alive=False),
                                      js["ship"].update(x=100, color={"r": 1, "g": 2, "b": 3, "a": 255})),
    }
    for game in GAMES:
        o = oracle_mod.OracleBatch(game, 1, seeds=[77])
        e = emu_lib.Emu(game)
        e.seed(77)
        e.new_game()
        legal = oracle_mod.LEGAL[game]
        for t in range(300):
            a = legal[oracle_mod.action_index(3, 0, t, len(legal))]
            o.step([a]); e.step(ale_action=a, auto_reset=True)
        js = e.state_json()
        edits[game](js)
        e.write_state_json(js)
        o.write_state_json(0, js)
        assert json_diff(e.state_json(), o.state_json(0)) == [], game
        for t in range(3000):
            a = legal[oracle_mod.action_index(4, 0, t, len(legal))]
            r, d, s, l = o.step([a], auto_reset=True)
            assert (int(r[0]), bool(d[0]), int(s[0]), int(l[0])) == e.step(ale_action=a, auto_reset=True), (game, t)
            if t % 300 == 0:
                assert json_diff(e.state_json(), o.state_json(0)) == [], (game, t)
                assert np.array_equal(e.render("rgba"), o.render("rgba")[0]), (game, t)


def test_config_interventions(oracle_mod):
    """write_config_json + new_game (toybox/interventions/base.py:401-403) with edited configs."""
    cases = {
        "breakout": dict(start_lives=2, paddle_discrete_segments=7, ball_speed_slow=3.0, row_scores=[9, 9, 5, 5, 2, 2],
                         bg_color={"r": 10, "g": 20, "b": 30, "a": 255}),
        "amidar": dict(start_lives=1, start_jumps=2, chase_time=20, jump_time=10, box_bonus=7),
        "space_invaders": dict(start_lives=1, jitter=0.1, enemy_protocol="Random", row_scores=[1, 2, 3, 4, 5, 6]),
    }
    for game, edit in cases.items():
        e = emu_lib.Emu(game)
        cfg = e.config_json()
        cfg.update(edit)
        e.write_config_json(cfg)
        e.new_game()
        o = oracle_mod.OracleBatch(game, 1, cfg_json=cfg)
        o.new_game()
        assert json_diff(e.config_json(), oracle_mod.CODEC[game][2](o.cfg)) == []
        legal = oracle_mod.LEGAL[game]
        e.write_state_json(o.state_json(0))        # align the rng lineages, then compare dynamics under the new config
        for t in range(3000):
            a = legal[oracle_mod.action_index(8, 0, t, len(legal))]
            r, d, s, l = o.step([a], auto_reset=False)
            assert (int(r[0]), bool(d[0]), int(s[0]), int(l[0])) == e.step(ale_action=a), (game, t)
        assert json_diff(e.state_json(), o.state_json(0)) == [], game
        assert np.array_equal(e.render("rgb"), o.render("rgb")[0]), game


def test_bad_json_is_rejected_and_leaves_state_untouched():
    e = emu_lib.Emu("breakout")
    before = e.state_json()
    js = e.state_json()
    js["balls"] = js["balls"] * 5
    with pytest.raises(ValueError):
        e.write_state_json(js)
    del js["balls"]
    with pytest.raises(ValueError):
        e.write_state_json(js)
    assert e.state_json() == before
    a = emu_lib.Emu("amidar")
    js = a.state_json()
    js["enemies"] = js["enemies"] * 2
    with pytest.raises(ValueError):
        a.write_state_json(js)


def test_brick_tables_are_interned():
    e = emu_lib.Emu("breakout")
    js = e.state_json()
    e.write_state_json(js)
    assert e.n_tables() == 1                      # unchanged geometry maps onto the default table
    js["bricks"][0]["points"] = 99
    e.write_state_json(js)
    e.write_state_json(js)
    assert e.n_tables() == 2


def test_schema_has_the_keys_the_intervention_classes_require():
    want = {"breakout": {"score", "lives", "rand", "level", "paddle", "paddle_width", "paddle_speed", "ball_radius", "balls", "bricks",
                         "reset", "is_dead"},
            "amidar": {"score", "lives", "rand", "level", "enemies", "player", "jumps", "jump_timer", "chase_timer", "board"},
            "space_invaders": {"score", "lives", "rand", "level", "ship", "ship_laser", "enemies", "enemies_movement", "enemy_lasers",
                               "shields", "ufo", "life_display_timer", "enemy_shot_delay"}}
    for game, keys in want.items():
        sch = emu_lib.schema_for_state(game)
        assert set(sch["required"]) == keys == set(sch["properties"]) == set(emu_lib.Emu(game).state_json())
        assert sch["properties"]["lives"]["type"] == "integer"
        assert set(emu_lib.schema_for_config(game)["required"]) == set(emu_lib.Emu(game).config_json())
