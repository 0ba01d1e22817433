"""GPU tier: vectorised property access (tbx_field_get / tbx_field_set) against the JSON state of the same
envs and against the oracle after the same edits."""
import numpy as np
import pytest
import torch

from conftest import json_diff

pytestmark = pytest.mark.gpu

PATHS = {
    "breakout": ["lives", "score", "level", "paddle.position.x", "paddle.velocity.x", "paddle_width", "ball_radius", "is_dead", "reset",
                 "balls[0].position.y", "balls[0].velocity.x", "bricks[0].alive", "bricks[37].alive", "bricks[107].alive"],
    "space_invaders": ["lives", "score", "ship.x", "ship.alive", "ship.death_counter", "ufo.appearance_counter", "ufo.x", "life_display_timer",
                       "enemy_shot_delay", "enemies[0].x", "enemies[35].y", "enemies[17].alive", "enemies[3].death_counter", "enemies[9].points",
                       "enemies_movement.move_counter", "shields[2].x"],
    "amidar": ["lives", "score", "jumps", "jump_timer", "chase_timer", "player.position.x", "player.position.y", "player.caught",
               "enemies[0].position.x", "enemies[4].position.y", "enemies[2].speed", "board.boxes[0].painted", "board.boxes[28].painted"],
}


def _dig(state, path):
    from toybox_b200.interventions import get_property
    return get_property(state, path)


@pytest.mark.parametrize("game", list(PATHS))
def test_get_matches_json(tbx, oracle_mod, game):
    n = 40
    pool = tbx.BatchedToybox(game, n, seeds=21)
    legal = np.asarray(pool.get_legal_action_set(), np.int32)
    for t in range(300):
        pool.apply_ale_action(legal[[oracle_mod.action_index(7, i, t, len(legal)) for i in range(n)]], auto_reset=True)
    states = pool.to_state_json()
    for path in PATHS[game]:
        got = pool.get_property(path).cpu().numpy()
        for i in range(n):
            want = _dig(states[i], path)
            if want is None:
                assert got[i] == -2 ** 31, (path, i)
            else:
                assert got[i] == want, (path, i, got[i], want)
    with pytest.raises(Exception):
        pool.get_property("no.such.field")
    pool.close()


def test_set_matches_json_edit_and_rollout(tbx, oracle_mod):
    """masked vectorised edits == the same edits through the oracle's JSON, and the games then evolve identically"""
    from toybox_b200 import interventions as IV
    n = 24
    pool = tbx.BatchedToybox("breakout", n, seeds=5)
    ref = oracle_mod.OracleBatch("breakout", n, seeds=5 + np.arange(n))
    legal = np.asarray(pool.get_legal_action_set(), np.int32)

    def both_step(t):
        acts = legal[[oracle_mod.action_index(9, i, t, len(legal)) for i in range(n)]]
        pool.apply_ale_action(acts, auto_reset=True)
        ref.step(acts, auto_reset=True)

    for t in range(50):
        both_step(t)
    mask = np.arange(n) % 3 == 0
    pool.set_property("lives", 2, mask)
    xs = np.linspace(40.0, 200.0, n)
    pool.set_property("paddle.position.x", xs, mask)
    pool.set_property("score", 77, mask)
    IV.breakout_add_channel(pool, 4, mask)
    for i in np.nonzero(mask)[0]:
        js = ref.state_json(int(i))
        js["lives"] = 2
        js["paddle"]["position"]["x"] = float(xs[i])
        js["score"] = 77
        for b in js["bricks"]:
            if b["col"] == 4:
                b["alive"] = False
        ref.write_state_json(int(i), js)
    states = pool.to_state_json()
    for i in range(n):
        assert json_diff(states[i], ref.state_json(i)) == [], i
    assert np.array_equal(IV.breakout_channel_count(pool).cpu().numpy(), mask.astype(np.int32))
    for t in range(50, 250):
        both_step(t)
    states = pool.to_state_json()
    for i in range(n):
        assert json_diff(states[i], ref.state_json(i)) == [], i
    assert np.array_equal(pool.render(obs="gray84").cpu().numpy().reshape(n, -1), ref.render("gray84").reshape(n, -1))
    pool.close()


def test_option_and_amidar_mode(tbx, oracle_mod):
    from toybox_b200 import interventions as IV
    si = tbx.BatchedToybox("space_invaders", 8, seeds=1)
    si.set_property("ufo.appearance_counter", None, [1, 0, 0, 0, 0, 0, 0, 1])
    si.set_property("ufo.appearance_counter", 5, [0, 1, 0, 0, 0, 0, 0, 0])
    st = si.to_state_json()
    assert st[0]["ufo"]["appearance_counter"] is None and st[7]["ufo"]["appearance_counter"] is None and st[1]["ufo"]["appearance_counter"] == 5
    si.close()
    am = tbx.BatchedToybox("amidar", 6, seeds=1)
    IV.amidar_set_mode(am, "jump", mask=[1, 1, 0, 0, 0, 0])
    IV.amidar_set_mode(am, "chase", time=33, mask=[0, 0, 1, 0, 0, 0])
    jt, ct = am.get_property("jump_timer").cpu().numpy(), am.get_property("chase_timer").cpu().numpy()
    cfg = am.config_to_json()
    assert list(jt) == [cfg["jump_time"]] * 2 + [0] * 4 and ct[2] == 33 and ct[0] == 0
    am.close()
