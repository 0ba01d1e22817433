"""Property tests (hypothesis): JSON round trip = identity.  For states reached by random rollouts and
then edited at random leaves (numbers, booleans, colours, Option fields, list lengths within the capacity limits):
    write_state_json(js); to_state_json() == js
on the product's host codec (tests/emu: the same tbx_host.cpp the CUDA library links) and on the oracle, and both agree;
the GPU tier repeats it through the device planes (gather / scatter kernels) of a BatchedToybox."""
import copy

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import emu_lib
from conftest import json_diff

GAMES = ["breakout", "amidar", "space_invaders"]
_BASE = {}


def base_states(oracle_mod, game):
    """a few states along a rollout (fresh, mid-game), computed once"""
    if game not in _BASE:
        o = oracle_mod.OracleBatch(game, 1, seeds=[99])
        legal = oracle_mod.LEGAL[game]
        out = [o.state_json(0)]
        for t in range(900):
            o.step([legal[oracle_mod.action_index(3, 0, t, len(legal))]], auto_reset=True)
            if t % 300 == 299:
                out.append(o.state_json(0))
        _BASE[game] = out
    return _BASE[game]


def leaves(js, path=()):
    """paths of the editable scalar leaves of a state document"""
    if isinstance(js, dict):
        for k, v in js.items():
            if k in ("rand", "data"):                # u64 pairs and shield pixels (opaque / transparent only): edited separately
                continue
            yield from leaves(v, path + (k,))
    elif isinstance(js, list):
        for i, v in enumerate(js):
            yield from leaves(v, path + (i,))
    elif isinstance(js, (bool, int, float)) or js is None:
        yield path


def get(js, path):
    for k in path:
        js = js[k]
    return js


def put(js, path, v):
    for k in path[:-1]:
        js = js[k]
    js[path[-1]] = v


def edit(js, game, data):
    """apply a handful of type-preserving random edits"""
    js = copy.deepcopy(js)
    paths = list(leaves(js))
    for _ in range(data.draw(st.integers(0, 12))):
        p = paths[data.draw(st.integers(0, len(paths) - 1))]
        old = get(js, p)
        key = p[-1]
        if isinstance(old, bool):
            new = data.draw(st.booleans())
        elif key in ("r", "g", "b", "a"):
            new = data.draw(st.integers(0, 255))
        elif isinstance(old, float):
            new = float(data.draw(st.floats(-500.0, 500.0, allow_nan=False, width=64)))
        elif old is None:
            new = data.draw(st.one_of(st.none(), st.integers(1, 400)))
        elif key in ("tx", "ty", "row", "col", "id", "depth"):
            new = data.draw(st.integers(0, 30))
        elif key in ("lives", "jumps", "level", "levels_completed"):
            new = data.draw(st.integers(0, 9))
        else:
            new = data.draw(st.integers(-300, 3000)) if not (game == "space_invaders" and key in ("w", "h")) else old
        put(js, p, new)
    js["rand"]["state"] = [data.draw(st.integers(0, 2 ** 64 - 1)), data.draw(st.integers(1, 2 ** 64 - 1))]
    if game == "breakout":
        k = data.draw(st.integers(0, 4))
        js["balls"] = (js["balls"] + [{"position": {"x": 50.0, "y": 60.0}, "velocity": {"x": -1.5, "y": 2.0}}] * 4)[:k]
    elif game == "amidar":
        js["enemies"] = js["enemies"][:data.draw(st.integers(0, len(js["enemies"])))]
    else:
        js["enemy_lasers"] = js["enemy_lasers"][:data.draw(st.integers(0, len(js["enemy_lasers"])))]
        opaque = next(px for sh in js["shields"] for row in sh["data"] for px in row if px["a"])
        for _ in range(data.draw(st.integers(0, 6))):          # a shield pixel is opaque (the shield's colour) or transparent
            sh = js["shields"][data.draw(st.integers(0, 2))]
            row = sh["data"][data.draw(st.integers(0, len(sh["data"]) - 1))]
            row[data.draw(st.integers(0, len(row) - 1))] = dict(opaque) if data.draw(st.booleans()) else {"r": 0, "g": 0, "b": 0, "a": 0}
    return js


@pytest.mark.parametrize("game", GAMES)
@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(data=st.data())
def test_round_trip_identity_host_codec_and_oracle(oracle_mod, game, data):
    base = base_states(oracle_mod, game)
    js = edit(base[data.draw(st.integers(0, len(base) - 1))], game, data)
    e = emu_lib.Emu(game)
    o = oracle_mod.OracleBatch(game, 1, seeds=[1])
    try:
        e.write_state_json(js)
    except ValueError:                               # outside the product's stated limits (SPEC.md B8 / A8 / S14, e.g. a box
        return                                       # corner off the board): refused with a message, nothing to round-trip
    o.write_state_json(0, js)
    back = e.state_json()
    assert json_diff(back, js) == []
    assert json_diff(o.state_json(0), js) == []
    e.write_state_json(back)
    assert json_diff(e.state_json(), back) == []     # and once more: a fixed point


@pytest.mark.gpu
@pytest.mark.parametrize("game", GAMES)
@settings(max_examples=8, deadline=None, suppress_health_check=list(HealthCheck))
@given(data=st.data())
def test_round_trip_identity_device_planes(tbx, oracle_mod, game, data):
    base = base_states(oracle_mod, game)
    n = 24
    docs = [edit(base[data.draw(st.integers(0, len(base) - 1))], game, data) for _ in range(n)]
    ref = emu_lib.Emu(game)
    ok = []
    for d in docs:                                    # keep the documents the codec accepts
        try:
            ref.write_state_json(d)
            ok.append(d)
        except ValueError:
            pass
    if not ok:
        return
    pool = tbx.BatchedToybox(game, len(ok), seeds=5)
    try:
        pool.write_state_json(ok)
        back = pool.to_state_json()
        for a, b in zip(back, ok):
            assert json_diff(a, b) == []
        ids = list(range(0, len(ok), 3))
        text = pool.to_state_json_text(ids)           # the text-level API carries the same documents
        pool.write_state_json(text, ids)
        for a, b in zip(pool.to_state_json(), ok):
            assert json_diff(a, b) == []
    finally:
        pool.close()
