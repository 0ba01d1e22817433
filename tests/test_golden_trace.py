"""the golden-trace recorder (tools/record_golden_trace.py) and its replayer, proven end to end before a real
ctoybox wheel ever shows up: traces are recorded from the oracle, from the host build of the product's engines and -- GPU
tier -- from the CUDA path (toybox_b200.ctoybox.Toybox), written to a temporary directory, read back and replayed on the
other implementations; a corrupted trace must be caught.  Whenever tests/golden/ctoybox_trace_<game>.json exists (recorded
from ctoybox==0.5.0 somewhere else) the same replayer judges the oracle, the host build and the CUDA path against it --
that is what would lift "parity unpinned" (oracle/SPEC.md)."""
import importlib.util
import json
import os

import pytest

GAMES = ["breakout", "amidar", "space_invaders"]
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def recorder():
    spec = importlib.util.spec_from_file_location("record_golden_trace", os.path.join(os.path.dirname(HERE), "tools", "record_golden_trace.py"))
    rec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rec)
    return rec


def _oracle(game):
    from oracle import oracle as O
    return O.OracleToybox(game)


def _emu(game):
    import shims
    return shims.EmuToybox(game)


def _cuda(game):
    from toybox_b200 import ctoybox
    return ctoybox.Toybox(game)


def _round_trip(tmp_path, game, make_rec, make_replay, steps=400, every=40):
    rec = recorder()
    trace = rec.record(game, steps, every, make_rec)
    path = tmp_path / ("ctoybox_trace_%s.json" % game)
    json.dump(trace, open(path, "w"))
    back = json.load(open(path))
    assert len(back["records"]) == steps // every + 1
    assert rec.replay(back, make_replay) == []
    return rec, back


@pytest.mark.parametrize("game", GAMES)
def test_recorder_and_replayer_on_cpu_implementations(oracle_mod, tmp_path, game):
    rec, trace = _round_trip(tmp_path, game, _oracle, _emu)
    assert rec.replay(trace, _oracle) == []
    # the replayer really compares: a changed score, a changed state field and a changed frame hash are all reported
    trace["records"][3]["score"] += 1
    trace["records"][5]["state"]["lives"] += 1
    trace["records"][7]["gray_sha256"] = "0" * 64
    bad = rec.replay(trace, _oracle)
    assert (trace["records"][3]["t"], "score/lives") in bad and (trace["records"][5]["t"], "state") in bad and (trace["records"][7]["t"], "gray frame") in bad
    assert list(oracle_mod.LEGAL[game]) == rec.LEGAL[game]
    for t in (0, 1, 7, 1000, 123456):
        assert rec.action_index(0xB200, 5, t, len(rec.LEGAL[game])) == oracle_mod.action_index(0xB200, 5, t, len(rec.LEGAL[game]))


@pytest.mark.gpu
@pytest.mark.parametrize("game", GAMES)
def test_recorder_and_replayer_on_the_cuda_path(tbx, oracle_mod, tmp_path, game):
    rec, trace = _round_trip(tmp_path, game, _cuda, _oracle)      # recorded on the GPU, replayed on the oracle
    _round_trip(tmp_path, game, _oracle, _cuda, steps=300, every=30)   # and the other way round


def _committed(game):
    path = os.path.join(GOLD, "ctoybox_trace_%s.json" % game)
    if not os.path.exists(path):
        pytest.skip("no ctoybox trace recorded (ctoybox==0.5.0 is not installable here): transition / raster parity stays unpinned")
    return json.load(open(path))


@pytest.mark.parametrize("game", GAMES)
@pytest.mark.parametrize("impl", ["oracle", "emu"])
def test_committed_ctoybox_trace(oracle_mod, game, impl):
    assert recorder().replay(_committed(game), _oracle if impl == "oracle" else _emu) == []


@pytest.mark.gpu
@pytest.mark.parametrize("game", GAMES)
def test_committed_ctoybox_trace_on_the_cuda_path(tbx, game):
    assert recorder().replay(_committed(game), _cuda) == []
