"""Runs the reference's OWN unit tests (test/interventions/*.py, scripts/utils/test_games.py logic) against a
ctoybox shim backed by (a) the host build of the product's engines + JSON codec and (b) the oracle.
Only possible where /root/reference exists (this container); skipped on the GPU box."""
import os
import sys
import unittest

import pytest

import shims

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def run_reference_tests(toybox_cls):
    shims.install_ctoybox(toybox_cls)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in [k for k in sys.modules if k == "test" or k.startswith("test.")]:
        del sys.modules[name]
    suite = unittest.defaultTestLoader.discover(os.path.join(REF, "test", "interventions"), top_level_dir=REF)
    res = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, "w")).run(suite)
    return res


@pytest.mark.parametrize("backend", ["emu", "oracle"])
def test_reference_intervention_tests(backend):
    res = run_reference_tests(shims.EmuToybox if backend == "emu" else shims.OracleToyboxWithSchema)
    msgs = ["%s: %s" % (t.id(), tb.strip().splitlines()[-1]) for t, tb in res.failures + res.errors]
    assert res.testsRun >= 30, res.testsRun
    assert not msgs, "\n".join(msgs)


def test_reference_smoke_script_logic():
    """scripts/utils/test_games.py:5-41 restated: score 0, lives > 0, 100 NOOPs, frames, JSON round trips."""
    for game in ("breakout", "amidar", "space_invaders"):
        tb = shims.EmuToybox(game)
        for _ in range(3):
            assert tb.get_score() == 0 and tb.get_lives() > 0
            for _ in range(100):
                tb.apply_ale_action(0)
            assert tb.get_rgb_frame().shape == (tb.get_height(), tb.get_width(), 3)
            assert tb.get_score() == 0 and tb.get_lives() > 0
            cfg, st = tb.config_to_json(), tb.to_state_json()
            tb.set_seed(1234)
            tb.write_config_json(cfg)
            tb.write_state_json(st)
            assert tb.to_state_json() == st
            tb.new_game()
