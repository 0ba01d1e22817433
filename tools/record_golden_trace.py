#!/usr/bin/env python
"""record a golden trace from a REAL ctoybox, to close the "parity unpinned" gap the day a ctoybox
wheel is available (it is not in this image: third-party Rust crate ctoybox==0.5.0, REQUIREMENTS.txt:12).

Pattern: scripts/utils/start_images_toybox:24-37 (seed 1234, step, save frames).  For each game: set_seed(1234),
new_game(), then `--steps` frames of the library's counter-based action stream (seed 0xB200, env 0); every
`--every` frames the state JSON and the SHA-256 of the RGBA and grayscale frames are written to
tests/golden/ctoybox_trace_<game>.json.  tests/test_golden_trace.py replays the same actions (`replay` below) on the oracle,
on the host build of the product's engines and -- GPU tier -- on the CUDA path, and compares whenever such a file exists;
it also proves this recorder / replayer pair on traces recorded from those three implementations themselves.

    python tools/record_golden_trace.py [--steps 2000] [--every 50] [--out tests/golden]
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEGAL = {"breakout": [0, 1, 3, 4], "amidar": [0, 1, 2, 3, 4, 5, 10, 11, 12, 13], "space_invaders": [0, 1, 3, 4, 11, 12]}
MASK = (1 << 64) - 1


def action_index(seed, env, t, n_legal):
    """tbx_action_index (toybox_b200/csrc/tbx_common.h), restated so that this script needs nothing but ctoybox"""
    z = (seed + env * 0x9E3779B97F4A7C15 + t * 0xD1B54A32D192ED03) & MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    z ^= z >> 31
    return ((z >> 32) * n_legal) >> 32


def record(game, steps, every, make):
    tb = make(game)
    tb.set_seed(1234)
    tb.new_game()
    legal = list(tb.get_legal_action_set()) if hasattr(tb, "get_legal_action_set") else LEGAL[game]
    trace = {"game": game, "seed": 1234, "action_seed": 0xB200, "legal": legal, "every": every, "records": []}
    for t in range(steps + 1):
        if t % every == 0:
            tb.grayscale = False
            rgba = tb.get_state()
            tb.grayscale = True
            gray = tb.get_state()
            trace["records"].append({"t": t, "state": tb.to_state_json(), "score": tb.get_score(), "lives": tb.get_lives(),
                                     "rgba_sha256": hashlib.sha256(rgba.tobytes()).hexdigest(), "rgba_shape": list(rgba.shape),
                                     "gray_sha256": hashlib.sha256(gray.tobytes()).hexdigest(), "gray_shape": list(gray.shape)})
        if t < steps:
            tb.apply_ale_action(legal[action_index(0xB200, 0, t, len(legal))])
    return trace


def _numbers_equal(a, b):
    """JSON documents compare by value (1 == 1.0), dict order ignored"""
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_numbers_equal(a[k], b[k]) for k in a)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_numbers_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, bool) or isinstance(b, bool):
        return a is b
    return a == b


def replay(trace, make):
    """Replay `trace` on the ctoybox-compatible Toybox that make(game) returns; a list of (t, what) mismatches, empty = parity."""
    tb = make(trace["game"])
    tb.set_seed(trace["seed"])
    tb.new_game()
    recs = {r["t"]: r for r in trace["records"]}
    legal = trace["legal"]
    bad = []
    for t in range(max(recs) + 1):
        if t in recs:
            r = recs[t]
            if not _numbers_equal(tb.to_state_json(), r["state"]):
                bad.append((t, "state"))
            if (tb.get_score(), tb.get_lives()) != (r["score"], r["lives"]):
                bad.append((t, "score/lives"))
            tb.grayscale = False
            if hashlib.sha256(tb.get_state().tobytes()).hexdigest() != r["rgba_sha256"]:
                bad.append((t, "rgba frame"))
            tb.grayscale = True
            if hashlib.sha256(tb.get_state().tobytes()).hexdigest() != r["gray_sha256"]:
                bad.append((t, "gray frame"))
        tb.apply_ale_action(legal[action_index(trace["action_seed"], 0, t, len(legal))])
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    try:
        from ctoybox import Toybox
    except ImportError:
        print("ctoybox is not importable here: nothing recorded (install ctoybox==0.5.0 and re-run)", file=sys.stderr)
        return 2
    for game in LEGAL:
        trace = record(game, args.steps, args.every, lambda g: Toybox(g, grayscale=True))
        path = os.path.join(args.out, "ctoybox_trace_%s.json" % game)
        json.dump(trace, open(path, "w"))
        print("wrote", path, len(trace["records"]), "records")
    return 0


if __name__ == "__main__":
    sys.exit(main())
