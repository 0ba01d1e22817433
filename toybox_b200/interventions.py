"""Interventions at batch N: the `with Intervention(tb) as iv:` round trip of the reference
(toybox/interventions/base.py:371-427 -- pull config + state JSON on enter, push what changed on exit) for a
list of envs of a `BatchedToybox`, plus the property-path access of toybox/interventions/core.py:271-304.

At batch size 1 the reference's own `toybox.interventions.*` classes run unchanged on `toybox_b200.ctoybox.Toybox`
(INTEGRATION.md); this module is for editing many device-resident envs in one call.
"""
import copy
import json
import re

import numpy as np


def parse_property_access(s):
    """'abc.def[7][8].y[5]' -> ['abc', 'def', 7, 8, 'y', 5]   (toybox/interventions/core.py:271-282)."""
    out = []
    for part in s.split("."):
        m = re.match(r"^([^\[\]]*)((\[\d+\])*)$", part)
        if m is None:
            raise ValueError("cannot parse property path %r" % s)
        if m.group(1):
            out.append(m.group(1))
        out.extend(int(i) for i in re.findall(r"\[(\d+)\]", m.group(2)))
    return out


def get_property(state, path):
    """Read `path` (e.g. 'bricks[1].col', 'paddle.position.x') from a decoded state."""
    cur = state
    for key in parse_property_access(path):
        cur = cur[key]
    return cur


def set_property(state, path, value):
    keys = parse_property_access(path)
    cur = state
    for key in keys[:-1]:
        cur = cur[key]
    cur[keys[-1]] = value


class BatchedIntervention:
    """Context manager over `env_ids` of a pool.

        with BatchedIntervention(pool, [3, 17, 99]) as iv:
            for s in iv.states: s["lives"] = 1
            iv.set("ufo.appearance_counter", 5)            # same edit on every listed env
            iv.config["jitter"] = 0.1                      # config edits apply to the whole pool (then new_game)

    On exit a changed config is written and every env of the pool starts a new game (the reference's rule,
    base.py:401-403); otherwise only the states that actually changed are written back (base.py:405-406).
    """

    def __init__(self, pool, env_ids=None):
        self.pool = pool
        self.env_ids = np.arange(pool.n_envs, dtype=np.int32) if env_ids is None else np.asarray(env_ids, dtype=np.int32)
        self.states = None
        self.config = None
        self.dirty_state = False
        self.dirty_config = False

    def __enter__(self):
        self.config = self.pool.config_to_json()
        self.states = self.pool.to_state_json(self.env_ids)
        self._config0 = copy.deepcopy(self.config)
        self._states0 = [json.dumps(s, sort_keys=True) for s in self.states]
        return self

    def get(self, path):
        return [get_property(s, path) for s in self.states]

    def set(self, path, value):
        for k, s in enumerate(self.states):
            set_property(s, path, value[k] if isinstance(value, (list, tuple, np.ndarray)) and len(value) == len(self.states) else value)

    def __exit__(self, exc_type, exc_value, traceback):
        if exc_type is not None:
            return False
        self.dirty_config = self.config != self._config0
        changed = [k for k, s in enumerate(self.states) if json.dumps(s, sort_keys=True) != self._states0[k]]
        self.dirty_state = bool(changed)
        if self.dirty_config:
            self.pool.write_config_json(self.config)
            self.pool.new_game()
        elif changed:
            self.pool.write_state_json([self.states[k] for k in changed], self.env_ids[changed])
        return False


# ---------------------------------------------------------------------------------------------------------------------
# The reference's per-game intervention helpers at batch N, on SoA planes instead of JSON.  `mask`
# (optional, N booleans) selects the envs an edit applies to.

def breakout_add_channel(pool, col, mask=None):
    """BreakoutIntervention.add_channel(col) (toybox/interventions/breakout.py:406-410) for every (masked) env: the
    bricks of column `col` are removed.  Bricks are column-major (test/interventions/test_get_property.py:41-44)."""
    _breakout_set_column(pool, col, 0, mask)


def breakout_fill_column(pool, col, mask=None):
    """BreakoutIntervention.fill_column(col) (breakout.py:412-416): every brick of the column alive again"""
    _breakout_set_column(pool, col, 1, mask)


def _breakout_set_column(pool, col, alive, mask):
    """one launch over the alive-mask words (tbx_breakout_columns); columns are the bricks' own `col` field, so envs whose brick
    tables were edited through JSON are handled like query_state_json('count_channels') handles them"""
    import torch
    from . import _lib
    from .pool import _ptr, _stream
    m = None
    if mask is not None:
        m = torch.as_tensor(mask, device=pool.device).to(torch.uint8).contiguous()
        if m.numel() != pool.n_envs:
            raise ValueError("expected a mask of %d entries" % pool.n_envs)
    _lib.check(pool.L.tbx_breakout_columns(pool._h, 1 if alive else 0, int(col), _ptr(m), None, _stream(pool.device)))


def breakout_channel_count(pool):
    """channels per env (columns with every brick dead), as ctoybox's breakout_channel_count query: int32[N], one launch"""
    import torch
    from . import _lib
    from .pool import _ptr, _stream
    out = torch.empty(pool.n_envs, dtype=torch.int32, device=pool.device)
    _lib.check(pool.L.tbx_breakout_columns(pool._h, 2, 0, None, _ptr(out), _stream(pool.device)))
    return out


def amidar_set_mode(pool, mode, time=None, mask=None):
    """AmidarIntervention.set_mode (toybox/interventions/amidar.py:406-419): 'jump' / 'chase' start the mode's timer
    (the config's jump_time / chase_time unless `time` is given), 'regular' clears both."""
    cfg = pool.config_to_json()
    if mode == "jump":
        pool.set_property("jump_timer", cfg["jump_time"] if time is None else time, mask)
    elif mode == "chase":
        pool.set_property("chase_timer", cfg["chase_time"] if time is None else time, mask)
    elif mode == "regular":
        pool.set_property("jump_timer", 0, mask)
        pool.set_property("chase_timer", 0, mask)
    else:
        raise ValueError("mode must be 'jump', 'chase' or 'regular'")
