"""Drop-in for the `ctoybox` python module at batch size 1: `Toybox`, `Input`, `Simulator`, `State` with
the surface the reference consumes (toybox/__init__.py:1-2, toybox/envs/atari/base.py:15-173,
toybox/envs/atari/constants.py:3-13, toybox/interventions/base.py:371-427, test/interventions/*,
scripts/utils/test_games.py, test/benchmark.py:44-58).  Every call runs on the GPU through
libtoybox_b200.so -- a `Toybox` is a `BatchedToybox` with one environment.

    import sys, toybox_b200.ctoybox
    sys.modules["ctoybox"] = toybox_b200.ctoybox      # the reference's `toybox` package now runs on this library
"""
import numpy as np

from .pool import BatchedToybox, schema_for_state, schema_for_config


class Input:
    """ctoybox.Input: six booleans plus the name constants toybox/envs/atari/constants.py reads."""
    _LEFT = "left"
    _RIGHT = "right"
    _UP = "up"
    _DOWN = "down"
    _BUTTON1 = "button1"
    _BUTTON2 = "button2"
    _NOOP = "noop"

    def __init__(self):
        self.reset()

    def reset(self):
        self.left = False
        self.right = False
        self.up = False
        self.down = False
        self.button1 = False
        self.button2 = False

    def __str__(self):
        return str(self.__dict__)

    def __repr__(self):
        return self.__str__()

    def set_input(self, input_dir, button=_NOOP):
        input_dir, button = input_dir.lower(), button.lower()
        if input_dir not in (Input._NOOP, Input._LEFT, Input._RIGHT, Input._UP, Input._DOWN):
            raise ValueError("Input value \"%s\" not found" % input_dir)
        if button not in (Input._NOOP, Input._BUTTON1, Input._BUTTON2):
            raise ValueError("Input button value \"%s\" not found" % button)
        if input_dir != Input._NOOP:
            setattr(self, input_dir, True)
        if button != Input._NOOP:
            setattr(self, button, True)

    def mask(self):
        return (1 * bool(self.left) | 2 * bool(self.right) | 4 * bool(self.up) | 8 * bool(self.down)
                | 16 * bool(self.button1) | 32 * bool(self.button2))


class _RState:
    """`toybox.rstate`: truthy handle with the two Breakout helpers baselines/run_get_seed_state.py:266-270 calls."""

    def __init__(self, tb):
        self._tb = tb

    def breakout_bricks_remaining(self):
        return self._tb.query_state_json("bricks_remaining")

    def breakout_channel_count(self):
        return self._tb.query_state_json("count_channels")

    def get_score(self):
        return self._tb.get_score()

    def lives(self):
        return self._tb.get_lives()

    def game_over(self):
        return self._tb.game_over()


class Toybox:
    """ctoybox.Toybox(game_name, grayscale=True, frameskip=0, seed=None)."""

    def __init__(self, game_name, grayscale=True, frameskip=0, seed=None, device=None):
        self.game_name = game_name
        self.grayscale = grayscale
        self.frames_per_action = frameskip + 1
        self._pool = BatchedToybox(game_name, 1, device=device, obs="gray" if grayscale else "rgba")
        self.rsimulator = self
        self.rstate = _RState(self)
        if seed is not None:
            self.set_seed(seed)
            self.new_game()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    def close(self):
        if self._pool is not None:
            self._pool.close()
            self._pool = None
            self.rstate = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- simulator-level
    def set_seed(self, seed):
        self._pool.set_seed(np.asarray([int(seed) & 0xFFFFFFFF], dtype=np.uint32))

    def new_game(self):
        self._pool.new_game()

    def get_height(self):
        return self._pool.get_height()

    def get_width(self):
        return self._pool.get_width()

    def get_legal_action_set(self):
        return sorted(self._pool.get_legal_action_set())

    # ---- transitions
    def apply_ale_action(self, action_int):
        for _ in range(self.frames_per_action):
            self._pool.apply_ale_action([int(action_int)])
            self._pool.check()     # ValueError("Expected to apply action, but failed: ...") for an unknown id

    def apply_action(self, action_input_obj):
        for _ in range(self.frames_per_action):
            self._pool.apply_action([action_input_obj.mask()])

    # ---- frames
    def get_state(self):
        return self._pool.render(obs="gray" if self.grayscale else "rgba")[0].cpu().numpy()

    def get_rgb_frame(self):
        return self._pool.render(obs="rgb")[0].cpu().numpy()

    def get_rgba_frame(self):
        return self._pool.render(obs="rgba")[0].cpu().numpy()

    def save_frame_image(self, path, grayscale=False):
        from .png import write_png
        write_png(path, self._pool.render(obs="gray" if grayscale else "rgba")[0].cpu().numpy())

    # ---- scalars
    def get_score(self):
        return int(self._pool.get_score()[0])

    def get_lives(self):
        return int(self._pool.get_lives()[0])

    def get_level(self):
        return int(self._pool.get_level()[0])

    def game_over(self):
        return self.get_lives() <= 0

    # ---- JSON
    def state_to_json(self):
        return self._pool.to_state_json([0])[0]

    to_state_json = state_to_json

    def to_json(self):
        return self.state_to_json()

    def write_state_json(self, js):
        self._pool.write_state_json([js], [0])

    def from_json(self, js):
        self.write_state_json(js)

    def config_to_json(self):
        return self._pool.config_to_json()

    def write_config_json(self, config_js):
        # ctoybox builds a fresh simulator from the config and a fresh state from it
        self._pool.write_config_json(config_js)
        self._pool.new_game()

    def query_state_json(self, query, args="null"):
        return self._pool.query_state_json(query, args, 0)

    def schema_for_state(self):
        return schema_for_state(self.game_name)

    def schema_for_config(self):
        return schema_for_config(self.game_name)


class Simulator:
    """ctoybox.Simulator: importable for `from ctoybox import Toybox, Simulator, State, Input` (toybox/__init__.py:2);
    the reference never drives it directly, so it only wraps a Toybox."""

    def __init__(self, game_name, sim=None):
        self.game_name = game_name
        self._tb = sim if sim is not None else Toybox(game_name)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._tb.close()

    def set_seed(self, seed):
        self._tb.set_seed(seed)

    def get_frame_width(self):
        return self._tb.get_width()

    def get_frame_height(self):
        return self._tb.get_height()

    def new_game(self):
        self._tb.new_game()
        return State(self)

    def to_json(self):
        return self._tb.config_to_json()

    def schema_for_state(self):
        return self._tb.schema_for_state()

    def schema_for_config(self):
        return self._tb.schema_for_config()


class State:
    """ctoybox.State view over a Simulator's current state."""

    def __init__(self, sim, state=None):
        self.sim = sim

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def lives(self):
        return self.sim._tb.get_lives()

    def score(self):
        return self.sim._tb.get_score()

    def level(self):
        return self.sim._tb.get_level()

    def game_over(self):
        return self.sim._tb.game_over()

    def render_frame(self, sim=None, grayscale=True):
        return self.sim._tb._pool.render(obs="gray" if grayscale else "rgba")[0].cpu().numpy()

    def to_json(self):
        return self.sim._tb.state_to_json()

    def query_json(self, query, args="null"):
        return self.sim._tb.query_state_json(query, args)
