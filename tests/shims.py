"""ctoybox look-alikes for the CPU tier, so the reference's own python (toybox.interventions, its unit tests)
can be exercised without a GPU: `EmuToybox` drives the host build of the PRODUCT's engine headers and JSON
codec (tests/emu), `oracle.OracleToybox` drives the oracle.  Test infrastructure only."""
import json
import sys
import types

import emu_lib
from oracle import oracle as O


class EmuToybox:
    def __init__(self, game_name, grayscale=True, frameskip=0, seed=None):
        self.game_name, self.grayscale = game_name, grayscale
        self.e = emu_lib.Emu(game_name)
        self.rstate = self
        if seed is not None:
            self.set_seed(seed)
            self.new_game()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def set_seed(self, seed):
        self.e.seed(int(seed))

    def new_game(self):
        self.e.new_game()

    def get_width(self):
        return emu_lib.DIMS[self.game_name][0]

    def get_height(self):
        return emu_lib.DIMS[self.game_name][1]

    def get_legal_action_set(self):
        return list(O.LEGAL[self.game_name])

    def apply_ale_action(self, a):
        self.e.step(ale_action=a)

    def apply_action(self, inp):
        self.e.step(input_mask=inp.mask())

    def get_state(self):
        f = self.e.render("gray" if self.grayscale else "rgba")
        return f[:, :, None] if f.ndim == 2 else f

    def get_rgb_frame(self):
        return self.e.render("rgb")

    def get_score(self):
        return self.e.state_json()["score"]

    def get_lives(self):
        return self.e.state_json()["lives"]

    def game_over(self):
        return self.get_lives() <= 0

    def to_state_json(self):
        return self.e.state_json()

    state_to_json = to_state_json
    to_json = to_state_json

    def write_state_json(self, js):
        self.e.write_state_json(js)

    def config_to_json(self):
        return self.e.config_json()

    def write_config_json(self, js):
        self.e.write_config_json(js)
        self.e.new_game()

    def schema_for_state(self):
        return emu_lib.schema_for_state(self.game_name)

    def schema_for_config(self):
        return emu_lib.schema_for_config(self.game_name)

    def query_state_json(self, query, args="null"):
        L = emu_lib.lib()
        L.emu_query_json.restype = emu_lib.C.c_void_p
        L.emu_query_json.argtypes = [emu_lib.C.c_void_p, emu_lib.C.c_char_p, emu_lib.C.c_char_p]
        a = args if isinstance(args, str) else json.dumps(args)
        return json.loads(emu_lib._take(L.emu_query_json(self.e.h, query.encode(), a.encode())))

    def breakout_bricks_remaining(self):
        return self.query_state_json("bricks_remaining")

    def breakout_channel_count(self):
        return self.query_state_json("count_channels")


class OracleToyboxWithSchema(O.OracleToybox):
    def schema_for_state(self):
        return emu_lib.schema_for_state(self.game_name)


def install_ctoybox(toybox_cls):
    """Make `import ctoybox` resolve to a shim whose Toybox is `toybox_cls`."""
    m = types.ModuleType("ctoybox")
    m.Toybox, m.Input, m.Simulator, m.State = toybox_cls, O.Input, object, object
    sys.modules["ctoybox"] = m
    for name in [k for k in sys.modules if k == "toybox" or k.startswith("toybox.")]:
        del sys.modules[name]
    return m
