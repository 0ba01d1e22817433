"""tuning: composition of the Space Invaders direct kernel's evaluated pixels (library built with -DTBX_SI_STATS)"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import toybox_b200
from toybox_b200 import _lib
n = 65536
pool = toybox_b200.BatchedToybox("space_invaders", n, device="cuda:0", obs="gray84", seeds=(1234 + np.arange(n)) & 0xFFFFFFFF)
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2000):
    pool.step_random(0xB200, t, 0, auto_reset=True)
L = _lib.lib()
out = (C.c_ulonglong * 48)()
pool.render(); L.tbx_debug_si_stats(out, 1)
pool.render(); L.tbx_debug_si_stats(out, 1)
v = np.array(out[:], dtype=np.float64); e = v[0]
print("envs", e, "entries/env %.1f patched %.1f evaluated %.1f (conf %.2f)  pixels/env %.1f (conf %.1f)  batches/env %.2f  row-lut entries/env %.2f pixels/env %.1f" % (v[1]/e, v[2]/e, v[3]/e, v[5]/e, v[4]/e, v[6]/e, v[7]/e, v[8]/e, v[9]/e))
for i, name in enumerate(["score digits", "lives digits", "shields", "enemies", "ship", "ufo", "lasers"]):
    print("  %-13s entries/env %.2f  pixels/env %.1f  conf entries/env %.2f" % (name, v[16+i]/e, v[24+i]/e, v[32+i]/e))
