"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol include/toybox_b200.h declares.
No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "toybox_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tbx_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    import toybox_b200
    toybox_b200.build()
    L = toybox_b200.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(toybox_b200._lib.EXPORTS) == names
    assert L.tbx_version() >= 100


def test_no_cpu_fallback_without_a_gpu():
    import torch
    import toybox_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(toybox_b200.ToyboxError):
        toybox_b200.BatchedToybox("breakout", 4)
    out = C.c_void_p()
    rc = toybox_b200.lib().tbx_pool_create(b"breakout", 4, 0, None, C.byref(out))
    assert rc != 0 and not out.value and b"no CPU path" in toybox_b200.lib().tbx_last_error()


def test_host_only_entry_points_work_without_a_gpu():
    import toybox_b200
    sch = toybox_b200.schema_for_state("amidar")
    assert "board" in sch["required"]
    assert "jitter" in toybox_b200.schema_for_config("space_invaders")["required"]
    out = C.c_void_p()
    assert toybox_b200.lib().tbx_pool_create(b"pong", 4, 0, None, C.byref(out)) != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "toybox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cuh", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle/|oracle\.py|tbo\.h", text, flags=re.M), (dirpath, f)
                assert "libtbo" not in text and "tbo_" not in text, (dirpath, f)
