import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the snapshot of the reference's own unit tests is a fixture that tests/test_reference_suite.py drives with a ctoybox module
# installed; pytest must not collect those files itself
collect_ignore_glob = ["golden/reference_py/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def tbx():
    """The product package with its CUDA library loaded; GPU tests fail loudly if it is not built."""
    import torch
    import toybox_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    toybox_b200.lib()
    return toybox_b200


def json_diff(a, b, path=""):
    """First few differences between two decoded JSON documents (numbers compare by value)."""
    num = (int, float)
    if isinstance(a, bool) != isinstance(b, bool) or (type(a) != type(b) and not (isinstance(a, num) and isinstance(b, num))):
        return [(path, a, b)]
    if isinstance(a, dict):
        out = []
        for k in sorted(set(a) | set(b)):
            if k not in a or k not in b:
                out.append((path + "." + k, "<missing>" if k not in a else "<present>", "<missing>" if k not in b else "<present>"))
            else:
                out += json_diff(a[k], b[k], path + "." + k)
        return out[:8]
    if isinstance(a, list):
        if len(a) != len(b):
            return [(path, "len %d" % len(a), "len %d" % len(b))]
        out = []
        for i, (x, y) in enumerate(zip(a, b)):
            out += json_diff(x, y, "%s[%d]" % (path, i))
            if len(out) > 8:
                break
        return out
    return [] if a == b else [(path, a, b)]
