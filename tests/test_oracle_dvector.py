"""Pins oracle/dvector_port.py against the REFERENCE's own pooling functions (scoring/extractdvector.py:19-58):
through the committed fixture they generated (tests/golden/make_golden.py: make_dvector) and live when the reference
tree exists (authoring container)."""
import os

import numpy as np
import pytest

from oracle import dvector_port as dp

REF = "/root/reference/scoring/extractdvector.py"


def test_port_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "dvector_pool.npz"))
    for (method, l2), fn in dp.METHODS.items():
        got = dp.pool(g["frames"], g["offsets"], method, l2)
        assert np.array_equal(got, g[fn.__name__]), fn.__name__
    # the *_nol2 functions return shape (1, d) (extractdvector.py:50,54,58)
    assert tuple(g["shape_nol2"]) == dp.extractdvectormean_nol2(g["frames"][:5]).shape


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the authoring container")
def test_port_matches_reference_live():
    with open(REF) as f:
        lines = f.readlines()
    ns = {"np": np}
    exec(compile("".join(lines[18:58]), REF, "exec"), ns)
    rng = np.random.RandomState(5)
    for dtype in (np.float64, np.float32):
        utt = (rng.randn(57, 24) * 2 + 0.3).astype(dtype)
        for fn in dp.METHODS.values():
            assert np.array_equal(np.asarray(fn(utt)), np.asarray(ns[fn.__name__](utt))), fn.__name__
