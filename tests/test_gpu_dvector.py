"""GPU parity of the d-vector pooling kernel (C ABI plda_dvector_pool, csrc/dvector.cu) against the reference's own
functions -- through the fixture they generated -- and against the CPU restatement on seeded inputs, including the
shapes the reference's per-utterance functions return, ragged / single-frame / long (chunked) utterances, fp32 input
and device-resident input.  Tolerance: fp64 arithmetic on both sides, different summation order -> 1e-12 relative
(fp32 input: the reference normalises in fp32, the kernel in fp64 -> 1e-5)."""
import os

import numpy as np
import pytest

from oracle import dvector_port as dp

pytestmark = pytest.mark.gpu


def rel(got, ref):
    return np.max(np.abs(np.asarray(got, dtype=np.float64) - ref) / np.maximum(1e-3, np.abs(ref)))


def test_fixture_from_reference(golden_dir):
    from plda_b200 import dvector as dv
    g = np.load(os.path.join(golden_dir, "dvector_pool.npz"))
    for (method, l2), fn in dp.METHODS.items():
        got = dv.pool_dvectors(g["frames"], g["offsets"], method, l2)
        assert got.dtype == np.float64 and got.shape == g[fn.__name__].shape
        assert rel(got, g[fn.__name__]) <= 1e-12, fn.__name__


def test_per_utterance_functions_keep_reference_shapes(golden_dir):
    from plda_b200 import dvector as dv
    g = np.load(os.path.join(golden_dir, "dvector_pool.npz"))
    utt = g["frames"][g["offsets"][3]:g["offsets"][4]]
    for fn in dp.METHODS.values():
        got = getattr(dv, fn.__name__)(utt)
        want = np.asarray(fn(utt))
        assert got.shape == want.shape, fn.__name__
        assert rel(got, want) <= 1e-12, fn.__name__


@pytest.mark.parametrize("d", [1, 31, 200, 512, 1024])
def test_ragged_and_long_utterances(d):
    """Utterances longer than the 2048-frame work item are chunked and merged; lengths 1 .. 5000."""
    from plda_b200 import dvector as dv
    rng = np.random.RandomState(d)
    lens = [1, 2, 31, 32, 33, 255, 2048, 2049, 5000]
    frames = rng.randn(sum(lens), d) * 3 + 1.5
    offsets = np.concatenate([[0], np.cumsum(lens)])
    for method in ("mean", "max", "var"):
        for l2 in (True, False):
            got = dv.pool_dvectors(frames, offsets, method, l2)
            assert rel(got, dp.pool(frames, offsets, method, l2)) <= 1e-11, (method, l2)


def test_float32_and_device_resident_input():
    import torch
    from plda_b200 import dvector as dv
    rng = np.random.RandomState(9)
    lens = [40, 7, 300]
    frames = (rng.randn(sum(lens), 64) + 0.2).astype(np.float32)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    want = dp.pool(frames.astype(np.float64), offsets, "mean", True)
    got_host = dv.pool_dvectors(frames, offsets, "mean", True)
    assert rel(got_host, want) <= 1e-12                       # fp32 values, fp64 arithmetic
    assert rel(got_host, dp.pool(frames, offsets, "mean", True)) <= 1e-5   # the reference's fp32 arithmetic
    got_dev = dv.pool_dvectors(torch.from_numpy(frames).cuda(), offsets, "mean", True)
    assert got_dev.is_cuda and np.array_equal(got_dev.cpu().numpy(), got_host)
    # a view with a row pitch (every other column block of a wider matrix)
    wide = torch.from_numpy(np.concatenate([frames, frames], axis=1)).cuda()
    assert np.array_equal(dv.pool_dvectors(wide[:, :64], offsets, "mean", True).cpu().numpy(), got_host)


def test_extractvectors_and_errors():
    from plda_b200 import dvector as dv
    rng = np.random.RandomState(2)
    data = {"spkA": [rng.randn(20, 16), rng.randn(5, 16)], "spkB": [rng.randn(9, 16)]}
    vecs, labels = dv.extractvectors(data, "var")
    assert vecs.shape == (3, 16) and list(labels) == ["spkA", "spkA", "spkB"]
    assert rel(vecs[1], dp.extractdvectorvar(data["spkA"][1])) <= 1e-12
    with pytest.raises(ValueError):
        dv.pool_dvectors(rng.randn(10, 4), [0, 5, 5, 10])      # an empty utterance
    with pytest.raises(ValueError):
        dv.pool_dvectors(rng.randn(10, 4), [0, 11])            # offsets past the end
    with pytest.raises(ValueError):
        dv.pool_dvectors(rng.randn(10, 4), [0, 10], method="median")
    with pytest.raises(ValueError):
        dv.pool_dvectors(np.arange(12).reshape(3, 4), [0, 3])  # not floats (as_matrix, like PLDA.fit)
