"""Pins for the PLDA oracle (oracle/kaldi_plda.py).

The reference has no golden vectors for this path and Kaldi cannot be built here
("parity unpinned"), so the oracle is pinned by the self-consistency invariants of
SURVEY.md section 8c plus a committed regression fixture.
"""
import os

import numpy as np
import pytest

from oracle import kaldi_plda as kp


@pytest.fixture(scope="module")
def fitted():
    d = 16
    a_b = kp.two_cov_generator(d, seed=1234)
    rng = np.random.RandomState(3)
    counts = rng.randint(2, 9, size=40)
    x, labels, _ = kp.synth_speakers(a_b, counts, seed=1234)
    m = kp.MPlda()
    objf = []
    assert m.fit(x, labels, 8, objf_log=objf) is None
    return m, x, labels, objf, a_b


def test_joint_diagonalisation(fitted):
    m, *_ = fitted
    est, p = m.estimator, m.plda
    a = p.transform
    d = a.shape[0]
    assert np.allclose(a @ est.within_var @ a.T, np.eye(d), atol=1e-10)
    assert np.allclose(a @ est.between_var @ a.T, np.diag(p.psi), atol=1e-10)
    assert np.all(np.diff(p.psi) <= 1e-15) and p.psi.min() >= 0.0
    assert np.allclose(p.offset, -a @ p.mean, atol=1e-13)


def test_em_objective_non_decreasing(fitted):
    _, _, _, objf, _ = fitted
    assert all(b >= a - 1e-12 for a, b in zip(objf, objf[1:]))


def test_counts_bookkeeping(fitted):
    m, x, labels, *_ = fitted
    est = m.estimator
    k = len(np.unique(labels))
    # App. A.3: W_count ends at exactly K, B_count = sum 1/n_s (reference weights 1/n_s)
    assert est.within_var_count == pytest.approx(k, rel=1e-13)
    n = np.bincount(labels.astype(np.int64))
    assert est.between_var_count == pytest.approx(np.sum(1.0 / n), rel=1e-13)


def test_length_normalisation(fitted):
    m, x, labels, *_ = fitted
    out = m.transform(x[:50], labels[:50])
    d = x.shape[1]
    for k, (n, y) in out.items():
        assert np.sum(y * y / (m.plda.psi + 1.0 / n)) == pytest.approx(d, rel=1e-12)


def test_llr_symmetry_n1(fitted):
    m, x, *_ = fitted
    p = m.plda
    a = p.transform_ivector(x[0], 1)
    b = p.transform_ivector(x[1], 1)
    assert p.log_likelihood_ratio(a, 1, b) == pytest.approx(p.log_likelihood_ratio(b, 1, a), abs=1e-12)


def test_grid_equals_per_pair(fitted):
    m, x, labels, _, a_b = fitted
    xe, le, _ = kp.synth_speakers(a_b, [1, 2, 3, 3, 5, 2], seed=5)
    xt, lt, _ = kp.synth_speakers(a_b, [1] * 9, seed=6)
    te, tt = m.transform(xe, le), m.transform(xt, lt)
    ek = sorted(te)
    e = np.stack([te[k][1] for k in ek])
    n = np.array([te[k][0] for k in ek])
    t = np.stack([tt[k][1] for k in sorted(tt)])
    grid = kp.score_grid(m.plda, e, n, t)
    for i, k in enumerate(ek):
        for j in range(t.shape[0]):
            assert grid[i, j] == pytest.approx(m.plda.log_likelihood_ratio(e[i], int(n[i]), t[j]), abs=1e-11)
    # batched transform == per-call transform
    u, c, mu = kp.group_means(xe, le)
    tb = kp.transform_batch(m.plda, mu, c)
    assert np.allclose(tb, e, atol=1e-12)


def test_sign_flip_invariance(fitted):
    m, x, labels, _, a_b = fitted
    xe, le, _ = kp.synth_speakers(a_b, [2, 3], seed=8)
    xt, lt, _ = kp.synth_speakers(a_b, [1, 1, 1], seed=9)
    p2 = m.plda.copy()
    flips = np.where(np.arange(p2.dim()) % 2 == 0, -1.0, 1.0)
    p2.transform = p2.transform * flips[:, None]
    p2.compute_derived_vars()
    for p in (m.plda, p2):
        u, c, mu = kp.group_means(xe, le)
        e = kp.transform_batch(p, mu, c)
        _, ct, mt = kp.group_means(xt, lt)
        t = kp.transform_batch(p, mt, ct)
        s = kp.score_grid(p, e, c, t)
        if p is m.plda:
            ref = s
    assert np.allclose(s, ref, atol=1e-11)


def test_diagonalised_em_equals_kaldi_loop(fitted):
    m, x, labels, *_ = fitted
    lab = labels.astype(np.int64)
    stats = kp.PldaStats()
    for s in range(lab.max() + 1):
        g = x[lab == s]
        stats.add_samples(1.0 / g.shape[0], g)
    stats.sort()
    est = kp.PldaEstimator(stats)
    means = np.stack([c.mean for c in stats.class_info])
    counts = np.array([c.num_examples for c in stats.class_info])
    weights = np.array([c.weight for c in stats.class_info])
    mu = stats.sum / stats.class_weight
    w, b = np.eye(x.shape[1]), np.eye(x.shape[1])
    for _ in range(4):
        est.estimate_one_iter()
        ws, bs = kp.em_iter_diag(stats.offset_scatter, means, counts, weights, mu, w, b)
        w = ws / float(len(counts))
        b = bs / weights.sum()
        assert np.allclose(w, est.within_var, rtol=1e-11, atol=1e-13)
        assert np.allclose(b, est.between_var, rtol=1e-11, atol=1e-13)


def test_znorm_vectorised_equals_shim(fitted):
    m, x, labels, _, a_b = fitted
    import copy
    m2 = kp.MPlda()
    m2.plda = m.plda.copy()
    xe, le, _ = kp.synth_speakers(a_b, [2, 3, 1, 4], seed=15)
    te = m2.transform(xe, le)
    bkg, _, _ = kp.synth_speakers(a_b, [1] * 12, seed=16)
    m2.norm(bkg, te)
    e = np.stack([te[k][1] for k in sorted(te)])
    mean, std = kp.znorm_stats(m2.plda, bkg, e)
    assert np.allclose(mean, [m2.meanz[k] for k in sorted(te)], atol=1e-12)
    assert np.allclose(std, [m2.stdvz[k] for k in sorted(te)], atol=1e-12)
    # second norm() never overwrites (std::unordered_map::insert, src/pldamodule.cpp:245,250)
    before = dict(m2.meanz)
    m2.norm(bkg[:5], te)
    assert before == m2.meanz


def test_recovers_discriminative_structure():
    d = 20
    a_b = kp.two_cov_generator(d, seed=1234)
    x, labels, _ = kp.synth_speakers(a_b, [10] * 60, seed=1234)
    m = kp.MPlda()
    m.fit(x, labels, 10)
    xe, le, _ = kp.synth_speakers(a_b, [3] * 30, seed=1235)
    # tests: one utterance of each enrol speaker (same z) -> regenerate with shared z
    rng = np.random.RandomState(1236)
    _, _, z = kp.synth_speakers(a_b, [3] * 30, seed=1235)
    xt = 0.5 + z @ a_b.T + rng.randn(30, d)
    te = m.transform(xe, le)
    tt = m.transform(xt, np.arange(30, dtype=np.uint64))
    e = np.stack([te[k][1] for k in sorted(te)])
    t = np.stack([tt[k][1] for k in sorted(tt)])
    s = kp.score_grid(m.plda, e, np.full(30, 3), t)
    tar = np.diag(s)
    non = s[~np.eye(30, dtype=bool)]
    assert tar.mean() > non.mean() + 1.0
    assert kp.eer_percent(tar, non) < 25.0


def test_reference_error_behaviour():
    m = kp.MPlda()
    x = np.random.RandomState(0).rand(20, 4)
    with pytest.raises(ValueError):
        m.fit(x, np.zeros(20, dtype=np.int64))            # not unsigned (src/pldamodule.cpp:55-58)
    with pytest.raises(ValueError):
        m.fit(x.astype(np.int32), np.zeros(20, dtype=np.uint64))   # not float (:59-62)
    with pytest.raises(ValueError):
        m.fit(x, np.zeros(20, dtype=np.uint64))           # one speaker (:83-86)
    m.fit(x, (np.arange(20) % 2).astype(np.uint64), 2)
    with pytest.raises(ValueError):
        m.transform(x, np.array(["a"] * 20))              # strings (:128-131)


def test_score_is_float32_rounded(fitted):
    m, x, labels, *_ = fitted
    out = m.transform(x[:6], np.arange(6, dtype=np.uint64))
    s = m.score(0, out[0], out[1])
    assert s == float(np.float32(s))


def test_regression_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "plda_small.npz"))
    m = kp.MPlda()
    m.fit(g["x"], g["labels"], int(g["iters"]))
    assert np.allclose(m.plda.psi, g["psi"], rtol=1e-9, atol=1e-12)
    assert np.allclose(m.plda.mean, g["mean"], rtol=1e-12)
    te = m.transform(g["xe"], g["le"])
    tt = m.transform(g["xt"], g["lt"])
    e = np.stack([te[k][1] for k in sorted(te)])
    t = np.stack([tt[k][1] for k in sorted(tt)])
    s = kp.score_grid(m.plda, e, g["enrol_counts"], t)
    assert np.allclose(s.astype(np.float32), g["scores"], rtol=1e-5, atol=1e-5)
    mean, std = kp.znorm_stats(m.plda, g["bkg"], e)
    assert np.allclose(mean, g["meanz"], atol=1e-9)
    assert np.allclose(std, g["stdvz"], atol=1e-9)
    assert np.allclose(((s - mean[:, None]) / std[:, None]).astype(np.float32), g["zscores"], rtol=1e-5, atol=1e-5)


def test_c1_shape_runs():
    """BASELINE configs[0]: 500x200 rand, 2 speakers (README.md:52-57)."""
    rng = np.random.RandomState(0)
    x = rng.rand(500, 200)
    y = rng.randint(0, 2, 500).astype("uint")
    m = kp.MPlda()
    assert m.fit(x, y, 10) is None
    out = m.transform(x[:10], np.arange(10, dtype="uint"))
    s = m.score(0, out[0], out[1])
    assert -100 <= s <= 100
