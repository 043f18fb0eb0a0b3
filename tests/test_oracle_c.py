"""Pins the C restatement (oracle/plda_ref.c, timed as the CPU baseline) against the numpy oracle."""
import numpy as np

from oracle import c_ref
from oracle import kaldi_plda as kp


def test_c_port_matches_numpy_oracle():
    d = 24
    a_b = kp.two_cov_generator(d, seed=1234)
    rng = np.random.RandomState(4)
    x, labels, _ = kp.synth_speakers(a_b, rng.randint(2, 9, size=50), seed=1234)
    ref = kp.MPlda()
    ref.fit(x, labels, 5)
    c = c_ref.RefPlda(x, labels)
    for _ in range(5):
        c.em_iter()
    w, b = c.covariances()
    assert np.allclose(w, ref.estimator.within_var, rtol=1e-9, atol=1e-12)
    assert np.allclose(b, ref.estimator.between_var, rtol=1e-9, atol=1e-12)
    mean, tr, psi = c.output()
    assert np.allclose(mean, ref.plda.mean, rtol=1e-12)
    assert np.allclose(psi, ref.plda.psi, rtol=1e-8, atol=1e-12)
    # scores are invariant to the eigenvector signs
    p = kp.Plda()
    p.mean, p.transform, p.psi = mean, tr, psi
    p.compute_derived_vars()
    e = np.stack([c_ref.transform(mean, tr, psi, x[i], 2) for i in range(6)])
    e_np = np.stack([p.transform_ivector(x[i], 2) for i in range(6)])
    assert np.allclose(e, e_np, atol=1e-11)
    t = np.stack([c_ref.transform(mean, tr, psi, x[10 + i], 1) for i in range(7)])
    e_ref = np.stack([ref.plda.transform_ivector(x[i], 2) for i in range(6)])
    t_ref = np.stack([ref.plda.transform_ivector(x[10 + i], 1) for i in range(7)])
    grid, used = c_ref.score_grid(psi, e, np.full(6, 2), t, threads=2)
    want = kp.score_grid(ref.plda, e_ref, np.full(6, 2), t_ref)
    assert used >= 1
    assert np.allclose(grid, want, rtol=1e-5, atol=1e-5)
