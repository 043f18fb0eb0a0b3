"""First-principles pins of the PLDA oracle (oracle/kaldi_plda.py).

The reference's arithmetic lives in Kaldi, which is neither vendored nor buildable here, and the reference's own tests
hold no golden vectors ("parity unpinned").  What CAN be pinned without Kaldi is that the
restated formulas are the two-covariance model they claim to be -- checked here by brute force against plain
multivariate-normal densities (scipy), with nothing taken from the oracle but the function under test:

* ``Plda.log_likelihood_ratio`` (App. A.7) == log p(e_1..e_n, t | same speaker) - log p(e_1..e_n) - log p(t) under
  ``u ~ N(0, diag(psi)), x = u + eps, eps ~ N(0, I)`` -- exactly, for any n.
* ``PldaEstimator.compute_objf`` (App. A.4) == the exact data log-likelihood of the model ``x_si = mu + b_s + w_si``,
  ``b ~ N(0, B)``, ``w ~ N(0, W)`` up to a parameter-free constant (unit class weights).
* ``PldaEstimator.estimate_one_iter`` (App. A.3) is an EM step of that model: the exact log-likelihood never
  decreases (unit class weights, ragged speaker sizes).
"""
import numpy as np
import pytest
from scipy.stats import multivariate_normal as mvn

from oracle import kaldi_plda as kp


def _joint_same_speaker_cov(psi, n):
    """Covariance of (x_1, ..., x_n) sharing one speaker variable u ~ N(0, diag psi), x_i = u + N(0, I)."""
    d = psi.shape[0]
    return np.kron(np.ones((n, n)), np.diag(psi)) + np.eye(n * d)


@pytest.mark.parametrize("n", [1, 2, 5])
def test_llr_is_the_log_density_ratio_of_the_model(n):
    rng = np.random.RandomState(10 + n)
    d = 3
    psi = np.array([2.5, 0.7, 0.05])
    plda = kp.Plda()
    plda.mean = np.zeros(d)
    plda.transform = np.eye(d)
    plda.psi = psi
    plda.compute_derived_vars()
    for _ in range(5):
        enrol = rng.randn(n, d) * 1.3
        test = rng.randn(d) * 1.3
        same = mvn(mean=np.zeros((n + 1) * d), cov=_joint_same_speaker_cov(psi, n + 1)).logpdf(
            np.concatenate([enrol.reshape(-1), test]))
        enrol_alone = mvn(mean=np.zeros(n * d), cov=_joint_same_speaker_cov(psi, n)).logpdf(enrol.reshape(-1))
        test_alone = mvn(mean=np.zeros(d), cov=np.diag(psi + 1.0)).logpdf(test)
        want = same - enrol_alone - test_alone
        got = plda.log_likelihood_ratio(enrol.mean(axis=0), n, test)
        assert abs(got - want) <= 1e-10 * max(1.0, abs(want))


def _exact_loglik(groups, mu, within, between):
    """sum_s log N(vec(X_s); 1 (x) mu, I_n (x) W + 1 1^T (x) B): the model's marginal likelihood, by brute force."""
    tot = 0.0
    for g in groups:
        n, d = g.shape
        cov = np.kron(np.eye(n), within) + np.kron(np.ones((n, n)), between)
        tot += mvn(mean=np.tile(mu, n), cov=cov).logpdf(g.reshape(-1))
    return tot


def _problem(seed, d=3, sizes=(2, 3, 4, 2, 5, 3, 4)):
    rng = np.random.RandomState(seed)
    a = rng.randn(d, d)
    groups = []
    for n in sizes:
        spk = a @ rng.randn(d) * 1.5
        groups.append(0.3 + spk + rng.randn(n, d) @ np.diag([1.0, 0.6, 1.4]))
    return groups


def test_objf_is_the_exact_loglik_up_to_a_constant_and_em_ascends():
    groups = _problem(3)
    stats = kp.PldaStats()
    for g in groups:
        stats.add_samples(1.0, g)               # unit class weights: Kaldi's own ivector-compute-plda usage
    stats.sort()
    est = kp.PldaEstimator(stats)
    mu = stats.sum / stats.class_weight
    exact, objf = [], []
    for _ in range(8):
        est.estimate_one_iter()
        exact.append(_exact_loglik(groups, mu, est.within_var, est.between_var))
        objf.append(est.compute_objf() * stats.example_weight)
    exact, objf = np.array(exact), np.array(objf)
    # (i) Kaldi's objective differs from the exact log-likelihood by a parameter-free constant
    #     (the Jacobian of (x_1..x_n) -> (mean, deviations): (d/2) sum_s log n_s)
    gap = exact - objf
    assert np.max(np.abs(gap - gap[0])) <= 1e-8 * max(1.0, abs(gap[0]))
    d = groups[0].shape[1]
    assert abs(gap[0] - (-0.5 * d * sum(np.log(g.shape[0]) for g in groups))) <= 1e-8
    # (ii) every EstimateOneIter is an EM step of the model: the exact log-likelihood does not decrease
    assert np.all(np.diff(exact) >= -1e-9)
    assert exact[-1] > exact[0]


def test_reference_weights_are_a_weighted_likelihood_that_em_still_ascends():
    """The reference passes weight 1/n_s (src/pldamodule.cpp:97).  That is EM on a WEIGHTED log-likelihood: the
    restated objective (which carries the weights) must be non-decreasing, and with equal speaker sizes the weights
    are one common factor, so the exact log-likelihood must ascend too."""
    groups = _problem(5, sizes=(3, 3, 3, 3, 3, 3))
    stats = kp.PldaStats()
    for g in groups:
        stats.add_samples(1.0 / g.shape[0], g)
    stats.sort()
    est = kp.PldaEstimator(stats)
    mu = stats.sum / stats.class_weight
    exact, objf = [], []
    for _ in range(6):
        est.estimate_one_iter()
        exact.append(_exact_loglik(groups, mu, est.within_var, est.between_var))
        objf.append(est.compute_objf())
    assert np.all(np.diff(objf) >= -1e-10)
    assert np.all(np.diff(exact) >= -1e-9)
