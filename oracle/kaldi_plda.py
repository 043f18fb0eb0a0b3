"""CPU fp64 ORACLE for the PLDA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``plda_b200``) never imports anything under ``oracle/``.

PARITY UNPINNED: the arithmetic of the reference lives in Kaldi's
``src/ivector/plda.{h,cc}`` (included at ``src/pldamodule.cpp:16``), a third-party
dependency that is neither vendored nor version-pinned by the reference
(``README.md:6-13``, ``cmake/FindKaldi.cmake:14-68``) and cannot be built in this
image (needs Kaldi + ATLAS + CPython 2).  The reference's own tests hold no golden
vectors for this path (``tests/pldatest.py:33,53,82`` only assert
``-100 <= score <= 100``).  This file therefore restates Kaldi's *published*
algorithm (``PldaStats``, ``PldaEstimator``, ``Plda`` in ivector/plda.cc, API with
the 4-argument ``TransformIvector`` used at ``src/pldamodule.cpp:171,224``) in the
operation order Kaldi uses, and follows ``src/pldamodule.cpp`` line by line for
everything the shim adds.  It is pinned only by the self-consistency invariants in
``tests/test_oracle_plda.py`` and by the first-principles checks in
``tests/test_oracle_first_principles.py`` (LLR, EM objective and EM step against brute-force
multivariate-normal densities of the two-covariance model) -- not by any output of Kaldi.

Two layers:

* ``PldaStats`` / ``PldaEstimator`` / ``Plda`` -- Kaldi's classes (call sites:
  ``src/pldamodule.cpp:70,97,100,103,106,159,171,224,235,266``).
* ``MPlda`` -- the reference's CPython type (``src/pldamodule.cpp:27-34``) with
  ``fit`` (:42-109), ``transform`` (:111-194), ``norm`` (:196-256), ``score``
  (:258-277).

Plus vectorised restatements (``score_grid``, ``transform_batch``) that the
tests prove equal to the per-call forms to ~1e-13, used where the per-pair loop
would be too slow for a test.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

M_LOG_2PI = 1.8378770664093454835606594728112


# --------------------------------------------------------------------------- #
# Kaldi ivector/plda.cc  (not in tree; restated)
# --------------------------------------------------------------------------- #
@dataclass
class ClassInfo:
    weight: float
    mean: np.ndarray
    num_examples: int


class PldaStats:
    """Kaldi ``PldaStats`` (constructed at ``src/pldamodule.cpp:70``)."""

    def __init__(self) -> None:
        self.dim = 0
        self.num_classes = 0
        self.num_examples = 0
        self.class_weight = 0.0
        self.example_weight = 0.0
        self.sum: Optional[np.ndarray] = None
        self.offset_scatter: Optional[np.ndarray] = None
        self.class_info: List[ClassInfo] = []

    def add_samples(self, weight: float, group: np.ndarray) -> None:
        """``PldaStats::AddSamples`` (called at ``src/pldamodule.cpp:97``)."""
        group = np.asarray(group, dtype=np.float64)
        if self.dim == 0:
            self.dim = group.shape[1]
            self.sum = np.zeros(self.dim)
            self.offset_scatter = np.zeros((self.dim, self.dim))
        n = group.shape[0]
        mean = group.sum(axis=0) * (1.0 / n)                 # AddRowSumMat(1/n)
        self.offset_scatter += weight * (group.T @ group)    # AddMat2(weight, group, kTrans)
        self.offset_scatter += (-n * weight) * np.outer(mean, mean)   # AddVec2(-n*weight, mean)
        self.class_info.append(ClassInfo(weight, mean, n))
        self.num_classes += 1
        self.num_examples += n
        self.class_weight += weight
        self.example_weight += weight * n
        self.sum += weight * mean

    def sort(self) -> None:
        """``PldaStats::Sort`` (``src/pldamodule.cpp:100``): ascending num_examples."""
        self.class_info.sort(key=lambda c: c.num_examples)

    def is_sorted(self) -> bool:
        ns = [c.num_examples for c in self.class_info]
        return all(a <= b for a, b in zip(ns, ns[1:]))


class Plda:
    """Kaldi ``Plda`` (member of ``MPlda``, ``src/pldamodule.cpp:29``)."""

    def __init__(self) -> None:
        self.mean: Optional[np.ndarray] = None
        self.transform: Optional[np.ndarray] = None
        self.psi: Optional[np.ndarray] = None
        self.offset: Optional[np.ndarray] = None

    def dim(self) -> int:
        return 0 if self.mean is None else self.mean.shape[0]

    def compute_derived_vars(self) -> None:
        self.offset = -1.0 * (self.transform @ self.mean)

    def get_normalization_factor(self, transformed: np.ndarray, num_examples: int) -> float:
        assert num_examples > 0
        sq = transformed ** 2
        inv_covar = 1.0 / (self.psi + 1.0 / num_examples)
        dot_prod = float(np.dot(inv_covar, sq))
        return math.sqrt(self.dim() / dot_prod)

    def transform_ivector(self, ivector: np.ndarray, num_examples: int,
                          normalize_length: bool = True,
                          simple_length_norm: bool = False) -> np.ndarray:
        """``Plda::TransformIvector`` with default ``PldaConfig``
        (``src/pldamodule.cpp:171,224``; the config is never changed, :31)."""
        y = self.offset + self.transform @ np.asarray(ivector, dtype=np.float64)
        if simple_length_norm:
            f = math.sqrt(y.shape[0]) / float(np.linalg.norm(y))
        else:
            f = self.get_normalization_factor(y, num_examples)
        if normalize_length:
            y = y * f
        return y

    def log_likelihood_ratio(self, train: np.ndarray, n: int, test: np.ndarray) -> float:
        """``Plda::LogLikelihoodRatio`` (``src/pldamodule.cpp:235,266``)."""
        psi = self.psi
        dim = self.dim()
        mean = n * psi / (n * psi + 1.0) * train
        variance = 1.0 + psi / (n * psi + 1.0)
        logdet = float(np.sum(np.log(variance)))
        sqdiff = (test - mean) ** 2
        loglike_given_class = -0.5 * (logdet + M_LOG_2PI * dim + float(np.dot(sqdiff, 1.0 / variance)))
        sqdiff = test ** 2
        variance = psi + 1.0
        logdet = float(np.sum(np.log(variance)))
        loglike_without_class = -0.5 * (logdet + M_LOG_2PI * dim + float(np.dot(sqdiff, 1.0 / variance)))
        return loglike_given_class - loglike_without_class

    def smooth_within_class_covariance(self, smoothing_factor: float) -> None:
        """``Plda::SmoothWithinClassCovariance`` (``src/pldamodule.cpp:159``)."""
        assert 0.0 <= smoothing_factor <= 1.0
        within = 1.0 + smoothing_factor * self.psi
        self.psi = self.psi / within
        self.transform = self.transform * (within ** -0.5)[:, None]   # MulRowsVec
        self.compute_derived_vars()

    def copy(self) -> "Plda":
        p = Plda()
        p.mean, p.transform, p.psi, p.offset = (self.mean.copy(), self.transform.copy(),
                                                 self.psi.copy(), self.offset.copy())
        return p


def _sp_invert(a: np.ndarray) -> np.ndarray:
    """``SpMatrix::Invert`` (LU under HAVE_ATLAS, ``CMakeLists.txt:90``); result
    re-symmetrised because Kaldi copies the lower triangle back into packed form."""
    inv = np.linalg.inv(a)
    low = np.tril(inv)
    return low + np.tril(inv, -1).T


class PldaEstimator:
    """Kaldi ``PldaEstimator`` (``src/pldamodule.cpp:103-106``)."""

    def __init__(self, stats: PldaStats) -> None:
        assert stats.is_sorted()
        self.stats = stats
        d = stats.dim
        self.within_var = np.eye(d)
        self.between_var = np.eye(d)
        self.within_var_stats = np.zeros((d, d))
        self.within_var_count = 0.0
        self.between_var_stats = np.zeros((d, d))
        self.between_var_count = 0.0

    # -- objective (ComputeObjf*; used only as an invariant: non-decreasing) -- #
    def compute_objf_part1(self) -> float:
        d = self.stats.dim
        within_class_count = self.stats.example_weight - self.stats.class_weight
        sign, within_logdet = np.linalg.slogdet(self.within_var)
        inv_within = _sp_invert(self.within_var)
        return -0.5 * (within_class_count * (within_logdet + M_LOG_2PI * d)
                       + float(np.sum(inv_within * self.stats.offset_scatter)))

    def compute_objf_part2(self) -> float:
        d = self.stats.dim
        tot = 0.0
        n = -1
        combined_inv = None
        combined_logdet = 0.0
        mu = self.stats.sum / self.stats.class_weight
        for info in self.stats.class_info:
            if info.num_examples != n:
                n = info.num_examples
                combined = self.between_var + self.within_var / n
                _, combined_logdet = np.linalg.slogdet(combined)
                combined_inv = _sp_invert(combined)
            m = info.mean - mu
            tot += info.weight * -0.5 * (combined_logdet + M_LOG_2PI * d + float(m @ combined_inv @ m))
        return tot

    def compute_objf(self) -> float:
        return (self.compute_objf_part1() + self.compute_objf_part2()) / self.stats.example_weight

    # -- EM -- #
    def reset_per_iter_stats(self) -> None:
        d = self.stats.dim
        self.within_var_stats = np.zeros((d, d))
        self.within_var_count = 0.0
        self.between_var_stats = np.zeros((d, d))
        self.between_var_count = 0.0

    def get_stats_from_intra_class(self) -> None:
        self.within_var_stats += self.stats.offset_scatter
        self.within_var_count += (self.stats.example_weight - self.stats.class_weight)

    def get_stats_from_class_means(self) -> None:
        between_var_inv = _sp_invert(self.between_var)
        within_var_inv = _sp_invert(self.within_var)
        mixed_var = None
        n = -1
        mu = self.stats.sum * (1.0 / self.stats.class_weight)
        for info in self.stats.class_info:
            weight = info.weight
            if info.num_examples != n:
                n = info.num_examples
                mixed_var = _sp_invert(between_var_inv + n * within_var_inv)
            m = info.mean - mu
            temp = n * (within_var_inv @ m)
            w = mixed_var @ temp
            m_w = m - w
            self.between_var_stats += weight * mixed_var
            self.between_var_stats += weight * np.outer(w, w)
            self.between_var_count += weight
            self.within_var_stats += (weight * n) * mixed_var
            self.within_var_stats += (weight * n) * np.outer(m_w, m_w)
            self.within_var_count += weight

    def estimate_from_stats(self) -> None:
        self.within_var = self.within_var_stats * (1.0 / self.within_var_count)
        self.between_var = self.between_var_stats * (1.0 / self.between_var_count)

    def estimate_one_iter(self) -> None:
        self.reset_per_iter_stats()
        self.get_stats_from_intra_class()
        self.get_stats_from_class_means()
        self.estimate_from_stats()

    def get_output(self) -> Plda:
        plda = Plda()
        plda.mean = self.stats.sum * (1.0 / self.stats.class_weight)
        # ComputeNormalizingTransform: C = chol(W) (lower), transform1 = C^-1
        c = np.linalg.cholesky(self.within_var)
        transform1 = np.linalg.inv(c)
        transform1 = np.tril(transform1)
        between_var_proj = transform1 @ self.between_var @ transform1.T
        between_var_proj = 0.5 * (between_var_proj + between_var_proj.T)
        s, u = np.linalg.eigh(between_var_proj)
        assert s.min() >= -1e-10 * max(1.0, abs(s.max())), "between-class eigenvalue negative"
        s = np.maximum(s, 0.0)                       # ApplyFloor(0.0)
        order = np.argsort(-np.abs(s), kind="stable")   # SortSvd: descending
        s, u = s[order], u[:, order]
        plda.transform = u.T @ transform1
        plda.psi = s
        plda.compute_derived_vars()
        return plda

    def estimate(self, num_em_iters: int = 10, objf_log: Optional[list] = None) -> Plda:
        for _ in range(num_em_iters):
            self.estimate_one_iter()
            if objf_log is not None:
                objf_log.append(self.compute_objf())
        return self.get_output()


# --------------------------------------------------------------------------- #
# src/pldamodule.cpp  (the reference's CPython shim)
# --------------------------------------------------------------------------- #
def _check_labels(labels: np.ndarray) -> np.ndarray:
    labels = np.asarray(labels)
    if labels.dtype.kind == "S" or labels.dtype.kind == "U":
        raise ValueError("Labels need to be numpy array of uints, not strings!")   # :128-131
    if labels.dtype.kind != "u":
        raise ValueError("Given labels (argument 2) are not an unsigned! Set the dtype to uint!")  # :55-58
    return labels.astype(np.int64)


class MPlda:
    """The reference's ``libplda.MPlda`` object (``src/pldamodule.cpp:27-34``)."""

    def __init__(self) -> None:
        self.plda = Plda()
        self.meanz: Dict[int, float] = {}
        self.stdvz: Dict[int, float] = {}
        self.estimator: Optional[PldaEstimator] = None   # kept for invariants only

    # :42-109
    def fit(self, features: np.ndarray, labels: np.ndarray, iters: int = 10,
            objf_log: Optional[list] = None) -> None:
        lab = _check_labels(labels)
        features = np.asarray(features)
        if features.dtype.kind != "f":
            raise ValueError("Given Input features (argument 1) are not floats! Set the dtype to float!")
        feats = np.ascontiguousarray(features, dtype=np.float64)
        assert lab.shape[0] == feats.shape[0]
        u_labels = np.unique(lab)
        num_speakers = u_labels.shape[0]
        if num_speakers == 1:
            raise ValueError("Number of speakers is 1. Aborting PLDA esimation, at least two speakers are required!")
        # :88-92 labels index an array of size num_speakers => must be dense 0..K-1
        assert lab.max() < num_speakers, "reference requires dense labels 0..K-1 (src/pldamodule.cpp:88-92)"
        stats = PldaStats()
        order = np.argsort(lab, kind="stable")
        bounds = np.searchsorted(lab[order], np.arange(num_speakers + 1))
        for spk in range(num_speakers):
            idx = order[bounds[spk]:bounds[spk + 1]]
            tmp = feats[idx]
            stats.add_samples(1.0 / idx.shape[0], tmp)         # :97  weight = 1/n_s
        stats.sort()                                           # :100
        est = PldaEstimator(stats)
        self.plda = est.estimate(iters, objf_log)              # :102-106
        self.estimator = est
        return None

    # :111-194
    def transform(self, features: np.ndarray, labels: np.ndarray,
                  targetdim: int = 0, smoothfactor: float = 1.0) -> Dict[int, Tuple[int, np.ndarray]]:
        lab = _check_labels(labels)
        feats = np.ascontiguousarray(features, dtype=np.float64)
        assert targetdim == 0, "reference targetdim plumbing asserts inside Kaldi (SURVEY App. B)"
        sums: Dict[int, np.ndarray] = {}
        sizes: Dict[int, int] = {}
        for i in range(feats.shape[0]):                       # :147-156
            k = int(lab[i])
            if k not in sums:
                sums[k] = np.zeros(feats.shape[1])
                sizes[k] = 0
            sizes[k] += 1
            sums[k] += feats[i]
        if smoothfactor != 1.0:                                # :158-160 (mutates model)
            self.plda.smooth_within_class_covariance(smoothfactor)
        out: Dict[int, Tuple[int, np.ndarray]] = {}
        for k in sorted(sums):                                 # std::map order, :164
            n = sizes[k]
            mean = sums[k] * (1.0 / n)
            out[k] = (n, self.plda.transform_ivector(mean, n))
        return out

    # :196-256
    def norm(self, bkgdata: np.ndarray, spktoutt: Dict[int, Tuple[int, np.ndarray]],
             numutts: int = 0, rows: Optional[Sequence[int]] = None) -> None:
        bkg = np.ascontiguousarray(bkgdata, dtype=np.float64)
        m = bkg.shape[0]
        if numutts == 0:
            numutts = m
        if rows is None:
            # reference: unseeded std::random_shuffle then first numutts (:204-213);
            # with numutts == all rows the subset is the full set.
            rows = list(range(m))[:numutts]
        scores: Dict[int, List[float]] = {}
        for r in rows:
            t = self.plda.transform_ivector(bkg[r], m)         # :224  num_examples = #bkg rows (sic)
            for k, (_n, vec) in spktoutt.items():
                s = self.plda.log_likelihood_ratio(t, 1, np.asarray(vec, dtype=np.float64))   # :235
                scores.setdefault(int(k), []).append(s)
        for k, v in scores.items():                            # :240-253
            a = np.asarray(v)
            mean = float(a.sum() / a.shape[0])
            sq = float(np.sum((a - mean) ** 2) / a.shape[0])
            if k not in self.meanz:                            # insert() never overwrites
                self.meanz[k] = mean
            if k not in self.stdvz:
                self.stdvz[k] = math.sqrt(sq)
        return None

    # :258-277
    def score(self, enrolemodelid: int, enrolemodel: Tuple[int, np.ndarray],
              testutt: Tuple[int, np.ndarray]) -> float:
        n = int(enrolemodel[0])
        e = np.asarray(enrolemodel[1], dtype=np.float64)
        t = np.asarray(testutt[1], dtype=np.float64)
        s = self.plda.log_likelihood_ratio(e, n, t)
        if len(self.meanz) and enrolemodelid in self.meanz:
            s = (s - self.meanz[enrolemodelid]) / self.stdvz[enrolemodelid]
        return float(np.float32(s))                            # Py_BuildValue("f") :276


# --------------------------------------------------------------------------- #
# Vectorised restatements (proved equal to the per-call forms in tests)
# --------------------------------------------------------------------------- #
def transform_batch(plda: Plda, means: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Row-wise ``TransformIvector``: means (R,d), counts (R,) -> (R,d)."""
    y = means @ plda.transform.T + plda.offset
    inv_covar = 1.0 / (plda.psi[None, :] + 1.0 / counts[:, None].astype(np.float64))
    f = np.sqrt(plda.dim() / np.sum(inv_covar * y * y, axis=1))
    return y * f[:, None]


def group_means(features: np.ndarray, labels: np.ndarray):
    """Per-label mean in ascending label order -> (uniq_labels, counts, means)."""
    lab = np.asarray(labels).astype(np.int64)
    uniq, inv, cnt = np.unique(lab, return_inverse=True, return_counts=True)
    sums = np.zeros((uniq.shape[0], features.shape[1]))
    np.add.at(sums, inv, np.asarray(features, dtype=np.float64))
    return uniq, cnt, sums / cnt[:, None]


def score_grid(plda: Plda, enrol: np.ndarray, enrol_counts: np.ndarray, test: np.ndarray) -> np.ndarray:
    """All-pairs ``LogLikelihoodRatio`` -> (Ne, Nt) fp64 (Gram form)."""
    psi = plda.psi
    ne = enrol.shape[0]
    out = np.empty((ne, test.shape[0]))
    t2 = test * test
    base = 0.5 * np.sum(np.log1p(psi))
    for n in np.unique(enrol_counts):
        sel = np.nonzero(enrol_counts == n)[0]
        a = n * psi / (n * psi + 1.0)
        v = 1.0 + psi / (n * psi + 1.0)
        c = -0.5 * np.sum(np.log(v)) + base
        q = 0.5 * (1.0 / (1.0 + psi) - 1.0 / v)
        e = enrol[sel]
        row = c - 0.5 * np.sum((a * a / v) * e * e, axis=1)
        col = t2 @ q
        out[sel] = (e * (a / v)) @ test.T + row[:, None] + col[None, :]
    return out


def znorm_stats(plda: Plda, bkg: np.ndarray, enrol: np.ndarray, rows=None):
    """Vectorised ``MPlda.norm``: returns (mean, std) per enrol row."""
    m = bkg.shape[0]
    if rows is None:
        rows = np.arange(m)
    t = transform_batch(plda, bkg[rows], np.full(len(rows), m))
    # LLR(train=bkg_t, n=1, test=enrol)  (argument order of src/pldamodule.cpp:235)
    s = score_grid(plda, t, np.ones(len(rows), dtype=np.int64), enrol)   # (M, Ne)
    mean = s.sum(axis=0) / s.shape[0]
    std = np.sqrt(np.sum((s - mean[None, :]) ** 2, axis=0) / s.shape[0])
    return mean, std


# --------------------------------------------------------------------------- #
# Diagonalised EM (SURVEY App. A.3 identity) -- the form the CUDA path uses;
# kept here so tests can prove it equals Kaldi's per-class loop.
# --------------------------------------------------------------------------- #
def joint_diag(within: np.ndarray, between: np.ndarray):
    c = np.linalg.cholesky(within)
    t1 = np.tril(np.linalg.inv(c))
    bp = t1 @ between @ t1.T
    bp = 0.5 * (bp + bp.T)
    s, u = np.linalg.eigh(bp)
    order = np.argsort(-s, kind="stable")
    s, u = np.maximum(s[order], 0.0), u[:, order]
    return u.T @ t1, s, c @ u       # A, psi, A^-1


def em_iter_diag(scatter, means, counts, weights, mu, within, between):
    a, psi, ainv = joint_diag(within, between)
    u = (means - mu) @ a.T
    n = counts.astype(np.float64)[:, None]
    r = psi[None, :] / (1.0 + n * psi[None, :])           # diag of A mixed A^T
    g = n * r
    gu = g * u
    hu = (1.0 - g) * u
    w = weights[:, None]
    bs = np.diag(np.sum(w * r, axis=0)) + (w * gu).T @ gu
    ws = np.diag(np.sum(w * n * r, axis=0)) + (w * n * hu).T @ hu
    b_stats = ainv @ bs @ ainv.T
    w_stats = scatter + ainv @ ws @ ainv.T
    return w_stats, b_stats


# --------------------------------------------------------------------------- #
# EER (definition of scoring/eer.py:68-73: threshold minimising |FAR-FRR|,
# report (FAR+FRR)/2*100).  bob.measure is not installed; restated.
# --------------------------------------------------------------------------- #
def eer_percent(target_scores: np.ndarray, nontarget_scores: np.ndarray) -> float:
    tar = np.sort(np.asarray(target_scores, dtype=np.float64))
    non = np.sort(np.asarray(nontarget_scores, dtype=np.float64))
    thr = np.unique(np.concatenate([tar, non]))
    # accept if score >= thr  (bob.measure.farfrr convention)
    far = 1.0 - np.searchsorted(non, thr, side="left") / non.shape[0]
    frr = np.searchsorted(tar, thr, side="left") / tar.shape[0]
    i = int(np.argmin(np.abs(far - frr)))
    return float((far[i] + frr[i]) / 2.0 * 100.0)


# --------------------------------------------------------------------------- #
# Synthetic generator
# --------------------------------------------------------------------------- #
def two_cov_generator(d: int, seed: int = 1234):
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rng.randn(d, d))
    spec = 2.0 * np.exp(-np.arange(d) / (0.15 * d))
    return q * np.sqrt(spec)[None, :]       # A_b = Q diag(sqrt(spec))


def synth_speakers(a_b: np.ndarray, counts: Sequence[int], seed: int):
    """x = 0.5*1 + A_b z_spk + e ; returns (X, labels uint64, z)."""
    rng = np.random.RandomState(seed)
    d = a_b.shape[0]
    counts = np.asarray(counts, dtype=np.int64)
    k = counts.shape[0]
    z = rng.randn(k, d)
    labels = np.repeat(np.arange(k), counts)
    x = 0.5 + (z @ a_b.T)[labels] + rng.randn(labels.shape[0], d)
    return x, labels.astype(np.uint64), z
