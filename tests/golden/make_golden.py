"""Generate the committed golden fixtures.  Run in the authoring container:

    python tests/golden/make_golden.py

* ``lda_<solver>.npz`` -- produced by the REFERENCE's own ``python/liblda/lda.py``
  (imported from /root/reference via oracle/ref_lda.py): inputs + coef / intercept /
  log-probas.  These pin ``oracle/lda_port.py`` and the CUDA LDA path.
* ``plda_small.npz`` -- produced by ``oracle/kaldi_plda.py`` (the reference's PLDA
  cannot be built here: Kaldi + ATLAS + CPython 2 are absent), so it is a
  REGRESSION pin of the oracle, not a reference-generated vector ("parity
  unpinned", see oracle/kaldi_plda.py header).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import kaldi_plda as kp   # noqa: E402
from oracle import ref_lda            # noqa: E402


def make_lda():
    ref = ref_lda.load()
    rng = np.random.RandomState(7)
    n, d, k, nt = 600, 12, 6, 40
    centers = rng.randn(k, d) * 1.5
    y = np.arange(n) % k
    x = centers[y] + rng.randn(n, d)
    xt = centers[np.arange(nt) % k] + rng.randn(nt, d)
    for solver in ("svd", "lsqr", "eigen"):
        m = ref.LDA(solver=solver)
        m.fit(x, y)
        out = dict(x=x, y=y, xt=xt, coef=m._coef, intercept=m._intercept, priors=m.priors,
                   decision=m.decision_function(xt), log_proba=m.predict_log_proba(xt),
                   proba=m.predict_proba(xt))
        if solver == "svd":
            out["xbar"] = m._xbar
            out["scalings"] = m._scalings
        np.savez_compressed(os.path.join(HERE, "lda_%s.npz" % solver), **out)
    # binary case + python/test.py shapes (100x10, 2 and 3 classes)
    rng = np.random.RandomState(11)
    x = rng.normal(size=(100, 10))
    t = rng.normal(size=(100, 10))
    out = dict(x=x, t=t)
    for kk in (2, 3):
        yk = np.arange(100) % kk
        m = ref.LDA()
        m.fit(x, yk)
        out["log_proba%d" % kk] = m.predict_log_proba(t)
    np.savez_compressed(os.path.join(HERE, "lda_pytest_shapes.npz"), **out)


def make_lda_eigen_full():
    """eigen solver with MORE classes than dimensions (K - 1 >= d): every generalised eigenvalue is simple, so the
    reference's coef / intercept / decision values are reproducible and can pin an independent eigensolver."""
    ref = ref_lda.load()
    rng = np.random.RandomState(13)
    n, d, k, nt = 1500, 6, 15, 50
    centers = rng.randn(k, d) * 1.2
    y = np.arange(n) % k
    x = centers[y] + rng.randn(n, d)
    xt = centers[np.arange(nt) % k] + rng.randn(nt, d)
    m = ref.LDA(solver="eigen")
    m.fit(x, y)
    np.savez_compressed(os.path.join(HERE, "lda_eigen_full.npz"), x=x, y=y, xt=xt, coef=m._coef,
                        intercept=m._intercept, priors=m.priors, scalings=m._scalings,
                        explained_variance_ratio=m.explained_variance_ratio_, decision=m.decision_function(xt),
                        log_proba=m.predict_log_proba(xt), proba=m.predict_proba(xt), transform=m.transform(xt))


def reference_dvector_functions():
    """The reference's own pooling functions: lines 19-58 of scoring/extractdvector.py executed as they are (the
    module itself cannot be imported under Python 3: cPickle, htkfeature)."""
    path = "/root/reference/scoring/extractdvector.py"
    with open(path) as f:
        lines = f.readlines()
    ns = {"np": np}
    exec(compile("".join(lines[18:58]), path, "exec"), ns)
    return ns


def make_dvector():
    ns = reference_dvector_functions()
    rng = np.random.RandomState(17)
    d = 40
    lens = [1, 7, 33, 64, 5, 130, 2]
    utts = [rng.randn(n, d) * rng.uniform(0.5, 3.0) + rng.uniform(-1, 1) for n in lens]
    out = dict(frames=np.concatenate(utts), offsets=np.concatenate([[0], np.cumsum(lens)]))
    for name in ("extractdvectormean", "extractdvectormax", "extractdvectorvar", "extractdvectormean_nol2",
                 "extractdvectorvar_nol2", "extractdvectormax_nol2"):
        out[name] = np.stack([np.asarray(ns[name](u)).reshape(-1) for u in utts])
    out["shape_nol2"] = np.array(np.asarray(ns["extractdvectormean_nol2"](utts[2])).shape)
    np.savez_compressed(os.path.join(HERE, "dvector_pool.npz"), **out)


def make_plda():
    d = 24
    a_b = kp.two_cov_generator(d, seed=1234)
    counts = [3, 5, 5, 2, 7, 4, 4, 6, 3, 5, 8, 2, 4, 4, 5, 6]
    x, labels, _ = kp.synth_speakers(a_b, counts, seed=1234)
    m = kp.MPlda()
    m.fit(x, labels, 5)
    xe, le, _ = kp.synth_speakers(a_b, [3, 1, 2, 3, 3, 1, 2, 4], seed=1235)
    xt, lt, _ = kp.synth_speakers(a_b, [1] * 20, seed=1236)
    te = m.transform(xe, le)
    tt = m.transform(xt, lt)
    scores = np.array([[m.score(k, te[k], tt[j]) for j in sorted(tt)] for k in sorted(te)])
    bkg, _, _ = kp.synth_speakers(a_b, [1] * 30, seed=1237)
    m.norm(bkg, te)
    zscores = np.array([[m.score(k, te[k], tt[j]) for j in sorted(tt)] for k in sorted(te)])
    np.savez_compressed(
        os.path.join(HERE, "plda_small.npz"),
        x=x, labels=labels, iters=5, xe=xe, le=le, xt=xt, lt=lt, bkg=bkg,
        mean=m.plda.mean, psi=m.plda.psi, transform=m.plda.transform,
        enrol=np.stack([te[k][1] for k in sorted(te)]),
        enrol_counts=np.array([te[k][0] for k in sorted(te)]),
        test=np.stack([tt[k][1] for k in sorted(tt)]),
        scores=scores, zscores=zscores,
        meanz=np.array([m.meanz[k] for k in sorted(te)]),
        stdvz=np.array([m.stdvz[k] for k in sorted(te)]))


if __name__ == "__main__":
    if "--only-eigen-full" in sys.argv:
        make_lda_eigen_full()
        sys.exit(0)
    if "--only-dvector" in sys.argv:
        make_dvector()
        sys.exit(0)
    make_lda()
    make_lda_eigen_full()
    make_dvector()
    make_plda()
    print("golden fixtures written to", HERE)
