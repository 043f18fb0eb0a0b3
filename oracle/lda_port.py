"""CPU fp64 ORACLE for the LDA path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's ``python/liblda/lda.py`` (``LDA.fit`` with the
``svd`` / ``lsqr`` / ``eigen`` solvers, ``decision_function``, ``predict_log_proba``,
``predict_proba``, ``transform``).  PINNED: ``tests/test_oracle_lda.py`` checks it
against the reference file itself (loaded through ``oracle/ref_lda.py`` in the
authoring container) and against the committed fixtures in ``tests/golden/lda_*.npz``
that the reference produced (``tests/golden/make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this.
"""
import numpy as np
from scipy.linalg import eigh
from scipy.special import logsumexp


def class_means(x, y):
    """``_class_means`` (``python/liblda/lda.py:53-71``)."""
    classes = np.unique(y)
    return np.asarray([x[y == g, :].mean(0) for g in classes])


def empirical_covariance(x):
    """``empirical_covariance`` (``lda.py:19-50``) for >1 rows."""
    x = np.asarray(x)
    cov = np.cov(x.T, bias=1)
    if cov.ndim == 0:
        cov = np.array([[cov]])
    return cov


def class_cov(x, y, priors):
    """``_class_cov`` (``lda.py:10-16``)."""
    classes = np.unique(y)
    covs = [np.atleast_2d(empirical_covariance(x[y == g, :])) for g in classes]
    return np.average(covs, axis=0, weights=priors)


class LDAOracle:
    def __init__(self, solver="svd", priors=None):
        self.priors = priors
        self.solver = solver

    def fit(self, features, labels):
        """``LDA.fit`` (``lda.py:106-138``)."""
        self._classes = np.unique(labels)
        if self.priors is None:
            _, y_t = np.unique(labels, return_inverse=True)
            self.priors = np.bincount(y_t) / float(len(labels))
        else:
            self.priors = np.asarray(self.priors)
        if self.priors.sum() != 1:
            self.priors = self.priors / self.priors.sum()
        getattr(self, "_solve_" + self.solver)(np.asarray(features, dtype=np.float64), np.asarray(labels))

    def _solve_svd(self, x, y):
        """``_solve_svd`` (``lda.py:178-221``)."""
        n_samples, _ = x.shape
        n_classes = len(self._classes)
        tol = 1e-4
        self._means = class_means(x, y)
        xc = np.concatenate([x[y == g, :] - self._means[i] for i, g in enumerate(self._classes)], axis=0)
        self._xbar = np.dot(self.priors, self._means)
        stddev = xc.std(axis=0)
        stddev[stddev == 0] = 1.0
        fac = 1.0 / (n_samples - n_classes)
        xs = np.sqrt(fac) * (xc / stddev)
        _, s, v = np.linalg.svd(xs, full_matrices=False)
        rank = np.sum(s > tol)
        scalings = (v[:rank] / stddev).T / s[:rank]
        xm = np.dot(((np.sqrt((n_samples * self.priors) * fac)) * (self._means - self._xbar).T).T, scalings)
        _, s2, v2 = np.linalg.svd(xm, full_matrices=0)
        rank2 = np.sum(s2 > tol * s2[0])
        self._scalings = np.dot(scalings, v2.T[:, :rank2])
        coef = np.dot(self._means - self._xbar, self._scalings)
        self._intercept = -0.5 * np.sum(coef ** 2, axis=1) + np.log(self.priors)
        self._coef = np.dot(coef, self._scalings.T)
        self._intercept -= np.dot(self._xbar, self._coef.T)

    def _solve_lsqr(self, x, y):
        """``_solve_lsqr`` (``lda.py:223-251``)."""
        self._means = class_means(x, y)
        cov = class_cov(x, y, self.priors)
        self._coef = np.linalg.lstsq(cov, self._means.T, rcond=-1)[0].T
        self._intercept = -0.5 * np.diag(np.dot(self._means, self._coef.T)) + np.log(self.priors)

    def _solve_eigen(self, x, y):
        """``_solve_eigen`` (``lda.py:140-176``)."""
        self._means = class_means(x, y)
        sw = class_cov(x, y, self.priors)
        st = empirical_covariance(x)
        sb = st - sw
        evals, evecs = eigh(sb, sw)
        evecs = evecs[:, np.argsort(evals)[::-1]]
        evecs /= np.apply_along_axis(np.linalg.norm, 0, evecs)
        self._scalings = evecs
        self._coef = np.dot(self._means, evecs).dot(evecs.T)
        self._intercept = -0.5 * np.diag(np.dot(self._means, self._coef.T)) + np.log(self.priors)

    def decision_function(self, x):
        """``decision_function`` (``lda.py:253-279``)."""
        if not hasattr(self, "_coef") or self._coef is None:
            raise ValueError("This LDA instance is not fitted yet")
        if x.shape[1] != self._coef.shape[1]:
            raise ValueError("X has %d features per sample; expecting %d" % (x.shape[1], self._coef.shape[1]))
        scores = np.dot(x, self._coef.T) + self._intercept
        return scores.ravel() if scores.shape[1] == 1 else scores

    def predict_log_proba(self, sample):
        """``predict_log_proba`` (``lda.py:306-325``)."""
        values = self.decision_function(sample)
        llk = values - values.max(axis=1)[:, np.newaxis]
        return llk - logsumexp(llk, axis=1)[:, np.newaxis]

    def predict_proba(self, sample):
        """``predict_proba`` (``lda.py:281-304``)."""
        prob = self.decision_function(sample)
        prob = 1.0 / (1.0 + np.exp(-prob))
        if len(self._classes) == 2:
            return np.column_stack([1 - prob, prob])
        return prob / prob.sum(axis=1).reshape((prob.shape[0], -1))

    def transform(self, x, n_components=None):
        """Evident intent of ``transform`` (``lda.py:328-349``; the svd branch is
        unreachable in the reference -- SURVEY App. B)."""
        if self.solver == "lsqr":
            raise NotImplementedError("transform not implemented for 'lsqr' solver (use 'svd' or 'eigen').")
        if self.solver == "svd":
            x_new = np.dot(x - self._xbar, self._scalings)
        else:
            x_new = np.dot(x, self._scalings)
        n_components = x.shape[1] if n_components is None else n_components
        return x_new[:, :n_components]
