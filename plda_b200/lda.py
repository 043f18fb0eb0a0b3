"""``LDA`` -- host-side mirror of the reference's ``liblda.LDA``
(``python/liblda/lda.py:88-349``) with the arithmetic on the device behind the C ABI
(``lda_*`` entry points of ``include/plda_b200.h``).

Hot-path rows of SURVEY.md section 8 implemented here: ``fit`` with the default ``svd``
solver (``lda.py:178-221``), ``decision_function`` (``:253-279``) and
``predict_log_proba`` (``:306-325``).  Section 8f "next" rows built on the same kernels: the
``lsqr`` solver (``:223-251``), ``predict_proba`` (``:281-304``) and ``transform`` (``:328-349``,
evident intent -- the reference's svd branch is unreachable) and the ``eigen`` solver (``:140-176``).
For ``eigen`` with K - 1 < d the generalised eigenvalue 0 is degenerate and the basis of its eigenspace is
arbitrary (LAPACK's choice in the reference, the Jacobi solver's here): that part only moves all decision
values of a sample by one common amount, so ``predict`` / ``predict_log_proba`` agree with the reference,
raw ``coef`` / ``decision_function`` / ``predict_proba`` (a sigmoid of the raw values) only when K > d.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

_SOLVERS = {"svd": 0, "lsqr": 1, "eigen": 2}


class LDA(object):
    def __init__(self, solver="svd", priors=None, device: int = 0, precision: str = "bf16x3"):
        self.priors = priors
        self.solver = solver
        self._lib = _ffi.lib()
        h = C.c_void_p()
        _ffi.check(self._lib.lda_create(int(device), C.byref(h)))
        self._h = h
        code = {"bf16x3": _ffi.PREC_BF16X3, "fp64": _ffi.PREC_FP64}.get(precision)
        if code is None:
            raise ValueError("precision must be 'bf16x3' or 'fp64'")
        _ffi.check(self._lib.lda_set_precision(self._h, code))
        self._coef = None
        self._intercept = None
        self._classes = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.lda_destroy(h)
            except Exception:
                pass
            self._h = None

    def launch_count(self) -> int:
        n = C.c_int64()
        _ffi.check(self._lib.lda_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def _after_torch(self, t) -> None:
        """Order the handle's stream after torch's current stream on ``t``'s device (``lda_stream_wait``; see
        ``PLDA._after_torch`` for the stream contract of CUDA-tensor operands)."""
        import torch
        s = torch.cuda.current_stream(t.device)
        _ffi.check(self._lib.lda_stream_wait(self._h, C.c_void_p(s.cuda_stream)))

    # ------------------------------------------------------------------ fit
    def fit(self, features, labels):
        """``LDA.fit`` (``lda.py:106-138``).  Returns None."""
        if self.solver not in _SOLVERS:
            # the reference silently fits nothing for an unknown solver (lda.py:131-138) and fails later
            raise ValueError("unknown solver %r (expected 'svd', 'lsqr' or 'eigen')" % (self.solver,))
        pri = None
        if self.priors is not None:
            pri = np.ascontiguousarray(self.priors, dtype=np.float64)
        fit_fn = {"svd": self._lib.lda_fit_svd, "lsqr": self._lib.lda_fit_lsqr,
                  "eigen": self._lib.lda_fit_eigen}[self.solver]
        if hasattr(features, "is_cuda") and features.is_cuda:
            # resident rows (superset): no host round trip of the N x d matrix; labels are 8 B / row on the host
            import torch
            x = features
            if x.dim() != 2 or x.dtype not in (torch.float32, torch.float64):
                raise ValueError("features must be a 2-D float32 / float64 tensor")
            if x.stride(1) != 1:
                x = x.contiguous()
            y = labels.detach().cpu().numpy() if hasattr(labels, "detach") else np.asarray(labels)
            if y.dtype.kind not in "iub":
                raise ValueError("labels must be integers")
            y = np.ascontiguousarray(y.astype(np.int64).reshape(-1))
            n, d = x.shape
            if y.shape[0] != n:
                raise ValueError("labels and features disagree on the number of samples")
            self._after_torch(x)
            _ffi.check(fit_fn(self._h, C.c_void_p(x.data_ptr()), n, d, x.stride(0),
                              _ffi.F32 if x.dtype == torch.float32 else _ffi.F64, _ffi.DEVICE, _ffi.ptr(y),
                              _ffi.ptr(pri), 0 if pri is None else pri.shape[0]))
            _, cnt = np.unique(y, return_counts=True)
            self._read_back(cnt, n)
            return None
        x, dtype, y = self._check_xy(features, labels)
        n, d = x.shape
        _ffi.check(fit_fn(self._h, _ffi.ptr(x), n, d, d, dtype, _ffi.HOST, _ffi.ptr(y), _ffi.ptr(pri),
                          0 if pri is None else pri.shape[0]))
        _, cnt = np.unique(y, return_counts=True)
        self._read_back(cnt, n)
        return None

    @staticmethod
    def _check_xy(features, labels):
        x, dtype = _ffi.as_matrix(np.asarray(features, dtype=np.float64) if np.asarray(features).dtype.kind != "f"
                                  else features, "features")
        y = np.asarray(labels)
        if y.dtype.kind not in "iub":
            raise ValueError("labels must be integers")
        y = np.ascontiguousarray(y.astype(np.int64).reshape(-1))
        if y.shape[0] != x.shape[0]:
            raise ValueError("labels and features disagree on the number of samples")
        return x, dtype, y

    def _read_back(self, class_counts, n):
        """coef / intercept / classes (and the svd state) from the handle; priors as the reference leaves them."""
        k = C.c_int64()
        dd = C.c_int64()
        _ffi.check(self._lib.lda_num_classes(self._h, C.byref(k), C.byref(dd)))
        self._coef = np.empty((k.value, dd.value))
        self._intercept = np.empty(k.value)
        self._classes = np.empty(k.value, dtype=np.int64)
        _ffi.check(self._lib.lda_get_coef(self._h, _ffi.ptr(self._coef), _ffi.ptr(self._intercept),
                                          _ffi.ptr(self._classes)))
        self._xbar = self._scalings = None
        self.explained_variance_ratio_ = None
        if self.solver == "eigen":
            ev = np.empty(dd.value)
            nev = C.c_int64()
            _ffi.check(self._lib.lda_get_eigenvalues(self._h, _ffi.ptr(ev), dd.value, C.byref(nev)))
            self.explained_variance_ratio_ = np.sort(ev / np.sum(ev))[::-1]        # lda.py:167
        if self.solver in ("svd", "eigen"):
            rank = C.c_int64()
            _ffi.check(self._lib.lda_get_svd(self._h, C.byref(rank), None, None))
            self._xbar = np.empty(dd.value)
            self._scalings = np.empty((dd.value, rank.value))
            _ffi.check(self._lib.lda_get_svd(self._h, C.byref(rank), _ffi.ptr(self._xbar), _ffi.ptr(self._scalings)))
        if self.priors is None:
            self.priors = np.asarray(class_counts) / float(n)
        else:
            p = np.asarray(self.priors, dtype=np.float64)
            self.priors = p / p.sum() if p.sum() != 1 else p

    def fit_distributed(self, features, labels, group=None):
        """Sharded ``fit``: every rank passes ITS rows; all rows of a class must live on one rank.

        Per rank one pass over its rows (class means + within scatter on the device), then ONE sum-all-reduce of the
        ``d x d`` scatter and ONE all-gather of the per-class rows (``plda_b200.dist.merge_class_stats``); the small
        solver runs replicated on every rank, so all ranks end with identical coefficients."""
        from . import dist as _dist
        if self.solver not in _SOLVERS:
            raise ValueError("unknown solver %r (expected 'svd', 'lsqr' or 'eigen')" % (self.solver,))
        sw, means, counts, classes = self.local_class_stats(features, labels)
        sw, means, counts, classes = _dist.merge_class_stats(sw, means, counts, classes, group)
        self.fit_from_stats(sw, means, counts, classes)
        return None

    def local_class_stats(self, features, labels):
        """Device pass over THIS rank's rows: ``(sw [d,d], means [k,d], counts [k], classes [k])``."""
        x, dtype, y = self._check_xy(features, labels)
        n_loc, d = x.shape
        k = C.c_int64()
        _ffi.check(self._lib.lda_class_stats(self._h, _ffi.ptr(x), n_loc, d, d, dtype, _ffi.HOST, _ffi.ptr(y),
                                             C.byref(k)))
        sw = np.empty((d, d))
        means = np.empty((k.value, d))
        counts = np.empty(k.value, dtype=np.int64)
        classes = np.empty(k.value, dtype=np.int64)
        _ffi.check(self._lib.lda_get_class_stats(self._h, _ffi.ptr(sw), _ffi.ptr(means), _ffi.ptr(counts),
                                                 _ffi.ptr(classes)))
        return sw, means, counts, classes

    def fit_from_stats(self, sw, means, counts, classes):
        """Solver stage on merged class statistics (see ``fit_distributed``)."""
        sw = np.ascontiguousarray(sw, dtype=np.float64)
        means = np.ascontiguousarray(means, dtype=np.float64)
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        classes = np.ascontiguousarray(classes, dtype=np.int64)
        kk, d = means.shape
        n = int(counts.sum())
        pri = None
        if self.priors is not None:
            pri = np.ascontiguousarray(self.priors, dtype=np.float64)
        _ffi.check(self._lib.lda_fit_from_stats(self._h, _SOLVERS[self.solver], n, kk, d, _ffi.ptr(sw),
                                                _ffi.ptr(means), _ffi.ptr(counts), _ffi.ptr(classes), _ffi.ptr(pri),
                                                0 if pri is None else pri.shape[0]))
        self._read_back(counts, n)

    def set_coef(self, coef, intercept, classes=None):
        """Install decision coefficients (replica of a model fitted elsewhere).  ``classes``: the label value of every
        row of ``coef`` (default ``arange(K)``) -- what ``predict`` returns."""
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        intercept = np.ascontiguousarray(intercept, dtype=np.float64)
        _ffi.check(self._lib.lda_set_coef(self._h, coef.shape[0], coef.shape[1], _ffi.ptr(coef), _ffi.ptr(intercept)))
        self._coef, self._intercept = coef, intercept
        self._classes = np.arange(coef.shape[0]) if classes is None else np.asarray(classes, dtype=np.int64).copy()
        if self._classes.shape[0] != coef.shape[0]:
            raise ValueError("set_coef: one class label per coefficient row")
        self._xbar = self._scalings = None

    # ------------------------------------------------------------------ predict
    def _predict(self, x, log_proba):
        if self._coef is None:
            raise ValueError("This %(name)s instance is not fitted yet" % {"name": type(self).__name__})
        if hasattr(x, "is_cuda") and x.is_cuda:
            import torch
            if x.dtype not in (torch.float32, torch.float64):
                raise ValueError("features must be float tensors")
            if x.stride(1) != 1:
                x = x.contiguous()
            nt, d = x.shape
            k = self._coef.shape[0]
            ldo = (k + 3) // 4 * 4
            buf = torch.empty((nt, ldo), dtype=torch.float32, device=x.device)
            self._after_torch(x)
            _ffi.check(self._lib.lda_predict(self._h, C.c_void_p(x.data_ptr()), nt, d, x.stride(0),
                                             _ffi.F32 if x.dtype == torch.float32 else _ffi.F64, _ffi.DEVICE,
                                             int(log_proba), C.c_void_p(buf.data_ptr()), ldo, _ffi.DEVICE))
            return buf[:, :k]
        xa = np.asarray(x)
        if xa.dtype.kind != "f":
            xa = xa.astype(np.float64)
        xa, dtype = _ffi.as_matrix(xa, "sample")
        nt, d = xa.shape
        k = self._coef.shape[0]
        out = np.empty((nt, k), dtype=np.float32)
        _ffi.check(self._lib.lda_predict(self._h, _ffi.ptr(xa), nt, d, d, dtype, _ffi.HOST, int(log_proba),
                                         _ffi.ptr(out), k, _ffi.HOST))
        return out

    def decision_function(self, X):
        """``decision_function`` (``lda.py:253-279``): ``X coef^T + intercept`` (float32 out)."""
        scores = self._predict(X, 0)
        return scores.ravel() if scores.shape[1] == 1 else scores

    def predict_log_proba(self, sample):
        """``predict_log_proba`` (``lda.py:306-325``): row log-softmax of the decision values."""
        return self._predict(sample, 1)

    def predict_proba(self, sample):
        """``predict_proba`` (``lda.py:281-304``): OvR sigmoid of the decision values, row-normalised for more than
        two classes; computed on the device.  Two classes: ``column_stack([1 - p, p])`` like the reference."""
        prob = self._predict(sample, 2)
        if len(self._classes) == 2:
            if hasattr(prob, "is_cuda"):
                import torch
                return torch.cat([1 - prob, prob], dim=1)
            return np.column_stack([1 - prob, prob])
        return prob

    def transform(self, X, n_components=None):
        """``transform`` (``lda.py:328-349``): svd solver ``(X - xbar) @ scalings[:, :n_components]`` (the
        reference's svd branch is unreachable -- SURVEY App. B -- this is its evident intent); eigen solver
        ``X @ scalings[:, :n_components]`` (the handle keeps ``xbar = 0`` for it)."""
        if self.solver == "lsqr":
            raise NotImplementedError("transform not implemented for 'lsqr' solver (use 'svd' or 'eigen').")
        if self._coef is None:
            raise ValueError("This %(name)s instance is not fitted yet" % {"name": type(self).__name__})
        if getattr(self, "_scalings", None) is None:
            raise ValueError("transform needs the fitted projection (this instance only holds coefficients installed "
                             "with set_coef)")
        xa = np.asarray(X)
        if xa.dtype.kind != "f":
            xa = xa.astype(np.float64)
        xa, dtype = _ffi.as_matrix(xa, "X")
        nt, d = xa.shape
        rank = self._scalings.shape[1]
        # the reference slices X_new[:, :n_components] with n_components defaulting to X.shape[1]
        n_comp = min(rank, d if n_components is None else int(n_components))
        out = np.empty((nt, n_comp), dtype=np.float32)
        _ffi.check(self._lib.lda_transform(self._h, _ffi.ptr(xa), nt, d, d, dtype, _ffi.HOST, n_comp, _ffi.ptr(out),
                                           n_comp, _ffi.HOST))
        return out

    def predict(self, sample):
        return self._classes[np.asarray(self.decision_function(sample)).argmax(axis=1)]
