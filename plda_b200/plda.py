"""``PLDA`` -- host-side mirror of the reference's ``liblda.PLDA``
(``python/liblda/plda.py:4-51``), which delegates 1:1 to the native ``MPlda`` type
(``src/pldamodule.cpp:280-295``).  Same method names, argument meaning, return
shapes and error behaviour; every method calls through ctypes into the C ABI of
``include/plda_b200.h`` (hand-written sm_100a kernels).  No CPU fallback.

Drop-in methods (reference signatures):
    fit(x, y, iters=10)                      -> None
    transform(x, y)                          -> {label: (n, ndarray[d] float64)}
    norm(vectors, transformedvecs, numutts=0)-> None
    score(target, xvec, yvec)                -> float (rounded through float32)

Batched supersets used at scale (SURVEY.md hard part 7):
    transform_batch(x, counts=None, ...)     -> (R, dim) array
    score_grid(enrol, enrol_counts, test, ...) -> (Ne, Nt) float32 array
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and hasattr(x, "data_ptr") and bool(x.is_cuda)


class PLDA(object):
    def __init__(self, device: int = 0, precision: str = "bf16x3"):
        self._lib = _ffi.lib()
        h = C.c_void_p()
        _ffi.check(self._lib.plda_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.set_precision(precision)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.plda_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ config
    def set_precision(self, precision: str) -> None:
        """'bf16x3' (tcgen05, default) or 'fp64' (exact SIMT mode)."""
        code = {"bf16x3": _ffi.PREC_BF16X3, "fp64": _ffi.PREC_FP64}.get(precision)
        if code is None:
            raise ValueError("precision must be 'bf16x3' or 'fp64'")
        _ffi.check(self._lib.plda_set_precision(self._h, code))
        self.precision = precision

    def launch_count(self) -> int:
        n = C.c_int64()
        _ffi.check(self._lib.plda_launch_count(self._h, C.byref(n)))
        return int(n.value)

    # ------------------------------------------------------------------ fit
    def fit(self, x, y, iters=10):
        """``MPlda_fit`` (``src/pldamodule.cpp:42-109``).  x: (n, d) float array (or a CUDA
        torch tensor), y: (n,) unsigned labels, iters: EM iterations.  Returns None.
        Raises ValueError for non-float features, string / negative labels or a single
        speaker -- the reference's three ValueErrors."""
        iters = int(iters)
        if _is_torch_cuda(x):
            xt, dtype = _torch_matrix(x)
            n, d = xt.shape
            lab = _ffi.as_labels(_to_numpy_labels(y), n)
            _ffi.check(self._lib.plda_fit(self._h, C.c_void_p(xt.data_ptr()), n, d, xt.stride(0), dtype, _ffi.DEVICE,
                                          _ffi.ptr(lab), iters))
            return None
        xa, dtype = _ffi.as_matrix(x, "features")
        lab = _ffi.as_labels(y, xa.shape[0])
        n, d = xa.shape
        _ffi.check(self._lib.plda_fit(self._h, _ffi.ptr(xa), n, d, d, dtype, _ffi.HOST, _ffi.ptr(lab), iters))
        return None

    def fit_distributed(self, x, y, iters=10, group=None):
        """Sharded fit: every rank passes ITS rows (whole speakers per rank -- a speaker must
        not be split across ranks; label values only need to be unique within a rank).  The stats pass exchanges
        one all-reduce (scatter, weighted mean sum, class weight, class count), every EM iteration one all-reduce of
        the two d x d statistics; the d x d factorizations are replicated.  Requires an initialised NCCL process
        group; without one this is `fit`."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self.fit(x, y, iters)
        d = int(x.shape[1])
        dev = torch.device("cuda", self.device)
        scratch = torch.zeros(2 * d * d + d + 2, dtype=torch.float64, device=dev)
        stream = torch.cuda.Stream(device=dev)
        failure = []

        def _cb(_user, count):
            try:
                dist.all_reduce(scratch[:count], group=group)
                return 0
            except Exception as e:  # surfaced as PLDA_E_INTERNAL by the library
                failure.append(e)
                return 1

        cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64)(_cb)
        torch.cuda.synchronize(dev)
        with torch.cuda.stream(stream):
            _ffi.check(self._lib.plda_set_stream(self._h, C.c_void_p(stream.cuda_stream)))
            _ffi.check(self._lib.plda_set_allreduce(self._h, C.cast(cb, C.c_void_p), None,
                                                    C.c_void_p(scratch.data_ptr()), scratch.numel()))
            try:
                self.fit(x, y, iters)
            finally:
                self._lib.plda_set_allreduce(self._h, None, None, None, 0)
                self._lib.plda_set_stream(self._h, C.c_void_p(None))
        stream.synchronize()
        if failure:
            raise failure[0]
        return None

    def fit_timings(self):
        """ms: dict(stats=, em=, output=, total=, iters=) of the last fit."""
        out = (C.c_double * 5)()
        _ffi.check(self._lib.plda_fit_timings(self._h, out))
        return dict(stats=out[0], em=out[1], output=out[2], total=out[3], iters=int(out[4]))

    # ------------------------------------------------------------------ model
    @property
    def dim(self) -> int:
        d = C.c_int64()
        _ffi.check(self._lib.plda_dim(self._h, C.byref(d)))
        return int(d.value)

    def get_model(self):
        """(mean[d], transform[d,d], psi[d]) -- Kaldi ``Plda{mean_, transform_, psi_}``."""
        d = self.dim
        if d == 0:
            raise ValueError("PLDA model is not fitted")
        mean = np.empty(d)
        tr = np.empty((d, d))
        psi = np.empty(d)
        _ffi.check(self._lib.plda_get_model(self._h, _ffi.ptr(mean), _ffi.ptr(tr), _ffi.ptr(psi)))
        return mean, tr, psi

    def set_model(self, mean, transform, psi) -> None:
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        tr = np.ascontiguousarray(transform, dtype=np.float64)
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        d = mean.shape[0]
        if tr.shape != (d, d) or psi.shape != (d,):
            raise ValueError("set_model: inconsistent shapes")
        _ffi.check(self._lib.plda_set_model(self._h, d, _ffi.ptr(mean), _ffi.ptr(tr), _ffi.ptr(psi)))

    def get_covariances(self):
        d = self.dim
        w = np.empty((d, d))
        b = np.empty((d, d))
        _ffi.check(self._lib.plda_get_covariances(self._h, _ffi.ptr(w), _ffi.ptr(b)))
        return w, b

    def save(self, path) -> None:
        """npz checkpoint of the model and z-norm tables (the reference has none)."""
        mean, tr, psi = self.get_model()
        ids, zm, zs = self.znorm_tables()
        np.savez(path, mean=mean, transform=tr, psi=psi, z_ids=ids, z_mean=zm, z_std=zs)

    def load(self, path) -> None:
        g = np.load(path)
        self.set_model(g["mean"], g["transform"], g["psi"])
        ids = np.ascontiguousarray(g["z_ids"], dtype=np.uint64)
        zm = np.ascontiguousarray(g["z_mean"], dtype=np.float64)
        zs = np.ascontiguousarray(g["z_std"], dtype=np.float64)
        _ffi.check(self._lib.plda_znorm_clear(self._h))
        _ffi.check(self._lib.plda_znorm_set(self._h, _ffi.ptr(ids), _ffi.ptr(zm), _ffi.ptr(zs), ids.shape[0]))

    # ------------------------------------------------------------------ transform
    def transform(self, x, y, targetdim=0, smoothing=1.0):
        """``Mplda_transform`` (``src/pldamodule.cpp:111-194``): average the rows of each label,
        apply the PLDA transform with length normalisation, return ``{label: (n, vec)}`` with
        labels in ascending order.  ``smoothing != 1.0`` smooths the within-class covariance and
        MUTATES the model, as in the reference (:158-160).  ``targetdim`` keeps the leading
        directions (the reference's own plumbing for it is broken)."""
        xa, dtype = _ffi.as_matrix(x, "features")
        lab = _ffi.as_labels(y, xa.shape[0])
        if float(smoothing) != 1.0:
            _ffi.check(self._lib.plda_smooth(self._h, float(smoothing)))
        n, d = xa.shape
        out_dim = int(targetdim) if targetdim else d
        out_labels = np.empty(n, dtype=np.uint64)
        out_counts = np.empty(n, dtype=np.int64)
        out_vecs = np.empty((n, out_dim), dtype=np.float64)
        n_out = C.c_int64()
        _ffi.check(self._lib.plda_transform(self._h, _ffi.ptr(xa), n, d, d, dtype, _ffi.HOST, _ffi.ptr(lab),
                                            int(targetdim), _ffi.ptr(out_labels), _ffi.ptr(out_counts),
                                            _ffi.ptr(out_vecs), C.byref(n_out)))
        r = int(n_out.value)
        vecs = out_vecs[:r]
        return {int(out_labels[i]): (int(out_counts[i]), vecs[i].copy()) for i in range(r)}

    def transform_batch(self, x, counts=None, targetdim=0, out_dtype=np.float64):
        """Row-wise ``TransformIvector``: row r is already the average of ``counts[r]`` utterances
        (default 1).  numpy in -> numpy out; CUDA torch tensor in -> CUDA torch tensor out."""
        const_count = 1
        cnt = None
        if counts is not None:
            if np.isscalar(counts):
                const_count = int(counts)
            else:
                cnt = np.ascontiguousarray(counts, dtype=np.int32)
        if _is_torch_cuda(x):
            import torch
            xt, dtype = _torch_matrix(x)
            n, d = xt.shape
            dim = int(targetdim) if targetdim else d
            tdt = torch.float32 if np.dtype(out_dtype) == np.float32 else torch.float64
            out = torch.empty((n, dim), dtype=tdt, device=xt.device)
            _ffi.check(self._lib.plda_transform_rows(
                self._h, C.c_void_p(xt.data_ptr()), n, d, xt.stride(0), dtype, _ffi.DEVICE, _ffi.ptr(cnt), const_count,
                int(targetdim), C.c_void_p(out.data_ptr()), dim, _ffi.F32 if tdt == torch.float32 else _ffi.F64,
                _ffi.DEVICE))
            return out
        xa, dtype = _ffi.as_matrix(x, "features")
        n, d = xa.shape
        dim = int(targetdim) if targetdim else d
        odt = np.dtype(out_dtype)
        out = np.empty((n, dim), dtype=odt)
        _ffi.check(self._lib.plda_transform_rows(self._h, _ffi.ptr(xa), n, d, d, dtype, _ffi.HOST, _ffi.ptr(cnt),
                                                 const_count, int(targetdim), _ffi.ptr(out), dim,
                                                 _ffi.F32 if odt == np.float32 else _ffi.F64, _ffi.HOST))
        return out

    # ------------------------------------------------------------------ norm / score
    def norm(self, vectors, transformedvecs, numutts=0, seed=0):
        """``MPlda_norm`` (``src/pldamodule.cpp:196-256``): z-norm statistics of every enrol model
        in ``transformedvecs`` (the dict returned by ``transform``) against the RAW background
        ``vectors``.  Returns None."""
        bkg, dtype = _ffi.as_matrix(vectors, "vectors")
        if not isinstance(transformedvecs, dict):
            raise TypeError("norm() expects the dict returned by transform()")
        if len(transformedvecs) == 0:
            return None
        ids = np.fromiter((int(k) for k in transformedvecs.keys()), dtype=np.uint64, count=len(transformedvecs))
        enrol = np.ascontiguousarray(np.stack([np.asarray(v[1], dtype=np.float64) for v in transformedvecs.values()]))
        m, d = bkg.shape
        ne, dim = enrol.shape
        _ffi.check(self._lib.plda_norm(self._h, _ffi.ptr(bkg), m, d, d, dtype, _ffi.HOST, _ffi.ptr(ids),
                                       _ffi.ptr(enrol), ne, dim, dim, _ffi.F64, _ffi.HOST, int(numutts), int(seed)))
        return None

    def norm_batch(self, vectors, enrol_ids, enrol_vectors, numutts=0, seed=0):
        """Batched form of ``norm`` for large enrol sets: ``enrol_ids`` (uint64 ``[Ne]``) and ``enrol_vectors``
        (transformed, ``[Ne, dim]``) as arrays instead of the ``{id: (n, vec)}`` dict -- same statistics
        (``src/pldamodule.cpp:196-256``; the per-model utterance count is not used by the reference's ``norm`` either,
        ``:235`` scores with ``n = 1``).  numpy arrays or CUDA tensors (both operands on the same side)."""
        ids = np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        if _is_torch_cuda(vectors) or _is_torch_cuda(enrol_vectors):
            if not (_is_torch_cuda(vectors) and _is_torch_cuda(enrol_vectors)):
                raise ValueError("norm_batch: background and enrol vectors must both be CUDA tensors or both host arrays")
            bt, bdt = _torch_matrix(vectors)
            et, edt = _torch_matrix(enrol_vectors)
            if et.shape[0] != ids.shape[0]:
                raise ValueError("norm_batch: enrol_ids length mismatch")
            if et.shape[0] == 0:
                return None
            _ffi.check(self._lib.plda_norm(self._h, C.c_void_p(bt.data_ptr()), bt.shape[0], bt.shape[1], bt.stride(0),
                                           bdt, _ffi.DEVICE, _ffi.ptr(ids), C.c_void_p(et.data_ptr()), et.shape[0],
                                           et.stride(0), et.shape[1], edt, _ffi.DEVICE, int(numutts), int(seed)))
            return None
        bkg, bdt = _ffi.as_matrix(vectors, "vectors")
        enrol, edt = _ffi.as_matrix(enrol_vectors, "enrol_vectors")
        if enrol.shape[0] != ids.shape[0]:
            raise ValueError("norm_batch: enrol_ids length mismatch")
        if enrol.shape[0] == 0:
            return None
        _ffi.check(self._lib.plda_norm(self._h, _ffi.ptr(bkg), bkg.shape[0], bkg.shape[1], bkg.shape[1], bdt, _ffi.HOST,
                                       _ffi.ptr(ids), _ffi.ptr(enrol), enrol.shape[0], enrol.shape[1], enrol.shape[1],
                                       edt, _ffi.HOST, int(numutts), int(seed)))
        return None

    def znorm_tables(self):
        n = C.c_int64()
        _ffi.check(self._lib.plda_znorm_size(self._h, C.byref(n)))
        cap = int(n.value)
        ids = np.empty(cap, dtype=np.uint64)
        mean = np.empty(cap)
        std = np.empty(cap)
        got = C.c_int64()
        _ffi.check(self._lib.plda_znorm_get(self._h, _ffi.ptr(ids), _ffi.ptr(mean), _ffi.ptr(std), cap, C.byref(got)))
        order = np.argsort(ids[:got.value])
        return ids[order], mean[order], std[order]

    def score(self, target, xvec, yvec):
        """``MPlda_score`` (``src/pldamodule.cpp:258-277``): LLR of one (enrol, test) pair, both
        ``(n, vec)`` tuples from ``transform``; z-normalised iff ``target`` was seen by ``norm``."""
        n_e = int(xvec[0])
        e = np.ascontiguousarray(xvec[1], dtype=np.float64)
        t = np.ascontiguousarray(yvec[1], dtype=np.float64)
        if e.shape != t.shape or e.ndim != 1:
            raise ValueError("score: enrol and test vectors must be 1-D and of equal length")
        out = C.c_float()
        _ffi.check(self._lib.plda_score_pair(self._h, int(target), n_e, _ffi.ptr(e), _ffi.ptr(t), e.shape[0],
                                             C.byref(out)))
        return float(out.value)

    def score_grid(self, enrol, enrol_counts, test, enrol_ids=None, out=None):
        """All-pairs LLR grid: ``out[e, t] = LogLikelihoodRatio(enrol[e], enrol_counts[e], test[t])``
        as float32 (Ne, Nt).  Inputs are transformed vectors: numpy (host) or CUDA torch tensors
        (resident; the result is then a CUDA tensor too).  ``enrol_ids``: apply z-norm for ids seen
        by ``norm``."""
        cnt = np.ascontiguousarray(enrol_counts, dtype=np.int32).reshape(-1)
        ids = None if enrol_ids is None else np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        if _is_torch_cuda(enrol):
            import torch
            et, dtype = _torch_matrix(enrol)
            tt, dtype2 = _torch_matrix(test)
            if dtype != dtype2:
                raise ValueError("score_grid: enrol and test dtypes differ")
            ne, dim = et.shape
            nt = tt.shape[0]
            if cnt.shape[0] != ne:
                raise ValueError("score_grid: enrol_counts length mismatch")
            if out is None:
                ldo = (nt + 3) // 4 * 4
                buf = torch.empty((ne, ldo), dtype=torch.float32, device=et.device)
                out = buf[:, :nt]
            _ffi.check(self._lib.plda_score_grid(
                self._h, C.c_void_p(et.data_ptr()), ne, et.stride(0), _ffi.ptr(cnt), _ffi.ptr(ids),
                C.c_void_p(tt.data_ptr()), nt, tt.stride(0), dim, dtype, _ffi.DEVICE, C.c_void_p(out.data_ptr()),
                out.stride(0), _ffi.DEVICE))
            return out
        ea, dtype = _ffi.as_matrix(enrol, "enrol")
        ta, dtype2 = _ffi.as_matrix(test, "test")
        if dtype != dtype2:
            ea = ea.astype(np.float64)
            ta = ta.astype(np.float64)
            dtype = _ffi.F64
        ne, dim = ea.shape
        nt = ta.shape[0]
        if ta.shape[1] != dim:
            raise ValueError("score_grid: enrol and test dimensions differ")
        if cnt.shape[0] != ne:
            raise ValueError("score_grid: enrol_counts length mismatch")
        if out is None:
            out = np.empty((ne, nt), dtype=np.float32)
        if ne == 0 or nt == 0:
            return out
        _ffi.check(self._lib.plda_score_grid(self._h, _ffi.ptr(ea), ne, dim, _ffi.ptr(cnt), _ffi.ptr(ids), _ffi.ptr(ta),
                                             nt, dim, dim, dtype, _ffi.HOST, _ffi.ptr(out), out.strides[0] // 4,
                                             _ffi.HOST))
        return out

    # ------------------------------------------------------------------ kernel-level hooks (tests)
    def _test_gemm(self, a, b, ksplit=1):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        m, k = a.shape
        n = b.shape[0]
        out = np.empty((m, n), dtype=np.float32)
        _ffi.check(self._lib.plda_test_gemm(self._h, _ffi.ptr(a), _ffi.ptr(b), m, n, k, int(ksplit), _ffi.ptr(out)))
        return out

    def _test_linalg(self, op, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        d = a.shape[0]
        out = np.empty((d, d))
        out2 = np.empty(d)
        _ffi.check(self._lib.plda_test_linalg(self._h, int(op), _ffi.ptr(a), d, _ffi.ptr(out), _ffi.ptr(out2)))
        return out, out2


def _torch_matrix(x):
    import torch
    if x.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if x.dtype == torch.float32:
        dtype = _ffi.F32
    elif x.dtype == torch.float64:
        dtype = _ffi.F64
    else:
        raise ValueError("Given Input features (argument 1) are not floats! Set the dtype to float!")
    if x.stride(1) != 1:
        x = x.contiguous()
    return x, dtype


def _to_numpy_labels(y):
    if hasattr(y, "detach"):
        return y.detach().cpu().numpy()
    return np.asarray(y)
