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

    def _after_torch(self, *tensors) -> None:
        """Stream contract of the CUDA-tensor entry points: the library launches on the handle's own stream, so it is
        first ordered AFTER torch's current stream on the tensors' device (``plda_stream_wait``: an event, no host
        synchronisation) -- operands still being produced by torch kernels / ``.to(device)`` copies are safe to pass,
        and so are output buffers the caching allocator has just recycled.  Outputs are complete when the call
        returns (own stream) or stream-ordered (a stream installed with ``plda_set_stream``)."""
        import torch
        for t in tensors:
            if t is not None and _is_torch_cuda(t):
                s = torch.cuda.current_stream(t.device)
                _ffi.check(self._lib.plda_stream_wait(self._h, C.c_void_p(s.cuda_stream)))
                return

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
            if _is_torch_cuda(y):
                # resident labels (superset): no 8 B / row upload; must be a non-negative integer tensor
                import torch
                if y.dtype in (torch.float16, torch.float32, torch.float64, torch.bfloat16):
                    raise ValueError("Given labels (argument 2) are not an unsigned! Set the dtype to uint!")
                if y.numel() != n:
                    raise ValueError("labels and features disagree on the number of samples")
                yl = y.reshape(-1).to(torch.int64).contiguous()
                if n and int(yl.min().item()) < 0:
                    raise ValueError("Given labels (argument 2) are not an unsigned! Set the dtype to uint!")
                self._after_torch(xt)
                _ffi.check(self._lib.plda_fit_labels(self._h, C.c_void_p(xt.data_ptr()), n, d, xt.stride(0), dtype,
                                                     _ffi.DEVICE, C.c_void_p(yl.data_ptr()), _ffi.DEVICE, iters))
                return None
            lab = _ffi.as_labels(_to_numpy_labels(y), n)
            self._after_torch(xt)
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

        events = []

        def _cb(_user, count):
            try:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                dist.all_reduce(scratch[:count], group=group)
                e1.record(stream)
                events.append((e0, e1))
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
        # device time spent in the exchanges of this fit: one after the stats pass, one per EM iteration
        self.last_allreduce_ms = float(sum(a.elapsed_time(b) for a, b in events))
        self.last_allreduce_count = len(events)
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
        np.savez(_npz_path(path), mean=mean, transform=tr, psi=psi, z_ids=ids, z_mean=zm, z_std=zs)

    def load(self, path) -> None:
        g = np.load(_npz_path(path))
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
            self._after_torch(xt)
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
            self._after_torch(bt)
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

    def norm_rows(self, vectors, enrol_vectors, numutts=0, seed=0):
        """Array form of ``norm`` for large enrol sets (``plda_norm_rows``): returns ``(mean, std)`` of every enrol row
        against the cohort (fp64 ``[Ne]``; CUDA tensors for CUDA operands, numpy otherwise) without touching the id
        table -- pass them to ``score_grid(..., znorm=(mean, std))`` / ``score_trials`` / ``score_hist``.  Same
        statistics as ``MPlda_norm`` (``src/pldamodule.cpp:196-256``)."""
        if _is_torch_cuda(vectors) or _is_torch_cuda(enrol_vectors):
            import torch
            if not (_is_torch_cuda(vectors) and _is_torch_cuda(enrol_vectors)):
                raise ValueError("norm_rows: background and enrol vectors must both be CUDA tensors or both host arrays")
            bt, bdt = _torch_matrix(vectors)
            et, edt = _torch_matrix(enrol_vectors)
            ne = et.shape[0]
            mean = torch.empty(ne, dtype=torch.float64, device=et.device)
            std = torch.empty(ne, dtype=torch.float64, device=et.device)
            if ne == 0:
                return mean, std
            self._after_torch(bt)
            _ffi.check(self._lib.plda_norm_rows(self._h, C.c_void_p(bt.data_ptr()), bt.shape[0], bt.shape[1],
                                                bt.stride(0), bdt, _ffi.DEVICE, C.c_void_p(et.data_ptr()), ne,
                                                et.stride(0), et.shape[1], edt, _ffi.DEVICE, int(numutts), int(seed),
                                                C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), _ffi.DEVICE))
            return mean, std
        bkg, bdt = _ffi.as_matrix(vectors, "vectors")
        enrol, edt = _ffi.as_matrix(enrol_vectors, "enrol_vectors")
        ne = enrol.shape[0]
        mean = np.empty(ne)
        std = np.empty(ne)
        if ne == 0:
            return mean, std
        _ffi.check(self._lib.plda_norm_rows(self._h, _ffi.ptr(bkg), bkg.shape[0], bkg.shape[1], bkg.shape[1], bdt,
                                            _ffi.HOST, _ffi.ptr(enrol), ne, enrol.shape[1], enrol.shape[1], edt,
                                            _ffi.HOST, int(numutts), int(seed), _ffi.ptr(mean), _ffi.ptr(std), _ffi.HOST))
        return mean, std

    def norm_selection(self, m, numutts, seed=0):
        """Rows of an ``m``-row background set that ``norm(..., numutts, seed)`` uses (``plda_norm_selection``)."""
        m, numutts = int(m), int(numutts)
        rows = np.empty(m if numutts == 0 else numutts, dtype=np.int32)
        _ffi.check(self._lib.plda_norm_selection(m, numutts, int(seed), _ffi.ptr(rows)))
        return rows

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

    def score_grid(self, enrol, enrol_counts, test, enrol_ids=None, out=None, znorm=None):
        """All-pairs LLR grid: ``out[e, t] = LogLikelihoodRatio(enrol[e], enrol_counts[e], test[t])``
        as float32 (Ne, Nt).  Inputs are transformed vectors: numpy (host) or CUDA torch tensors
        (resident; the result is then a CUDA tensor too).  ``enrol_ids``: apply z-norm for ids seen
        by ``norm``; ``znorm=(mean, std)``: z-norm given as per-row arrays (``norm_rows``) instead."""
        cnt = _counts_array(enrol_counts, enrol.shape[0] if hasattr(enrol, "shape") else len(enrol))
        ids = None if enrol_ids is None else np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        if _is_torch_cuda(enrol):
            import torch
            et, tt, dtype = self._cuda_pair(enrol, test, "score_grid")
            ne, dim = et.shape
            nt = tt.shape[0]
            if cnt.shape[0] != ne:
                raise ValueError("score_grid: enrol_counts length mismatch")
            if ids is not None and ids.shape[0] != ne:
                raise ValueError("score_grid: enrol_ids length mismatch")
            if out is None:
                ldo = (nt + 3) // 4 * 4
                buf = torch.empty((ne, ldo), dtype=torch.float32, device=et.device)
                out = buf[:, :nt]
            if ne == 0 or nt == 0:
                return out
            self._after_torch(et)
            if znorm is not None:
                zm, zs, zloc, _keep = _znorm_arrays(znorm, ne, et.device)
                _ffi.check(self._lib.plda_score_grid_z(
                    self._h, C.c_void_p(et.data_ptr()), ne, et.stride(0), _ffi.ptr(cnt), C.c_void_p(tt.data_ptr()), nt,
                    tt.stride(0), dim, dtype, _ffi.DEVICE, C.c_void_p(out.data_ptr()), out.stride(0), _ffi.DEVICE, zm, zs,
                    zloc))
                return out
            _ffi.check(self._lib.plda_score_grid(
                self._h, C.c_void_p(et.data_ptr()), ne, et.stride(0), _ffi.ptr(cnt), _ffi.ptr(ids),
                C.c_void_p(tt.data_ptr()), nt, tt.stride(0), dim, dtype, _ffi.DEVICE, C.c_void_p(out.data_ptr()),
                out.stride(0), _ffi.DEVICE))
            return out
        ea, ta, dtype = _host_pair(enrol, test, "score_grid")
        ne, dim = ea.shape
        nt = ta.shape[0]
        if cnt.shape[0] != ne:
            raise ValueError("score_grid: enrol_counts length mismatch")
        if ids is not None and ids.shape[0] != ne:
            raise ValueError("score_grid: enrol_ids length mismatch")
        if out is None:
            out = np.empty((ne, nt), dtype=np.float32)
        if ne == 0 or nt == 0:
            return out
        if znorm is not None:
            zm, zs, zloc, _keep = _znorm_arrays(znorm, ne, None)
            _ffi.check(self._lib.plda_score_grid_z(self._h, _ffi.ptr(ea), ne, dim, _ffi.ptr(cnt), _ffi.ptr(ta), nt, dim,
                                                   dim, dtype, _ffi.HOST, _ffi.ptr(out), out.strides[0] // 4, _ffi.HOST,
                                                   zm, zs, zloc))
            return out
        _ffi.check(self._lib.plda_score_grid(self._h, _ffi.ptr(ea), ne, dim, _ffi.ptr(cnt), _ffi.ptr(ids), _ffi.ptr(ta),
                                             nt, dim, dim, dtype, _ffi.HOST, _ffi.ptr(out), out.strides[0] // 4,
                                             _ffi.HOST))
        return out

    def _cuda_pair(self, enrol, test, what):
        et, dtype = _torch_matrix(enrol)
        if not _is_torch_cuda(test):
            raise ValueError("%s: enrol and test must both be CUDA tensors or both host arrays" % what)
        tt, dtype2 = _torch_matrix(test)
        if dtype != dtype2:
            raise ValueError("%s: enrol and test dtypes differ" % what)
        if tt.shape[1] != et.shape[1]:
            raise ValueError("%s: enrol and test dimensions differ" % what)
        if tt.device != et.device:
            raise ValueError("%s: enrol and test live on different devices" % what)
        return et, tt, dtype

    def score_trials(self, enrol, enrol_counts, test, trial_enrol, trial_test, enrol_ids=None, znorm=None, mode="auto"):
        """Scores of LISTED trials, computed and gathered on the device (``plda_score_trials``): what
        ``scoring/scorePLDA.py:302-318`` asks of ``MPlda_score`` one trial at a time.  ``trial_enrol[i]`` /
        ``trial_test[i]`` index rows of ``enrol`` / ``test``; returns float32 ``[n_trials]`` (CUDA tensor for CUDA
        operands).  ``mode``: 'direct' (one warp per trial, fp64 accumulation), 'grid' (grid slabs + gather on the
        device) or 'auto' (direct below ~0.5/dim list density)."""
        code = {"auto": 0, "direct": 1, "grid": 2}.get(mode)
        if code is None:
            raise ValueError("mode must be 'auto', 'direct' or 'grid'")
        cnt = _counts_array(enrol_counts, enrol.shape[0] if hasattr(enrol, "shape") else len(enrol))
        ids = None if enrol_ids is None else np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        if _is_torch_cuda(enrol):
            import torch
            et, tt, dtype = self._cuda_pair(enrol, test, "score_trials")
            ne, dim = et.shape
            nt = tt.shape[0]
            if cnt.shape[0] != ne:
                raise ValueError("score_trials: enrol_counts length mismatch")
            if _is_torch_cuda(trial_enrol):
                te = trial_enrol.reshape(-1).to(torch.int32).contiguous()
                tq = trial_test.reshape(-1).to(device=et.device, dtype=torch.int32).contiguous()
                n = te.numel()
                if tq.numel() != n:
                    raise ValueError("score_trials: index arrays differ in length")
                if n and (int(te.min()) < 0 or int(te.max()) >= ne or int(tq.min()) < 0 or int(tq.max()) >= nt):
                    raise ValueError("score_trials: trial index out of range")
                pe, pt, iloc = C.c_void_p(te.data_ptr()), C.c_void_p(tq.data_ptr()), _ffi.DEVICE
            else:
                te = np.ascontiguousarray(trial_enrol, dtype=np.int32).reshape(-1)
                tq = np.ascontiguousarray(trial_test, dtype=np.int32).reshape(-1)
                n = te.shape[0]
                if tq.shape[0] != n:
                    raise ValueError("score_trials: index arrays differ in length")
                pe, pt, iloc = _ffi.ptr(te), _ffi.ptr(tq), _ffi.HOST
            out = torch.empty(n, dtype=torch.float32, device=et.device)
            if n == 0:
                return out
            zm, zs, zloc, _keep = _znorm_arrays(znorm, ne, et.device)
            self._after_torch(et)
            _ffi.check(self._lib.plda_score_trials(
                self._h, C.c_void_p(et.data_ptr()), ne, et.stride(0), _ffi.ptr(cnt), _ffi.ptr(ids),
                C.c_void_p(tt.data_ptr()), nt, tt.stride(0), dim, dtype, _ffi.DEVICE, pe, pt, n, iloc,
                C.c_void_p(out.data_ptr()), _ffi.DEVICE, zm, zs, zloc, code))
            return out
        ea, ta, dtype = _host_pair(enrol, test, "score_trials")
        ne, dim = ea.shape
        nt = ta.shape[0]
        if cnt.shape[0] != ne:
            raise ValueError("score_trials: enrol_counts length mismatch")
        te = np.ascontiguousarray(trial_enrol, dtype=np.int32).reshape(-1)
        tq = np.ascontiguousarray(trial_test, dtype=np.int32).reshape(-1)
        if te.shape[0] != tq.shape[0]:
            raise ValueError("score_trials: index arrays differ in length")
        out = np.empty(te.shape[0], dtype=np.float32)
        if te.shape[0] == 0:
            return out
        zm, zs, zloc, _keep = _znorm_arrays(znorm, ne, None)
        _ffi.check(self._lib.plda_score_trials(self._h, _ffi.ptr(ea), ne, dim, _ffi.ptr(cnt), _ffi.ptr(ids), _ffi.ptr(ta),
                                               nt, dim, dim, dtype, _ffi.HOST, _ffi.ptr(te), _ffi.ptr(tq), te.shape[0],
                                               _ffi.HOST, _ffi.ptr(out), _ffi.HOST, zm, zs, zloc, code))
        return out

    def score_hist(self, enrol, enrol_count, test, enrol_spk, test_spk, lo, hi, nbins=1 << 16, theta_lo=-np.inf,
                   znorm=None):
        """Histogram sink (``plda_score_hist``): target / non-target score histograms of the whole ``Ne x Nt`` grid
        without materialising it.  Trial ``(e, t)`` is a target iff ``enrol_spk[e] == test_spk[t]``.  Non-targets
        scoring below ``theta_lo`` are only counted.  Returns ``(hist_target, hist_nontarget, below)`` (uint64
        numpy arrays of ``nbins`` and an int); ``plda_b200.eer.eer_from_hist`` turns them into the EER.  CUDA tensors
        only (resident operands), one enrol count for all rows."""
        import torch
        et, tt, dtype = self._cuda_pair(enrol, test, "score_hist")
        ne, dim = et.shape
        nt = tt.shape[0]
        es = torch.as_tensor(enrol_spk).reshape(-1).to(device=et.device, dtype=torch.int32).contiguous()
        ts = torch.as_tensor(test_spk).reshape(-1).to(device=et.device, dtype=torch.int32).contiguous()
        if es.numel() != ne or ts.numel() != nt:
            raise ValueError("score_hist: speaker id arrays must match the enrol / test rows")
        nbins = int(nbins)
        ht = np.zeros(nbins, dtype=np.uint64)
        hn = np.zeros(nbins, dtype=np.uint64)
        below = np.zeros(1, dtype=np.uint64)
        zm, zs, zloc, _keep = _znorm_arrays(znorm, ne, et.device)
        th = float(theta_lo)
        if not np.isfinite(th):
            th = -3.0e38
        self._after_torch(et)
        _ffi.check(self._lib.plda_score_hist(
            self._h, C.c_void_p(et.data_ptr()), ne, et.stride(0), int(enrol_count), C.c_void_p(tt.data_ptr()), nt,
            tt.stride(0), dim, dtype, _ffi.DEVICE, C.c_void_p(es.data_ptr()), C.c_void_p(ts.data_ptr()), _ffi.DEVICE,
            float(lo), float(hi), nbins, th, zm, zs, zloc, _ffi.ptr(ht), _ffi.ptr(hn), _ffi.ptr(below), _ffi.HOST))
        return ht, hn, int(below[0])

    # ------------------------------------------------------------------ kernel-level hooks (tests)
    def _test_gemm(self, a, b, ksplit=1):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        m, k = a.shape
        n = b.shape[0]
        out = np.empty((m, n), dtype=np.float32)
        _ffi.check(self._lib.plda_test_gemm(self._h, _ffi.ptr(a), _ffi.ptr(b), m, n, k, int(ksplit), _ffi.ptr(out)))
        return out

    def _test_scatter(self, x, labels, scale_by_count=True):
        xa, dtype = _ffi.as_matrix(x, "features")
        lab = _ffi.as_labels(labels, xa.shape[0])
        n, d = xa.shape
        k_max = int(np.unique(lab).shape[0])
        sc = np.empty((d, d))
        means = np.empty((k_max, d))
        k = C.c_int64()
        _ffi.check(self._lib.plda_test_scatter(self._h, _ffi.ptr(xa), n, d, dtype, _ffi.ptr(lab),
                                               1 if scale_by_count else 0, _ffi.ptr(sc), _ffi.ptr(means), means.size,
                                               C.byref(k)))
        return sc, means[:k.value]

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


def _npz_path(path):
    """``np.savez`` appends '.npz' to a bare name; ``save`` and ``load`` agree on the final path."""
    if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__"):
        import os
        p = os.fspath(path)
        if isinstance(p, bytes):
            p = p.decode()
        return p if p.endswith(".npz") else p + ".npz"
    return path         # an open file object


def _counts_array(enrol_counts, ne):
    if np.isscalar(enrol_counts):
        return np.full(int(ne), int(enrol_counts), dtype=np.int32)
    if hasattr(enrol_counts, "detach"):
        enrol_counts = enrol_counts.detach().cpu().numpy()
    return np.ascontiguousarray(enrol_counts, dtype=np.int32).reshape(-1)


def _host_pair(enrol, test, what):
    ea, dtype = _ffi.as_matrix(enrol, "enrol")
    ta, dtype2 = _ffi.as_matrix(test, "test")
    if dtype != dtype2:
        ea = ea.astype(np.float64)
        ta = ta.astype(np.float64)
        dtype = _ffi.F64
    if ta.shape[1] != ea.shape[1]:
        raise ValueError("%s: enrol and test dimensions differ" % what)
    return ea, ta, dtype


def _znorm_arrays(znorm, ne, device):
    """(mean, std) -> (void* mean, void* std, loc, keep-alive).  CUDA fp64 tensors stay on the device."""
    if znorm is None:
        return C.c_void_p(None), C.c_void_p(None), _ffi.HOST, None
    mean, std = znorm
    if _is_torch_cuda(mean) and _is_torch_cuda(std):
        import torch
        m = mean.reshape(-1).to(torch.float64).contiguous()
        s = std.reshape(-1).to(torch.float64).contiguous()
        if m.numel() != ne or s.numel() != ne:
            raise ValueError("znorm arrays must have one entry per enrol row")
        return C.c_void_p(m.data_ptr()), C.c_void_p(s.data_ptr()), _ffi.DEVICE, (m, s)
    if hasattr(mean, "detach"):
        mean, std = mean.detach().cpu().numpy(), std.detach().cpu().numpy()
    m = np.ascontiguousarray(mean, dtype=np.float64).reshape(-1)
    s = np.ascontiguousarray(std, dtype=np.float64).reshape(-1)
    if m.shape[0] != ne or s.shape[0] != ne:
        raise ValueError("znorm arrays must have one entry per enrol row")
    return _ffi.ptr(m), _ffi.ptr(s), _ffi.HOST, (m, s)


def _to_numpy_labels(y):
    if hasattr(y, "detach"):
        return y.detach().cpu().numpy()
    return np.asarray(y)
