"""d-vector pooling on the device -- host-side mirror of ``scoring/extractdvector.py:19-58,160-168``.

The reference turns the frame-level activations of an utterance into one d-vector by L2-normalising every frame
(``getnormalizedvector``) and taking the mean / max / variance over the frames, one utterance at a time in Python
(``extractvectors``).  Here the frames of all utterances go through ONE call of ``plda_dvector_pool``
(``csrc/dvector.cu``: a segmented reduction that reads every frame from HBM once).  The function names and return
shapes of the reference are kept for single utterances; ``pool_dvectors`` is the batched entry point.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

_MODES = {"mean": 0, "max": 1, "var": 2}
_handles = {}


def _handle(device: int):
    """The pooling kernel needs no model: one lazily created handle per device provides the stream."""
    h = _handles.get(device)
    if h is None:
        from .plda import PLDA
        h = _handles[device] = PLDA(device=device)
    return h


def pool_dvectors(frames, offsets, method: str = "mean", l2norm: bool = True, device: int = 0):
    """Pool utterance ``u`` = ``frames[offsets[u]:offsets[u+1]]`` into row ``u`` of the result (float64
    ``[len(offsets)-1, d]``).  ``frames``: float32/float64 numpy array or CUDA tensor ``[n_frames, d]``;
    ``method``: ``'mean'`` / ``'max'`` / ``'var'`` (``extractdvector.py:32-46``); ``l2norm=False`` gives the
    ``*_nol2`` variants (``:49-58``).  A CUDA tensor in gives a CUDA tensor out."""
    if method not in _MODES:
        raise ValueError("method must be 'mean', 'max' or 'var'")
    off = np.ascontiguousarray(offsets, dtype=np.int64).reshape(-1)
    if off.shape[0] < 1 or (off.shape[0] > 1 and np.any(np.diff(off) <= 0)):
        raise ValueError("offsets must be strictly increasing (every utterance needs at least one frame)")
    n_utts = off.shape[0] - 1
    lib = _ffi.lib()
    is_cuda = hasattr(frames, "is_cuda") and frames.is_cuda
    if is_cuda:
        import torch
        from .plda import _torch_matrix
        ft, dtype = _torch_matrix(frames)
        n, d = ft.shape
        if n_utts and (off[0] < 0 or off[-1] > n):
            raise ValueError("offsets out of range")
        out = torch.empty((n_utts, d), dtype=torch.float64, device=ft.device)
        h = _handle(ft.device.index or 0)
        torch.cuda.current_stream(ft.device).synchronize()      # the handle launches on its own stream
        _ffi.check(lib.plda_dvector_pool(h._h, C.c_void_p(ft.data_ptr()), n, d, ft.stride(0) if n else d, dtype,
                                         _ffi.DEVICE, _ffi.ptr(off), n_utts, _MODES[method], 1 if l2norm else 0,
                                         C.c_void_p(out.data_ptr()), d, _ffi.DEVICE))
        return out
    fa, dtype = _ffi.as_matrix(frames, "frames")
    n, d = fa.shape
    if n_utts and (off[0] < 0 or off[-1] > n):
        raise ValueError("offsets out of range")
    out = np.empty((n_utts, d), dtype=np.float64)
    _ffi.check(lib.plda_dvector_pool(_handle(device)._h, _ffi.ptr(fa), n, d, d, dtype, _ffi.HOST, _ffi.ptr(off),
                                     n_utts, _MODES[method], 1 if l2norm else 0, _ffi.ptr(out), d, _ffi.HOST))
    return out


def _single(utt, method, l2norm):
    utt = np.asarray(utt)
    if utt.ndim != 2:
        raise ValueError("an utterance is a 2-D array (n_frames, featdim)")
    if utt.dtype.kind != "f":
        utt = utt.astype(np.float64)
    return pool_dvectors(utt, [0, utt.shape[0]], method, l2norm)[0]


# ---- the reference's per-utterance functions (same names, same return shapes) --------------------------------
def extractdvectormax(utt):
    return _single(utt, "max", True)


def extractdvectormean(utt):
    return _single(utt, "mean", True)


def extractdvectorvar(utt):
    return _single(utt, "var", True)


def extractdvectormean_nol2(uttvec):
    return _single(uttvec, "mean", False)[np.newaxis, :]      # the reference returns shape (1, d), :50


def extractdvectorvar_nol2(uttvec):
    return _single(uttvec, "var", False)[np.newaxis, :]


def extractdvectormax_nol2(uttvec):
    return _single(uttvec, "max", False)[np.newaxis, :]


def extractvectors(datadict, method: str = "mean", l2norm: bool = True):
    """``extractvectors`` (``extractdvector.py:160-168``): ``{speaker: [utterance arrays]}`` -> ``(dvectors, labels)``,
    all utterances pooled by one device call instead of a Python loop."""
    utts, labels = [], []
    for spk, v in datadict.items():
        for u in v:
            utts.append(np.asarray(u))
            labels.append(spk)
    if not utts:
        return np.empty((0, 0)), np.array(labels)
    lens = np.array([u.shape[0] for u in utts], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    frames = np.concatenate(utts, axis=0)
    if frames.dtype.kind != "f":
        frames = frames.astype(np.float64)
    return pool_dvectors(frames, offsets, method, l2norm), np.array(labels)
