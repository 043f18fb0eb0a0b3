"""Equal error rate on the device -- the metric the north star is checked with ("EER identical").

Definition of the reference's ``scoring/eer.py:68-73`` (``bob.measure.eer_threshold`` + ``farfrr``): the threshold
minimising ``|FAR - FRR|`` over all observed scores (accept iff score >= threshold), reported as
``(FAR + FRR) / 2 * 100``.  ``bob`` is not installable here; this is an exact restatement for score grids that
live in HBM (C3 / C4 grids are 20-100 GB and never leave the device): one ``torch.sort`` of the scores plus
prefix counts.  It is a metric utility, not part of the timed hot path, so it uses torch's library sort.
"""
from __future__ import annotations


def eer_percent(scores, target_mask=None, enrol_labels=None, test_labels=None, max_nontargets=None, seed=0):
    """EER (%) of a CUDA score grid ``scores[e, t]``.

    Targets are given either by a boolean ``target_mask`` of the same shape or by integer label vectors
    (``enrol_labels[e] == test_labels[t]``).  ``max_nontargets``: evaluate on a seeded random subset of the
    non-target trials (the full 1e10+ grids of C3/C4 do not need every non-target for +-0.01 %)."""
    import torch
    if target_mask is None:
        el = torch.as_tensor(enrol_labels, device=scores.device).view(-1, 1)
        tl = torch.as_tensor(test_labels, device=scores.device).view(1, -1)
        target_mask = el == tl
    tar = scores[target_mask].double()
    non = scores[~target_mask]
    if max_nontargets is not None and non.numel() > max_nontargets:
        g = torch.Generator(device=scores.device)
        g.manual_seed(seed)
        idx = torch.randint(0, non.numel(), (max_nontargets,), device=scores.device, generator=g)
        non = non[idx]
    non = non.double()
    n_tar, n_non = tar.numel(), non.numel()
    if n_tar == 0 or n_non == 0:
        raise ValueError("EER needs at least one target and one non-target trial")
    tar_s, _ = torch.sort(tar)
    non_s, _ = torch.sort(non)
    thr = torch.unique(torch.cat([tar_s, non_s]))               # sorted candidate thresholds
    far = 1.0 - torch.searchsorted(non_s, thr, right=False).double() / n_non    # non-targets with score >= thr
    frr = torch.searchsorted(tar_s, thr, right=False).double() / n_tar          # targets with score <  thr
    i = torch.argmin(torch.abs(far - frr))
    return float((far[i] + frr[i]) / 2.0 * 100.0)


def eer_from_hist(target_scores, hist_nontarget, below, lo, hi):
    """EER (%) from the histogram sink (``PLDA.score_hist`` / ``plda_score_hist``) -- for grids that are never
    materialised (BASELINE configs[3]: 2e11 trials).

    ``target_scores``: the EXACT target scores (few: one per test utterance; ``PLDA.score_trials`` on the target
    pairs), so FRR is exact.  ``hist_nontarget`` (uint64 ``[nbins]`` over ``[lo, hi)``, last bin closed above) holds the
    non-targets that scored at least ``theta_lo``; ``below`` counts the rest.  Candidate thresholds are the bin edges at
    or above the lowest populated non-target bin: FAR is exact at a bin edge, FRR exact everywhere, so the result is
    the reference's EER (``scoring/eer.py:68-73``: threshold minimising |FAR - FRR|, report their mean) up to the
    restriction of the threshold to bin edges (one non-target bin of FAR, at most one target step of FRR).

    Returns ``(eer_percent, valid)``; ``valid`` is False when the crossing lies below the histogrammed tail
    (FAR at the lowest usable edge is already below FRR there) -- re-run the sink with a lower ``theta_lo``.  A safe
    choice is a quantile q of the target scores with q below the expected EER (FRR(theta_lo) = q <= EER puts the
    crossing at or above theta_lo); ``theta_lo = -inf`` always works (every non-target is then binned)."""
    import numpy as np
    hn = np.asarray(hist_nontarget, dtype=np.float64)
    nbins = hn.shape[0]
    tar = np.sort(np.asarray(target_scores, dtype=np.float64).reshape(-1))
    n_non = float(hn.sum() + float(below))
    if tar.shape[0] == 0 or n_non == 0:
        raise ValueError("EER needs at least one target and one non-target trial")
    edges = lo + (hi - lo) * np.arange(nbins, dtype=np.float64) / nbins          # lower edge of every bin
    above = np.cumsum(hn[::-1])[::-1]                                           # non-targets with bin >= b
    nz = np.nonzero(hn)[0]
    first = int(nz[0]) + 1 if nz.size else nbins - 1        # edges above the bin that may hold clamped / cut scores
    first = min(first, nbins - 1)
    far = above[first:] / n_non
    frr = np.searchsorted(tar, edges[first:], side="left") / tar.shape[0]
    i = int(np.argmin(np.abs(far - frr)))
    valid = bool(far[0] >= frr[0])
    return float((far[i] + frr[i]) / 2.0 * 100.0), valid
