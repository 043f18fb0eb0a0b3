"""EER restatement (plda_b200/eer.py) against the oracle's definition (scoring/eer.py:68-73) -- CPU tensors here,
the same code runs on CUDA grids (tests/test_gpu_scale.py uses the oracle EER on device-produced scores)."""
import numpy as np
import pytest
import torch

from oracle import kaldi_plda as kp
from plda_b200.eer import eer_percent


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_eer_matches_oracle_definition(seed):
    rng = np.random.RandomState(seed)
    ne = nt = 120
    scores = rng.randn(ne, nt) * 3.0
    scores[np.arange(ne), np.arange(nt)] += 4.0
    scores = np.round(scores, 2)                  # ties between targets and non-targets
    mask = np.eye(ne, dtype=bool)
    want = kp.eer_percent(scores[mask], scores[~mask])
    got = eer_percent(torch.from_numpy(scores), target_mask=torch.from_numpy(mask))
    assert got == pytest.approx(want, abs=1e-9)
    got2 = eer_percent(torch.from_numpy(scores), enrol_labels=np.arange(ne), test_labels=np.arange(nt))
    assert got2 == pytest.approx(want, abs=1e-9)


def test_eer_rejects_degenerate_input():
    with pytest.raises(ValueError):
        eer_percent(torch.zeros(3, 3), target_mask=torch.zeros(3, 3, dtype=torch.bool))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_eer_from_hist_matches_exact(seed):
    """Histogram-sink EER (exact target scores + tail histogram of the non-targets) against the exact EER."""
    from plda_b200.eer import eer_from_hist
    rng = np.random.RandomState(seed)
    tar = rng.randn(2000) * 4.0 + 12.0
    non = rng.randn(400000) * 5.0 - 8.0
    want = kp.eer_percent(tar, non)
    theta = float(np.quantile(tar, 0.002))      # FRR(theta) = 0.2 % < EER: the crossing lies above theta
    hi = float(max(tar.max(), non.max())) + 1.0
    nbins = 1 << 16
    keep = non[non >= theta]
    b = np.clip(np.floor((keep - theta) * nbins / (hi - theta)), 0, nbins - 1).astype(np.int64)
    hist = np.bincount(b, minlength=nbins).astype(np.uint64)
    got, valid = eer_from_hist(tar, hist, int((non < theta).sum()), theta, hi)
    assert valid
    assert abs(got - want) <= 0.01
    # a tail that starts above the crossing is reported as invalid
    theta_bad = float(np.quantile(tar, 0.5))
    keep = non[non >= theta_bad]
    b = np.clip(np.floor((keep - theta_bad) * nbins / (hi - theta_bad)), 0, nbins - 1).astype(np.int64)
    _, valid = eer_from_hist(tar, np.bincount(b, minlength=nbins), int((non < theta_bad).sum()), theta_bad, hi)
    assert not valid
