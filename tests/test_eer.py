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
