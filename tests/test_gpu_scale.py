"""GPU parity at BASELINE.json sizes and shapes: full C2 grid (10k x 10k, d = 200) against the oracle's
Gram-form grid (every score), size-independent properties (n = 1 symmetry, enrol-block additivity), the C3
shape (d = 256, targetdim = 150, z-norm) and a ragged-speaker fit."""
import numpy as np
import pytest

from oracle import kaldi_plda as kp

pytestmark = pytest.mark.gpu




def tol_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    return np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)


@pytest.fixture(scope="module")
def model200():
    """A d = 200 model fitted by the ORACLE on 40 k rows (400 speakers x 100), installed on the GPU with set_model
    so that the scoring path is tested in isolation from the fit."""
    d = 200
    a_b = kp.two_cov_generator(d, seed=1234)
    x, labels, _ = kp.synth_speakers(a_b, [100] * 400, seed=1234)
    ref = kp.MPlda()
    ref.fit(x, labels, 5)
    return ref, a_b


def test_c2_full_grid_every_score(model200):
    from plda_b200 import PLDA
    ref, a_b = model200
    d, ne, nt = 200, 10_000, 10_000
    xe, le, z = kp.synth_speakers(a_b, [3] * ne, seed=1235)
    rng = np.random.RandomState(1236)
    xt = 0.5 + z @ a_b.T + rng.randn(nt, d)                     # test utterance t belongs to enrol speaker t
    g = PLDA()
    g.set_model(ref.plda.mean, ref.plda.transform, ref.plda.psi)
    means = xe.reshape(ne, 3, d).mean(axis=1)
    e_g = g.transform_batch(means, counts=3)
    t_g = g.transform_batch(xt, counts=1)
    e_r = kp.transform_batch(ref.plda, means, np.full(ne, 3))
    t_r = kp.transform_batch(ref.plda, xt, np.full(nt, 1))
    assert np.max(np.abs(e_g - e_r)) <= 2e-4 * np.max(np.abs(e_r))
    grid = g.score_grid(e_g, np.full(ne, 3, dtype=np.int32), t_g)
    want = kp.score_grid(ref.plda, e_r, np.full(ne, 3), t_r)
    err = tol_err(grid, want)
    assert err.max() <= 1e-3, err.max()
    # EER on 10 k target / ~1e8 non-target trials: identical to +-0.01 % absolute
    mask = np.eye(ne, dtype=bool)
    sub = np.random.RandomState(0).rand(ne, nt) < 0.02            # 2 % of the non-targets keeps the sort cheap
    sub &= ~mask
    eer_g = kp.eer_percent(grid[mask].astype(np.float64), grid[sub].astype(np.float64))
    eer_r = kp.eer_percent(want[mask], want[sub])
    assert abs(eer_g - eer_r) <= 0.01


def test_n1_symmetry_and_block_additivity(model200):
    """Size-independent properties: LLR(e, 1, t) is symmetric in (e, t); scoring enrol blocks separately gives
    the same slab as the full grid (the multi-GPU sharding invariant)."""
    from plda_b200 import PLDA
    ref, a_b = model200
    g = PLDA()
    g.set_model(ref.plda.mean, ref.plda.transform, ref.plda.psi)
    x, _, _ = kp.synth_speakers(a_b, [1] * 3000, seed=77)
    v = g.transform_batch(x, counts=1)
    ones = np.ones(3000, dtype=np.int32)
    s = g.score_grid(v, ones, v).astype(np.float64)
    assert np.max(np.abs(s - s.T) / np.maximum(np.abs(s), 1.0)) <= 1e-3
    top = g.score_grid(v[:1234], ones[:1234], v)
    bot = g.score_grid(v[1234:], ones[1234:], v)
    assert np.array_equal(np.vstack([top, bot]), s.astype(np.float32))


def test_c3_shape_targetdim_znorm():
    """BASELINE configs[2] shape at test size: d = 256, fit, targetdim = 150 (leading directions, length norm with
    dim = 150), z-norm against a held-out cohort, grid vs the oracle with the same defined semantics."""
    from plda_b200 import PLDA
    d, r = 256, 150
    a_b = kp.two_cov_generator(d, seed=1234)
    x, labels, _ = kp.synth_speakers(a_b, [20] * 300, seed=1234)
    ref = kp.MPlda()
    ref.fit(x, labels, 4)
    g = PLDA()
    g.fit(x, labels, 4)
    _, _, psi = g.get_model()
    assert np.allclose(psi, ref.plda.psi, rtol=1e-3, atol=1e-6)
    xe, le, _ = kp.synth_speakers(a_b, [3] * 500, seed=1235)
    xt, _, _ = kp.synth_speakers(a_b, [1] * 700, seed=1236)
    bkg, _, _ = kp.synth_speakers(a_b, [1] * 400, seed=1237)
    means = xe.reshape(500, 3, d).mean(axis=1)

    def ref_transform(rows, n):
        y = (rows - ref.plda.mean) @ ref.plda.transform[:r].T
        f = np.sqrt(r / np.sum(y * y / (ref.plda.psi[:r] + 1.0 / n), axis=1))
        return y * f[:, None]

    pr = kp.Plda()
    pr.mean, pr.transform, pr.psi = ref.plda.mean[:r], np.eye(r), ref.plda.psi[:r]
    pr.compute_derived_vars()
    e_r, t_r, b_r = ref_transform(means, 3), ref_transform(xt, 1), ref_transform(bkg, 400)
    e_g = g.transform_batch(means, counts=3, targetdim=r)
    t_g = g.transform_batch(xt, counts=1, targetdim=r)
    assert e_g.shape == (500, r)
    raw = kp.score_grid(pr, e_r, np.full(500, 3), t_r)
    cohort = kp.score_grid(pr, b_r, np.ones(400, dtype=np.int64), e_r)          # (M, Ne), LLR(bkg, 1, enrol)
    zm, zs = cohort.mean(axis=0), cohort.std(axis=0)
    want = (raw - zm[:, None]) / zs[:, None]
    ids = np.arange(500, dtype=np.uint64)
    # norm() takes the dict API: {id: (n, vec)}; the cohort is raw, transformed inside with num_examples = 400
    # and (defined semantics) the enrol dimension
    g.norm(bkg, {int(i): (3, e_g[i]) for i in ids})
    got = g.score_grid(e_g, np.full(500, 3, dtype=np.int32), t_g, enrol_ids=ids)
    assert tol_err(got, want).max() <= 1e-3


def test_ragged_speakers_fit():
    """Log-normal speaker sizes (many distinct counts): the diagonalised EM does not depend on the number of
    distinct counts; psi and the covariances match the oracle's per-class loop."""
    from plda_b200 import PLDA
    d = 64
    a_b = kp.two_cov_generator(d, seed=1234)
    rng = np.random.RandomState(9)
    counts = np.maximum(1, np.round(np.exp(rng.randn(400) * 0.8 + 2.0))).astype(int)
    assert len(np.unique(counts)) > 20
    x, labels, _ = kp.synth_speakers(a_b, counts, seed=1234)
    ref = kp.MPlda()
    ref.fit(x, labels, 6)
    g = PLDA()
    g.fit(x, labels, 6)
    _, _, psi = g.get_model()
    w, b = g.get_covariances()
    assert np.allclose(psi, ref.plda.psi, rtol=1e-3, atol=1e-6)
    assert np.allclose(w, ref.estimator.within_var, rtol=1e-3, atol=1e-3 * np.abs(ref.estimator.within_var).max())
    assert np.allclose(b, ref.estimator.between_var, rtol=1e-3, atol=1e-3 * np.abs(ref.estimator.between_var).max())


def test_lda_c5_shape_scaled_down():
    """BASELINE configs[4] shape, scaled: 100k x 200, 500 classes, 20k test rows, log-probas vs the numpy port."""
    from oracle.lda_port import LDAOracle
    from plda_b200 import LDA
    rng = np.random.RandomState(3)
    k, d, n, nt = 500, 200, 100_000, 20_000
    centers = rng.randn(k, d) * 0.7
    y = np.arange(n) % k
    x = centers[y] + rng.randn(n, d)
    yt = rng.randint(0, k, nt)
    t = centers[yt] + rng.randn(nt, d)
    m = LDA()
    m.fit(x, y)
    o = LDAOracle("svd")
    o.fit(x, y)
    lp = np.asarray(m.predict_log_proba(t), dtype=np.float64)
    ref = o.predict_log_proba(t)
    assert np.max(np.abs(lp - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-3
    assert (lp.argmax(1) == ref.argmax(1)).mean() > 0.9999


def test_grid_is_bitwise_repeatable(model200):
    """Regression: the epilogue's shared column-term cache used to be rewritten while a slow sibling warp was still
    reading it (32 x 32 blocks of wrong scores in ~1 of 5 launches).  Same inputs -> the same bits, every launch."""
    import torch
    from plda_b200 import PLDA
    ref, a_b = model200
    g = PLDA()
    g.set_model(ref.plda.mean, ref.plda.transform, ref.plda.psi)
    x, _, _ = kp.synth_speakers(a_b, [1] * 16000, seed=99)
    v = torch.as_tensor(g.transform_batch(x, counts=1), device="cuda", dtype=torch.float32)
    e, t = v[:8000], v[8000:]
    n = np.full(8000, 3, dtype=np.int32)
    first = g.score_grid(e, n, t).clone()
    for _ in range(60):
        again = g.score_grid(e, n, t)
        assert torch.equal(again, first)


def test_c2_full_size_fit_vs_oracle():
    """BASELINE configs[1] at FULL size: fit 100k x 200, 1k speakers, 10 EM iterations on the device (fused stats pass
    + diagonalised EM) against the oracle's Kaldi loop on the same rows: psi, within and between covariances."""
    from plda_b200 import PLDA
    d = 200
    a_b = kp.two_cov_generator(d, seed=1234)
    x, labels, _ = kp.synth_speakers(a_b, [100] * 1000, seed=1234)
    ref = kp.MPlda()
    ref.fit(x, labels, 10)
    g = PLDA()
    g.fit(x, labels, 10)
    mean, a, psi = g.get_model()
    w, b = g.get_covariances()
    assert np.allclose(mean, ref.plda.mean, rtol=1e-6, atol=1e-7)
    assert np.allclose(psi, ref.plda.psi, rtol=1e-3, atol=1e-6)
    est = ref.estimator
    assert np.max(np.abs(w - est.within_var)) <= 1e-3 * np.abs(est.within_var).max()
    assert np.max(np.abs(b - est.between_var)) <= 1e-3 * np.abs(est.between_var).max()
    assert np.allclose(a @ w @ a.T, np.eye(d), atol=1e-8)
    # fp32 rows (the resident bench path) give the same model
    import torch
    g32 = PLDA()
    g32.fit(torch.from_numpy(x).to("cuda", dtype=torch.float32), torch.from_numpy(labels.astype(np.int64)).to("cuda"), 10)
    _, _, psi32 = g32.get_model()
    assert np.allclose(psi32, ref.plda.psi, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("targetdim", [0, 150])
def test_d512_fit_and_random_tiles(targetdim):
    """BASELINE configs[3]'s dimension (d = 512): device fit vs oracle fit, then >= 16 random 1024 x 1024 tiles of an
    8192 x 8192 grid (SURVEY 8d parity reporting) -- full dimension and targetdim = 150 -- and the EER of the sampled
    tiles, each side scoring with ITS OWN fitted model (scores are invariant to eigenvector signs)."""
    from plda_b200 import PLDA
    d = 512
    a_b = kp.two_cov_generator(d, seed=1234)
    x, labels, _ = kp.synth_speakers(a_b, [40] * 600, seed=1234)
    ref = kp.MPlda()
    ref.fit(x, labels, 3)
    g = PLDA()
    g.fit(x, labels, 3)
    _, _, psi = g.get_model()
    w, b = g.get_covariances()
    assert np.allclose(psi, ref.plda.psi, rtol=1e-3, atol=1e-6)
    assert np.max(np.abs(w - ref.estimator.within_var)) <= 1e-3 * np.abs(ref.estimator.within_var).max()
    assert np.max(np.abs(b - ref.estimator.between_var)) <= 1e-3 * np.abs(ref.estimator.between_var).max()

    n_side, tile = 8192, 1024
    xe, _, z = kp.synth_speakers(a_b, [3] * n_side, seed=1235)
    rng = np.random.RandomState(1236)
    xt = 0.5 + z @ a_b.T + rng.randn(n_side, d)              # test t belongs to enrol speaker t
    means = xe.reshape(n_side, 3, d).mean(axis=1)
    r = targetdim if targetdim else d

    def ref_transform(rows, n):
        y = (rows - ref.plda.mean) @ ref.plda.transform[:r].T
        f = np.sqrt(r / np.sum(y * y / (ref.plda.psi[:r] + 1.0 / n), axis=1))
        return y * f[:, None]

    pr = kp.Plda()
    pr.mean, pr.transform, pr.psi = ref.plda.mean[:r], np.eye(r), ref.plda.psi[:r]
    pr.compute_derived_vars()
    e_r, t_r = ref_transform(means, 3), ref_transform(xt, 1)
    import torch
    e_g = g.transform_batch(torch.from_numpy(means).cuda(), counts=3, targetdim=targetdim, out_dtype=np.float32)
    t_g = g.transform_batch(torch.from_numpy(xt).cuda(), counts=1, targetdim=targetdim, out_dtype=np.float32)
    grid = g.score_grid(e_g, 3, t_g)                          # 8192 x 8192 on the device
    pick = np.random.RandomState(7)
    tiles = [(0, 0), (7, 7), (3, 3), (5, 5)] + [tuple(pick.randint(0, n_side // tile, 2)) for _ in range(12)]
    worst, tar_g, non_g, tar_r, non_r = 0.0, [], [], [], []
    for (bi, bj) in tiles:
        rs, cs = slice(bi * tile, (bi + 1) * tile), slice(bj * tile, (bj + 1) * tile)
        got = grid[rs, cs].cpu().numpy().astype(np.float64)
        want = kp.score_grid(pr, e_r[rs], np.full(tile, 3), t_r[cs])
        worst = max(worst, float(tol_err(got, want).max()))
        if bi == bj:
            m = np.eye(tile, dtype=bool)
            tar_g.append(got[m]); tar_r.append(want[m]); non_g.append(got[~m]); non_r.append(want[~m])
        else:
            non_g.append(got.ravel()); non_r.append(want.ravel())
    assert worst <= 1e-3, worst
    eer_g = kp.eer_percent(np.concatenate(tar_g), np.concatenate(non_g))
    eer_r = kp.eer_percent(np.concatenate(tar_r), np.concatenate(non_r))
    assert abs(eer_g - eer_r) <= 0.01, (eer_g, eer_r)
