"""CPU emulation of the DEVICE arithmetic (split-bf16 x3 operands, fp32 accumulate, the diagonalised
EM, centre-before-split) against the fp64 oracle: shows on the authoring box, without a GPU, that the
precision design meets the parity bar (scores <= 1e-3*max(|s|,1), psi rel << 1e-3).  The GPU tests
check the real kernels; this pins the *design* so a precision regression is caught before GPU time
is spent."""
import numpy as np
import pytest

from oracle import kaldi_plda as kp


def bf16(x):
    f = np.asarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def split(x):
    hi = bf16(x)
    return hi, bf16(np.asarray(x, dtype=np.float64) - hi)


def mm3(a, b):
    """a[M,K] @ b[N,K]^T the way gemm_bf16x3_kernel computes it."""
    ah, al = split(a)
    bh, bl = split(b)
    return (ah @ bh.T + ah @ bl.T + al @ bh.T).astype(np.float32).astype(np.float64)


def device_fit(x, labels, iters):
    lab = labels.astype(np.int64)
    uniq, inv, cnt = np.unique(lab, return_inverse=True, return_counts=True)
    k, d = len(uniq), x.shape[1]
    sums = np.zeros((k, d))
    np.add.at(sums, inv, x)
    means = sums / cnt[:, None]
    xc = (x - means[inv]) / np.sqrt(cnt[inv])[:, None]
    s = mm3(xc.T.copy(), xc.T.copy())
    s = 0.5 * (s + s.T)
    w = 1.0 / cnt
    cw = w.sum()
    mu = (w[:, None] * means).sum(0) / cw
    mc = means - mu
    within, between = np.eye(d), np.eye(d)
    for _ in range(iters):
        a, psi, ainv = kp.joint_diag(within, between)
        u = mm3(mc, a)
        n = cnt[:, None].astype(float)
        r = psi[None, :] / (1 + n * psi[None, :])
        g = n * r
        p = np.sqrt(w)[:, None] * g * u
        q = (1 - g) * u
        bs = mm3(p.T.copy(), p.T.copy())
        ws = mm3(q.T.copy(), q.T.copy())
        bs = 0.5 * (bs + bs.T) + np.diag((w[:, None] * r).sum(0))
        ws = 0.5 * (ws + ws.T) + np.diag(r.sum(0))
        between = ainv @ bs @ ainv.T / cw
        within = (s + ainv @ ws @ ainv.T) / k
        between, within = 0.5 * (between + between.T), 0.5 * (within + within.T)
    a, psi, _ = kp.joint_diag(within, between)
    m = kp.Plda()
    m.mean, m.transform, m.psi = mu, a, psi
    m.compute_derived_vars()
    return m


def device_transform(m, means, counts):
    y = mm3(means - m.mean, m.transform)
    f = np.sqrt(m.dim() / np.sum(y * y / (m.psi[None, :] + 1.0 / counts[:, None]), axis=1))
    return y * f[:, None]


def device_grid(m, e, n, t):
    psi = m.psi
    nn = n[:, None].astype(float)
    a = nn * psi / (nn * psi + 1)
    v = 1 + psi / (nn * psi + 1)
    lmat = e * a / v
    row = (0.5 * np.sum(np.log1p(psi) - np.log(v) - a * a * e * e / v, axis=1)).astype(np.float32)
    g = mm3(lmat, t)
    out = np.empty_like(g)
    for i in range(e.shape[0]):
        vv = 1 + psi / (n[i] * psi + 1)
        q = 0.5 * (1 / (1 + psi) - 1 / vv)
        out[i] = (g[i].astype(np.float32) + row[i] + (t * t @ q).astype(np.float32)).astype(np.float64)
    return out


CASES = {
    "ragged_d40": dict(d=40, counts="ragged", iters=6),
    "config1_rand_500x200": dict(d=200, counts="c1", iters=10),
    "d200_structured": dict(d=200, counts="uniform", iters=10),
}


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_device_arithmetic_meets_parity_bar(name):
    c = CASES[name]
    d = c["d"]
    if c["counts"] == "c1":
        rng = np.random.RandomState(0)
        x = rng.rand(500, 200)
        labels = rng.randint(0, 2, 500).astype("uint")
        xe, xt = rng.rand(120, 200), rng.rand(150, 200)
        le, lt = np.arange(120, dtype="uint"), np.arange(150, dtype="uint")
    else:
        a_b = kp.two_cov_generator(d, 1234)
        rng = np.random.RandomState(3)
        cnt = rng.randint(2, 12, size=120) if c["counts"] == "ragged" else [20] * 150
        x, labels, _ = kp.synth_speakers(a_b, cnt, 1234)
        xe, le, _ = kp.synth_speakers(a_b, rng.randint(1, 5, size=37), 1235)
        xt, lt, _ = kp.synth_speakers(a_b, [1] * 53, 1236)
    ref = kp.MPlda()
    ref.fit(x, labels, c["iters"])
    dev = device_fit(x, labels, c["iters"])
    assert np.max(np.abs(dev.psi - ref.plda.psi) / np.maximum(ref.plda.psi, 1e-12)) < 1e-4
    _, ce, me = kp.group_means(xe, le)
    _, ct, mt = kp.group_means(xt, lt)
    want = kp.score_grid(ref.plda, kp.transform_batch(ref.plda, me, ce), ce, kp.transform_batch(ref.plda, mt, ct))
    got = device_grid(dev, device_transform(dev, me, ce), ce, device_transform(dev, mt, ct))
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 5e-4, err.max()      # bar is 1e-3


def device_grid_f32_rows(m, e, n, t):
    """The resident hot path: fp32 rows in.  The operand producer (score_prep_uniform_kernel, csrc/prep.cu) forms the
    enrol operand in fp32 (e * fp32(a/v), ONE fp32 rounding), splits fp32 values (hi = bf16(x), lo = bf16(x - hi),
    the subtraction is exact), rounds each square to fp32 and accumulates the row / column terms in fp64."""
    f32 = np.float32
    e = np.asarray(e, dtype=f32)
    t = np.asarray(t, dtype=f32)
    psi = m.psi
    out = np.empty((e.shape[0], t.shape[0]))
    t_hi = bf16(t)
    t_lo = bf16(t.astype(np.float64) - t_hi)
    tsq = (t * t).astype(np.float64)                   # fp32 products
    for count in np.unique(n):
        sel = np.nonzero(n == count)[0]
        a = count * psi / (count * psi + 1)
        v = 1 + psi / (count * psi + 1)
        lm = (e[sel] * (a / v).astype(f32)).astype(f32)               # fp32 multiply
        l_hi = bf16(lm)
        l_lo = bf16(lm.astype(np.float64) - l_hi)
        esq = (e[sel] * e[sel]).astype(np.float64)
        row = (0.5 * (np.sum(np.log1p(psi) - np.log(v)) - esq @ (a * a / v))).astype(f32)
        col = (tsq @ (0.5 * (1 / (1 + psi) - 1 / v))).astype(f32)
        g = (l_hi @ t_hi.T + l_hi @ t_lo.T + l_lo @ t_hi.T).astype(f32)
        out[sel] = ((g + col[None, :]) + row[:, None]).astype(f32).astype(np.float64)
    return out


@pytest.mark.parametrize("d", [40, 200])
def test_emulated_fp32_row_producer_meets_parity_bar(d):
    """Scores from fp32 transformed vectors through the fp32 operand path against the fp64 oracle on the SAME fp32
    values: the producer's fp32 rounding is far inside the bar the bf16x3 contraction already sets."""
    a_b = kp.two_cov_generator(d, 1234)
    rng = np.random.RandomState(8)
    x, labels, _ = kp.synth_speakers(a_b, [10] * 120, 1234)
    xe, le, _ = kp.synth_speakers(a_b, [3] * 60, 1235)
    xt, lt, _ = kp.synth_speakers(a_b, [1] * 80, 1236)
    ref = kp.MPlda()
    ref.fit(x, labels, 8)
    _, ce, me = kp.group_means(xe, le)
    _, ct, mt = kp.group_means(xt, lt)
    e32 = kp.transform_batch(ref.plda, me, ce).astype(np.float32)
    t32 = kp.transform_batch(ref.plda, mt, ct).astype(np.float32)
    want = kp.score_grid(ref.plda, e32.astype(np.float64), ce, t32.astype(np.float64))
    got = device_grid_f32_rows(ref.plda, e32, ce, t32)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 5e-4, err.max()
    # and the fp32 path stays close to the fp64-row path of the same design (the bench's parity_spot check)
    alt = device_grid(ref.plda, e32.astype(np.float64), ce, t32.astype(np.float64))
    assert np.max(np.abs(got - alt)) < 2e-4


def test_ragged_column_terms_ride_exactly_inside_the_operands():
    """The ragged-count producer (csrc/prep.cu, score_prep_grouped_vec_kernel) carries the column term c_g of every count
    group in two extra K columns of the test operand, as four bf16 pieces -- hi/lo of the fp32 value, hi/lo of what those
    two left -- against a one-hot pair (1, 1) in the enrol row.  The bf16x3 scheme then delivers hi*hi + hi*lo of both
    columns: that sum must BE the fp32 value (so the grid equals the epilogue-added form), and a row of another group must
    receive exactly zero."""
    rng = np.random.RandomState(7)
    for scale in (1e-6, 1e-2, 1.0, 37.5, 4.0e3, 2.5e6):
        cf = (scale * rng.randn(2000)).astype(np.float32).astype(np.float64)
        h1 = bf16(cf)
        r1 = (cf - h1).astype(np.float32).astype(np.float64)            # exact in fp32
        l1 = bf16(r1)
        r2 = (r1 - l1).astype(np.float32).astype(np.float64)
        h2 = bf16(r2)
        l2 = bf16((r2 - h2).astype(np.float32).astype(np.float64))
        # enrol one-hot pair: A_hi = (1, 1), A_lo = (0, 0); test columns: (h1 | l1) and (h2 | l2) as (B_hi | B_lo)
        acc = (1.0 * h1 + 1.0 * l1 + 0.0 * h1) + (1.0 * h2 + 1.0 * l2 + 0.0 * h2)
        assert np.array_equal(acc.astype(np.float32), cf.astype(np.float32))
        # the pieces are bf16-representable (what the operand planes can hold)
        for piece in (h1, l1, h2, l2):
            assert np.array_equal(bf16(piece), piece)
        # a row of another group multiplies the same columns by (0, 0)
        assert np.all(0.0 * h1 + 0.0 * l1 + 0.0 * h2 + 0.0 * l2 == 0.0)
