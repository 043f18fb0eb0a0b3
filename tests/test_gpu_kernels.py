"""GPU parity tests of the kernel-level pieces, through the C ABI (plda_test_gemm /
plda_test_linalg): tensor-core GEMM vs numpy, Cholesky / triangular inverse / Jacobi eig vs numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plda():
    from plda_b200 import PLDA
    return PLDA()


def _bf16_round(x):
    """round-to-nearest-even to bf16, returned as float64"""
    f = np.asarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def _split(x):
    hi = _bf16_round(x)
    lo = _bf16_round(np.asarray(x, dtype=np.float64) - hi)
    return hi, lo


def emulate_bf16x3(a, b):
    ah, al = _split(a)
    bh, bl = _split(b)
    return ah @ bh.T + ah @ bl.T + al @ bh.T


SHAPES = [
    (128, 256, 64),     # exactly one tile, one k-block
    (128, 256, 128),    # two k-blocks (pipeline wrap)
    (128, 256, 16),     # single UMMA_K step
    (100, 70, 40),      # ragged everything
    (500, 500, 200),    # BASELINE config 1 grid, d=200 -> K16=208 (partial last k-block)
    (1000, 777, 208),
    (300, 200, 200),    # N < 256 -> BN=208, single n tile (transform / EM shapes)
    (257, 513, 512),    # tile edges + 8 k-blocks
    (5, 3, 7),          # tiny
    (2048, 4096, 256),  # many tiles per CTA (accumulator double-buffer wrap)
    (260, 300, 150),    # K16 = 160 = 64 + 64 + 32 -> 64-byte-swizzle tail box (C3 targetdim=150)
    (130, 258, 80),     # 64 + 16 -> 32-byte-swizzle tail box
    (64, 64, 96),       # 64 + 32
    (1000, 1000, 600),  # 9 full k-blocks + 32 tail
    (129, 33, 1),       # K = 1
]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_bf16x3_matches_numpy(plda, m, n, k):
    rng = np.random.RandomState(m * 7 + n * 3 + k)
    a = rng.randn(m, k)
    b = rng.randn(n, k)
    out = plda._test_gemm(a, b, 1).astype(np.float64)
    assert not np.isnan(out).any(), "NaN canary survived: some output elements were never written"
    exact = a @ b.T
    emu = emulate_bf16x3(a, b)
    scale = np.sqrt(k)
    # vs the emulated split arithmetic: only fp32 accumulation-order noise
    assert np.max(np.abs(out - emu)) <= 2e-6 * scale * 8
    # vs exact fp64: the bf16x3 bound (~2^-16 relative per product)
    assert np.max(np.abs(out - exact)) <= 3e-5 * scale * 4


@pytest.mark.parametrize("m,n,k,ks", [(200, 200, 5000, 8), (200, 200, 100000, 64), (64, 48, 1000, 3),
                                      (512, 512, 20000, 16), (200, 200, 64, 4)])
def test_gemm_splitk(plda, m, n, k, ks):
    rng = np.random.RandomState(k)
    a = rng.randn(m, k)
    b = rng.randn(n, k)
    out = plda._test_gemm(a, b, ks).astype(np.float64)
    exact = a @ b.T
    assert np.max(np.abs(out - exact)) <= 3e-5 * np.sqrt(k) * 4


@pytest.mark.parametrize("d", [4, 33, 200, 256, 257, 512])
def test_cholesky_and_inverse(plda, d):
    rng = np.random.RandomState(d)
    g = rng.randn(d, 2 * d)
    a = g @ g.T / (2 * d) + 0.1 * np.eye(d)
    l, _ = plda._test_linalg(0, a)
    ref = np.linalg.cholesky(a)
    assert np.allclose(l, ref, rtol=1e-10, atol=1e-12)
    inv, _ = plda._test_linalg(1, ref)
    assert np.allclose(inv @ ref, np.eye(d), atol=1e-9)
    assert np.allclose(np.triu(inv, 1), 0.0)


@pytest.mark.parametrize("d", [1, 4, 31, 32, 33, 200, 256, 257, 480, 512])
def test_fused_cholesky_inverse(plda, d):
    """The single-launch cluster kernel (Cholesky + triangular inverse, one 32-row block per CTA) vs numpy."""
    rng = np.random.RandomState(d + 7)
    g = rng.randn(d, 2 * d + 3)
    a = g @ g.T / (2 * d) + 0.1 * np.eye(d)
    ref = np.linalg.cholesky(a)
    l, _ = plda._test_linalg(3, a)
    assert np.allclose(l, ref, rtol=1e-10, atol=1e-12)
    assert np.allclose(np.triu(l, 1), 0.0)
    inv, _ = plda._test_linalg(4, a)
    assert np.allclose(inv @ ref, np.eye(d), atol=1e-9)
    assert np.allclose(np.triu(inv, 1), 0.0)


# 64 .. 256: two CTAs per block pair (halves of the column length, partial Gram matrices over distributed shared
# memory); 63 and below / above 256: one CTA per pair; 600: grid-barrier form with 8-column blocks
@pytest.mark.parametrize("d", [2, 7, 63, 64, 77, 130, 200, 201, 255, 256, 257, 512, 600])
def test_jacobi_eig(plda, d):
    rng = np.random.RandomState(d + 1)
    q, _ = np.linalg.qr(rng.randn(d, d))
    lam = np.sort(2.0 * np.exp(-np.arange(d) / (0.15 * d)) + 1e-4)[::-1]
    a = (q * lam) @ q.T
    a = 0.5 * (a + a.T)
    v, w = plda._test_linalg(2, a)
    assert np.allclose(w, lam, rtol=1e-9, atol=1e-12)
    assert np.all(np.diff(w) <= 0)
    assert np.allclose(v.T @ v, np.eye(d), atol=1e-10)
    assert np.allclose(a @ v, v * w, atol=1e-9)


def test_jacobi_degenerate_spectrum(plda):
    """C1-like spectrum: one dominant eigenvalue, the rest nearly equal and tiny."""
    d = 120
    rng = np.random.RandomState(5)
    q, _ = np.linalg.qr(rng.randn(d, d))
    lam = np.full(d, 4.4e-4)
    lam[0] = 0.73
    lam[1:] += 1e-7 * rng.rand(d - 1)
    lam = np.sort(lam)[::-1]
    a = (q * lam) @ q.T
    a = 0.5 * (a + a.T)
    v, w = plda._test_linalg(2, a)
    assert np.allclose(w, lam, rtol=1e-8, atol=1e-13)
    assert np.allclose(v.T @ v, np.eye(d), atol=1e-10)


def test_gram_kernel_with_enrol_operand_in_tmem_matches_default():
    """The opt-in TS form of the score-grid kernel (csrc/gemm_ts.cu: enrol operand in tensor memory, PLDA_B200_TS=1)
    issues the same three products per k-step in the same order: its grid is bit-identical to the default kernel's,
    with and without the z-norm affine, at tile-edge shapes."""
    import os
    import torch
    from plda_b200 import PLDA
    d = 200
    rs = np.random.RandomState(5)
    q, _ = np.linalg.qr(rs.randn(d, d))
    model = (np.full(d, 0.5), q, 2.0 * np.exp(-np.arange(d) / (0.15 * d)))
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    ne, nt = 1300, 1111
    e = torch.randn(ne, d, device="cuda", generator=g)
    t = torch.randn(nt, d, device="cuda", generator=g)
    zm = torch.randn(ne, device="cuda", generator=g).double()
    zs = (1.0 + torch.rand(ne, device="cuda", generator=g)).double()
    outs = {}
    old = os.environ.get("PLDA_B200_TS")
    try:
        for mode in ("0", "1"):
            os.environ["PLDA_B200_TS"] = mode          # read when the handle is created
            p = PLDA()
            p.set_model(*model)
            outs[mode] = (p.score_grid(e, 3, t).clone(), p.score_grid(e, 3, t, znorm=(zm, zs)).clone())
            del p
    finally:
        if old is None:
            os.environ.pop("PLDA_B200_TS", None)
        else:
            os.environ["PLDA_B200_TS"] = old
    assert torch.equal(outs["0"][0], outs["1"][0])
    assert torch.equal(outs["0"][1], outs["1"][1])
    assert float(outs["0"][0].abs().max()) > 1.0
