"""CPU checks of the closed-form values / gradients that the native measure stages implement
(csrc/dense.cu k_dense_scalars + dense_measure.DenseMeasure.step, csrc/ndmeasure.cu, csrc/kl2.cu) against autograd of the
oracle's dense formulation (oracle/pgd_oracle.py, which follows the reference line by line).  fp64, small n."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pgd_oracle as O  # noqa: E402



@pytest.fixture(autouse=True)
def _fp64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _operands(n=23, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n * (n - 1) // 2, generator=g)
    M = O.expand(x, n)
    A = O.normalize(M)
    z = torch.nn.functional.normalize(torch.relu(torch.randn(n, 16, generator=g)), dim=1)
    M1 = torch.relu(z @ z.t())
    M1.fill_diagonal_(0.0)
    F = torch.rand(n, n, generator=g)
    F = torch.sigmoid(torch.relu(F @ F.t() - torch.eye(n)))
    return A, M1, F


def _dense_closed_form(measure, A, M1, F, k1c, k2c):
    """Mirror of DenseMeasure.step + k_dense_scalars: returns (c1, c2, GA, GM) with sign applied to the gradients."""
    n = A.shape[0]
    cen = 0.0 if measure == "DP" else 1.0
    sign = -1.0 if measure == "HSIC" else 1.0
    H = torch.eye(n) - cen * torch.ones(n, n) / n
    Kf = H @ (F @ F.t()) @ H
    a, m = A.sum(1), M1.sum(1)
    G1 = Kf @ A
    S1 = (G1 * A).sum()
    T = A @ M1 - cen * torch.outer(a, m / n)
    hAM = (T ** 2).sum()
    SA = A @ A - cen * torch.outer(a, a / n)
    SM = M1 @ M1 - cen * torch.outer(m, m / n)
    hAA, hMM, hFF = (SA ** 2).sum(), (SM ** 2).sum(), (Kf ** 2).sum()
    a1 = a2 = a3 = a4 = a5 = 0.0
    if measure == "HSIC":
        c1, c2 = k1c * S1, k2c * hAM
        a1, a2, a4 = 2 * k1c, 2 * k2c, 2 * k2c
    elif measure == "DP":
        c1, c2 = k1c * S1.sqrt(), k2c * hAM.sqrt()
        a1 = k1c / S1.sqrt()
        a2 = a4 = k2c / hAM.sqrt()
    else:
        d1, d2 = (hFF * hAA).sqrt(), (hAA * hMM).sqrt()
        c1, c2 = k1c * S1 / d1, k2c * hAM / d2
        a1 = 2 * k1c / d1
        a2 = a4 = 2 * k2c / d2
        a3 = -2 * c1 / hAA - 2 * c2 / hAA
        a5 = -2 * c2 / hMM
    ones = torch.ones(n)
    GA = a1 * G1 + a2 * (M1 @ T.t() - cen * torch.outer(ones, T @ (m / n))) \
        + a3 * (A @ SA - cen * torch.outer(ones, (a / n) @ SA))
    GM = a4 * (A @ T - cen * torch.outer(ones, (a / n) @ T)) + a5 * (M1 @ SM - cen * torch.outer(ones, (m / n) @ SM))
    return c1, c2, sign * GA, sign * GM


@pytest.mark.parametrize("measure", ["HSIC", "DP", "CKA"])
def test_dense_measure_closed_form(measure):
    A, M1, F = _operands()
    k1c, k2c = 0.7, 1.3
    calc = O.pick_measure(measure)
    sign = -1.0 if measure == "HSIC" else 1.0
    Ar, Mr = A.clone().requires_grad_(True), M1.clone().requires_grad_(True)
    c1r, c2r = k1c * calc(F, Ar), k2c * calc(Ar, Mr)
    gA, gM = torch.autograd.grad(sign * (c1r + c2r), [Ar, Mr])
    c1, c2, GA, GM = _dense_closed_form(measure, A, M1, F, k1c, k2c)
    assert abs(float(c1 - c1r)) <= 1e-9 * abs(float(c1r))
    assert abs(float(c2 - c2r)) <= 1e-9 * abs(float(c2r))
    # the tiled pipeline consumes G_ij + G_ji (lower triangle) and the diagonal of GA; M1's diagonal is constant
    symA, symAr = GA + GA.t(), gA + gA.t()
    symM, symMr = GM + GM.t(), gM + gM.t()
    off = ~torch.eye(A.shape[0], dtype=torch.bool)
    assert torch.allclose(symA, symAr, rtol=1e-8, atol=1e-10 * float(symAr.abs().max()))
    assert torch.allclose(symM[off], symMr[off], rtol=1e-8, atol=1e-10 * float(symMr.abs().max()))


def _nd_closed_form(measure, X, Y, w, m, weight):
    """Mirror of csrc/ndmeasure.cu: value and per-node gradient w.r.t. Y from weighted moments."""
    sx, sy = (w[:, None] * X).sum(0), (w[:, None] * Y).sum(0)
    Sxy, Syy, Sxx = X.t() @ (w[:, None] * Y), Y.t() @ (w[:, None] * Y), X.t() @ (w[:, None] * X)
    if measure == "DP":
        V = m * (Sxy ** 2).sum().sqrt()
        P, Q, cx, cy = weight * m * Sxy / V, torch.zeros_like(Syy), torch.zeros_like(sx), torch.zeros_like(sy)
        value = V
    else:
        Cxy, Cyy, Cxx = m * (Sxy - torch.outer(sx, sy)), m * (Syy - torch.outer(sy, sy)), m * (Sxx - torch.outer(sx, sx))
        hxy, hyy, hxx = (Cxy ** 2).sum(), (Cyy ** 2).sum(), (Cxx ** 2).sum()
        if measure == "HSIC":
            value, sP, sQ = hxy, 2.0, 0.0
        else:
            den = hxx.sqrt() * hyy.sqrt()
            value = hxy / den
            sP, sQ = 2.0 / den, -2.0 * value / hyy
        P, Q, cx, cy = weight * sP * Cxy, weight * sQ * Cyy, sx, sy
    g = (w * m)[:, None] * ((X - cx) @ P + (Y - cy) @ Q)
    return weight * value, g


@pytest.mark.parametrize("measure", ["HSIC", "DP", "CKA"])
def test_nd_measure_closed_form(measure):
    g = torch.Generator().manual_seed(3)
    n, c, m = 40, 5, 55
    idx = torch.randint(0, n, (m,), generator=g)               # sub-sample WITH repetition (idx_attack is any index list)
    w = torch.bincount(idx, minlength=n).double() / m
    HA, em = torch.randn(n, 16, generator=g), torch.randn(n, 16, generator=g)
    Wl, bl = torch.randn(c, 16, generator=g), torch.randn(c, generator=g)
    YA = torch.log_softmax(torch.randn(n, c, generator=g), 1)
    calc = O.pick_measure(measure)
    w9, w10 = -0.3, 1.7
    emr = em.clone().requires_grad_(True)
    c9r = w9 * calc(HA[idx], emr[idx])
    p2 = torch.softmax(torch.log_softmax(emr @ Wl.t() + bl, 1), 1)
    c10r = w10 * calc(YA[idx], p2[idx])
    (gr,) = torch.autograd.grad(c9r + c10r, [emr])
    c9, g9 = _nd_closed_form(measure, HA, em, w, float(m), w9)
    p = torch.softmax(em @ Wl.t() + bl, 1)
    c10, gp = _nd_closed_form(measure, YA, p, w, float(m), w10)
    dz = p * (gp - (gp * p).sum(1, keepdim=True))
    gtot = g9 + dz @ Wl
    assert abs(float(c9 - c9r)) <= 1e-9 * abs(float(c9r)) and abs(float(c10 - c10r)) <= 1e-9 * abs(float(c10r))
    assert torch.allclose(gtot, gr, rtol=1e-7, atol=1e-10 * float(gr.abs().max()))


def test_kl2_closed_form():
    """csrc/kl2.cu: c2 = KL(softmax(A) || softmax(M1)) rows, c1 = KL(softmax(F) || softmax(A)) rows (calc_kl, :483-487)."""
    A, M1, F = _operands(n=19, seed=5)
    n = A.shape[0]
    k1c, k2c = 0.9, 2.1
    Ar, Mr = A.clone().requires_grad_(True), M1.clone().requires_grad_(True)
    c1r, c2r = k1c * O.calc_kl(F, Ar), k2c * O.calc_kl(Ar, Mr)
    gA, gM = torch.autograd.grad(c1r + c2r, [Ar, Mr])
    lseA, lseM, lseF = torch.logsumexp(A, 1), torch.logsumexp(M1, 1), torch.logsumexp(F, 1)
    pA, qM, pF = torch.exp(A - lseA[:, None]), torch.exp(M1 - lseM[:, None]), torch.exp(F - lseF[:, None])
    kl = (pA * (A - M1)).sum(1)
    c2 = k2c / n * (kl - lseA + lseM).sum()
    c1 = k1c / n * ((pF * (F - A)).sum(1) - (lseF - lseA)).sum()
    GA = k2c / n * pA * (A - M1 - kl[:, None]) + k1c / n * (pA - pF)
    GM = k2c / n * (qM - pA)
    assert abs(float(c1 - c1r)) <= 1e-10 * abs(float(c1r)) and abs(float(c2 - c2r)) <= 1e-10 * abs(float(c2r))
    assert torch.allclose(GA, gA, rtol=1e-8, atol=1e-14)
    assert torch.allclose(GM, gM, rtol=1e-8, atol=1e-14)


def test_fp16x2_image_precision():
    """The operand image of csrc/dense.cu (hi = fp16(s x), lo = fp16(s x - hi), s = 2^k with max |s x| <= 2^14) keeps
    ~2^-22 of the row maximum, and the three-term product hi*hi + hi*lo + lo*hi matches an fp32 GEMM's error class."""
    rng = np.random.RandomState(0)
    X = (rng.randn(64, 512) * np.exp(rng.randn(64, 1) * 3)).astype(np.float32)
    Y = rng.rand(48, 512).astype(np.float32) * 1e-3

    def image(M):
        mx = np.abs(M).max(1)
        e = np.floor(np.log2(mx)) + 1                      # mx < 2^e
        s = np.exp2(14 - e)[:, None].astype(np.float32)
        hi = (M * s).astype(np.float16)
        lo = (M * s - hi.astype(np.float32)).astype(np.float16)
        return hi.astype(np.float64), lo.astype(np.float64), s.astype(np.float64)

    xh, xl, sx = image(X)
    yh, yl, sy = image(Y)
    rec = (xh + xl) / sx
    assert np.max(np.abs(rec - X) / np.abs(X).max(1, keepdims=True)) < 2.0 ** -22
    acc = xh @ yh.T + xh @ yl.T + xl @ yh.T
    C = acc / sx / sy.T
    ref = X.astype(np.float64) @ Y.astype(np.float64).T
    scale = np.abs(X).astype(np.float64) @ np.abs(Y).astype(np.float64).T
    assert np.max(np.abs(C - ref) / scale) < 2.0 ** -20


@pytest.mark.parametrize("name", ["linear_HSIC", "linear_CKA", "DP"])
def test_mcgpb_penalties_from_factor_moments(name, monkeypatch):
    """mcgpb_gcn.{linear_HSIC, linear_CKA, DP}: the MC-GPB penalties IAZ(N N^T, Z) (MC-GPB/utils.py:774-797) as functions
    of second moments of the factors, value and both gradients against the fixture produced by the unmodified reference
    (tests/golden/make_golden_mcgpb.py).  The moment kernel is replaced by its torch definition here (CPU)."""
    from mcgra_b200 import mcgpb_gcn as G
    d = np.load(os.path.join(ROOT, "tests", "golden", "mcgpb_penalties.npz"))

    def moments(Nf, Z):
        Nf, Z = Nf.double(), Z.double()
        return Nf.sum(0), Z.sum(0), Nf.t() @ Z, Z.t() @ Z, Nf.t() @ Nf
    monkeypatch.setattr(G, "_moments", moments)
    Z = torch.from_numpy(d["Z"]).double().requires_grad_(True)
    Zn = torch.from_numpy(d["Znext"]).double().requires_grad_(True)
    v = getattr(G, name)(Zn, Z)
    gz, gzn = torch.autograd.grad(v.double(), [Z, Zn])
    ref = float(d[f"{name}_value"])
    assert abs(float(v) - ref) <= 2e-4 * abs(ref)          # the fixture itself is an fp32 n^3 evaluation
    for g, key in ((gz, "gZ"), (gzn, "gZnext")):
        r = d[f"{name}_{key}"]
        assert np.max(np.abs(g.numpy() - r)) <= 2e-3 * np.max(np.abs(r))


def _kde_scalars_mirror(mom, d, m, weight):
    """Mirror of csrc/kde.cu k_kde_scalars: value and d value / d moments from mom = [s1 | s2 | Sxy | Syy] (weights sum 1)."""
    eps, ln2 = 1e-10, float(np.log(2.0))
    s1, s2, Sxy = mom[:d], mom[d:2 * d], mom[2 * d:2 * d + d * d]
    fp = lambda p: -torch.log2(p + eps) - p / ((p + eps) * ln2)
    N1, N2, NJ = eps + s1.sum(), eps + s2.sum(), eps + m * Sxy.sum()
    p, q, pj = s1 / N1, s2 / N2, m * Sxy / NJ
    H1, H2, H12 = -(p * torch.log2(p + eps)).sum(), -(q * torch.log2(q + eps)).sum(), -(pj * torch.log2(pj + eps)).sum()
    Hs = H1 + H2
    value = 2 * (Hs - H12) / Hs
    cH, cJ = 2 * H12 / Hs ** 2, -2 / Hs
    g = torch.zeros_like(mom)
    g[:d] = weight * cH * (fp(p) - (fp(p) * p).sum()) / N1
    g[d:2 * d] = weight * cH * (fp(q) - (fp(q) * q).sum()) / N2
    g[2 * d:2 * d + d * d] = weight * cJ * m * (fp(pj) - (fp(pj) * pj).sum()) / NJ
    return weight * value, g


@pytest.mark.parametrize("shape", ["nn", "nd"])
def test_kde_closed_form(shape):
    """--measure KDE: the engine keeps the first 8 columns of an n x n operand (the others underflow in the Gaussian
    kernel) and works from weighted second moments of the kernel values; value and gradient must equal autograd of the
    oracle's utils.MutualInformation restatement on the full operands."""
    g = torch.Generator().manual_seed(5)
    if shape == "nn":
        A, M1, _ = _operands(n=29, seed=3)
        X, Y, bins = A.clone().requires_grad_(True), M1.clone().requires_grad_(True), 29
        nb = 8
    else:
        X = torch.rand(40, 16, generator=g).requires_grad_(True)
        Y = torch.softmax(torch.randn(40, 16, generator=g), 1).requires_grad_(True)
        bins, nb = 16, 16
    want = O.kde_mi(X, Y, bins) * 3.0
    gX, gY = torch.autograd.grad(want, (X, Y))
    m = X.shape[0]
    step = bins / (bins - 1.0)
    cols = torch.arange(nb, dtype=torch.float64) * step
    xs, ys = X.detach()[:, :nb], Y.detach()[:, :nb]
    kx, ky = torch.exp(-0.5 * ((xs - cols) / 0.32) ** 2), torch.exp(-0.5 * ((ys - cols) / 0.32) ** 2)
    mom = torch.cat([kx.mean(0), ky.mean(0), (kx.t() @ ky / m).reshape(-1), torch.zeros(nb * nb)])
    val, gm = _kde_scalars_mirror(mom, nb, float(m), 3.0)
    assert abs(float(val) - float(want.detach())) <= 1e-9 * abs(float(want.detach()))
    gS = gm[2 * nb:2 * nb + nb * nb].view(nb, nb)
    gkx = (gm[:nb] + ky @ gS.t()) / m             # mcgra_cross_moments_bwd with w = 1/m
    gky = (gm[nb:2 * nb] + kx @ gS) / m
    dX = gkx * kx * (-(xs - cols) / 0.32 ** 2)    # mcgra_kde_chain
    dY = gky * ky * (-(ys - cols) / 0.32 ** 2)
    scale = float(gX.abs().max())
    assert float((dX - gX[:, :nb]).abs().max()) <= 1e-8 * scale
    assert float((dY - gY[:, :nb]).abs().max()) <= 1e-8 * max(float(gY.abs().max()), 1e-30)
    if shape == "nn":        # the dropped columns carry no gradient
        assert float(gX[:, nb:].abs().max()) <= 1e-12 * scale
        assert float(gY[:, nb:].abs().max()) <= 1e-12 * float(gY.abs().max())
