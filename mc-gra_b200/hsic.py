"""Gaussian-kernel HSIC library with the reference's function names (MC-GRA/hsic.py), on fused native kernels:
no m x m kernel matrix is stored for the scalar statistics (mcgra_gauss_stats), dense outputs only where the
reference function itself returns a matrix.  Inputs must live on a CUDA device (no CPU fallback)."""
import numpy as np
import torch

from . import _native as N
from ._native import call, ptr


def _f32(X):
    if not X.is_cuda:
        raise N.NativeError("mcgra_b200.hsic needs CUDA tensors (no CPU fallback)")
    return X.detach().to(torch.float32).contiguous()


def distmat(X):
    """Squared pairwise distances (hsic.py:20-27)."""
    X = _f32(X)
    m, d = X.shape
    out = torch.empty(m, m, dtype=torch.float32, device=X.device)
    call("mcgra_pair_dense", ptr(X), d, m, ptr(X), m, 0, 0.0, ptr(out), N.stream_ptr())
    return out


def sigma_estimation(X, Y):
    """Median heuristic (hsic.py:5-17): median of the strict lower triangle of distmat([X; Y])."""
    D = distmat(torch.cat([X, Y]))
    m = D.shape[0]
    ii = torch.tril_indices(m, m, -1, device=D.device)
    tri = D[ii[0], ii[1]].sort().values
    k = tri.numel()
    med = float(tri[k // 2]) if k % 2 else 0.5 * float(tri[k // 2 - 1] + tri[k // 2])
    if med <= 0:
        med = float(tri.mean())
    if med < 1e-2:
        med = 1e-2
    return med


def _gamma(X, sigma):
    s = sigma if sigma else sigma_estimation(X, X)
    return 1.0 / (2.0 * s * s)


def kernelmat(X, sigma):
    """exp(-D / 2 sigma^2) @ H, i.e. every row minus its mean (hsic.py:30-47)."""
    X = _f32(X)
    m, d = X.shape
    K = torch.empty(m, m, dtype=torch.float32, device=X.device)
    call("mcgra_pair_dense", ptr(X), d, m, ptr(X), m, 1, float(_gamma(X, sigma)), ptr(K), N.stream_ptr())
    return K - K.mean(1, keepdim=True)


def _stats(X, Y, gx, gy):
    X, Y = _f32(X), _f32(Y)
    m = X.shape[0]
    rowK = torch.zeros(m, dtype=torch.float32, device=X.device)
    rowL = torch.zeros(m, dtype=torch.float32, device=X.device)
    out = torch.zeros(3, dtype=torch.float64, device=X.device)
    call("mcgra_gauss_stats", ptr(X), X.shape[1], ptr(Y), Y.shape[1], m, float(gx), float(gy), ptr(rowK), ptr(rowL),
         ptr(out), N.stream_ptr())
    return m, rowK.double(), rowL.double(), out


def _tr_khlh(X, Y, gx, gy):
    m, rk, rl, o = _stats(X, Y, gx, gy)
    return o[0] - (2.0 / m) * (rk * rl).sum() + o[1] * o[2] / (m * m), m


def distcorr(X, sigma=1.0):
    """mean(exp(-D / 2 sigma^2)) (hsic.py:50-53)."""
    m, _, _, o = _stats(X, X, 1.0 / (2.0 * sigma * sigma), 0.0)
    return (o[1] / (m * m)).float()


def compute_kernel(x, y):
    """exp(-mean_k (x_ik - y_jk)^2 / dim) (hsic.py:56-66) -> dense [x_size, y_size]."""
    x, y = _f32(x), _f32(y)
    d = x.shape[1]
    out = torch.empty(x.shape[0], y.shape[0], dtype=torch.float32, device=x.device)
    call("mcgra_pair_dense", ptr(x), d, x.shape[0], ptr(y), y.shape[0], 1, 1.0 / (d * float(d)), ptr(out), N.stream_ptr())
    return out


def _mean_kernel(x, y, gamma):
    x, y = _f32(x), _f32(y)
    K = torch.empty(x.shape[0], y.shape[0], dtype=torch.float32, device=x.device)
    call("mcgra_pair_dense", ptr(x), x.shape[1], x.shape[0], ptr(y), y.shape[0], 1, float(gamma), ptr(K), N.stream_ptr())
    return K.double().mean()


def mmd(x, y, sigma=None, use_cuda=True, to_numpy=False):
    """hsic.py:69-90."""
    if sigma:
        gx = gy = 1.0 / (2.0 * sigma * sigma)
        sxy = sigma
    else:
        gx, gy = _gamma(x, None), _gamma(y, None)
        sxy = sigma_estimation(x, y)
    mx, _, _, ox = _stats(x, x, gx, 0.0)
    my, _, _, oy = _stats(y, y, gy, 0.0)
    val = ox[1] / (mx * mx) + oy[1] / (my * my) - 2 * _mean_kernel(x, y, 1.0 / (sxy * sxy))
    return val.float()


def mmd_pxpy_pxy(x, y, sigma=None, use_cuda=True, to_numpy=False):
    """hsic.py:93-114: mean(Kx.Ky) - 2 mean(colmean Kx . colmean Ky) + mean Kx mean Ky."""
    m, rk, rl, o = _stats(x, y, _gamma(x, sigma), _gamma(y, sigma))
    A = o[0] / (m * m)
    B = ((rk / m) * (rl / m)).mean()
    Cc = (o[1] / (m * m)) * (o[2] / (m * m))
    return (A - 2 * B + Cc).float()


def hsic_regular(x, y, sigma=None, use_cuda=True, to_numpy=False):
    """mean((Kx H) . (Ky H)^T) = tr(Kx H Ky H) / m^2 (hsic.py:117-124)."""
    tr, m = _tr_khlh(x, y, _gamma(x, sigma), _gamma(y, sigma))
    return (tr / (m * m)).float()


def hsic_normalized(x, y, sigma=None, use_cuda=True, to_numpy=True):
    """hsic.py:127-135."""
    Pxy = hsic_regular(x, y, sigma)
    Px = torch.sqrt(hsic_regular(x, x, sigma))
    Py = torch.sqrt(hsic_regular(y, y, sigma))
    return Pxy / (Px * Py)


def hsic_normalized_cca(x, y, sigma=None, use_cuda=True, to_numpy=True):
    raise NotImplementedError("hsic_normalized_cca (two dense m x m inverses, unused by the reference) is out of scope "
                              "(SURVEY.md 2 row 4)")
