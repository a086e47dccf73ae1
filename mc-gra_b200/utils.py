"""Math utilities with the reference's names (MC-GRA/utils.py hot subset)."""
import math

import numpy as np
import scipy.sparse as sp
import torch

Align_Parameter_Cora = {"c1": 100, "c2": 1000, "c3": 100, "c4": 10, "c5": 10, "c6": 10, "c7": 10, "c8": 0.01,
                        "c9": 1, "c10": 1}          # utils.py:1100-1111


def to_tensor(adj, features, labels=None, device='cpu'):
    """scipy / numpy -> torch tensors on `device` (utils.py:92-120)."""
    def conv(m):
        if sp.issparse(m):
            m = m.tocoo().astype(np.float32)
            idx = torch.from_numpy(np.vstack((m.row, m.col)).astype(np.int64))
            return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data), torch.Size(m.shape))
        return torch.as_tensor(np.asarray(m), dtype=torch.float32)
    adj, features = conv(adj), conv(features)
    if labels is None:
        return adj.to(device), features.to(device)
    return adj.to(device), features.to(device), torch.as_tensor(np.asarray(labels), dtype=torch.int64).to(device)


def normalize_adj_tensor(adj, sparse=False):
    """D^-1/2 (A+I) D^-1/2 of a dense tensor (utils.py:211-230) as a row/column scaling -- value-identical to
    the reference's two dense-diagonal matmuls ((r_i * m_ij) * r_j), without the O(n^3) work; autograd-capable."""
    if sparse or adj.is_sparse:
        adj = adj.to_dense()
    mx = adj + torch.eye(adj.shape[0], device=adj.device, dtype=adj.dtype)
    r_inv = mx.sum(1).pow(-1 / 2).flatten()
    r_inv = torch.where(torch.isinf(r_inv), torch.zeros_like(r_inv), r_inv)
    return (r_inv[:, None] * mx) * r_inv[None, :]


def accuracy(output, labels):
    """utils.py:286-308."""
    if type(labels) is not torch.Tensor:
        labels = torch.LongTensor(labels)
    preds = output.max(1)[1].type_as(labels)
    return preds.eq(labels).double().sum() / len(labels)


# ----------------------------------------------------------------------------------------------------------------
# HSIC / CKA surface (utils.py:803-822, 1056-1097) on the native fused kernels (csrc/hsic.cu)
# ----------------------------------------------------------------------------------------------------------------
def pairwise_distances(x):
    """utils.py:803-806."""
    from . import hsic as _h
    return _h.distmat(x)


def GaussianKernelMatrix(x, sigma=5):
    """exp(-dist / sigma) (utils.py:809-811) -> dense."""
    from . import _native as N
    x = x.detach().to(torch.float32).contiguous()
    m, d = x.shape
    out = torch.empty(m, m, dtype=torch.float32, device=x.device)
    N.call("mcgra_pair_dense", N.ptr(x), d, m, N.ptr(x), m, 1, 1.0 / float(sigma), N.ptr(out), N.stream_ptr())
    return out


def HSIC(x, y, s_x=1, s_y=1):
    """tr(L H K H) / (m-1)^2 with K = exp(-dist/s_x), L = exp(-dist/s_y) (utils.py:814-822)."""
    from . import hsic as _h
    tr, m = _h._tr_khlh(x, y, 1.0 / float(s_x), 1.0 / float(s_y))
    return (tr / ((m - 1) ** 2)).float()


class CudaCKA(object):
    """Same methods as the reference's CudaCKA (utils.py:1056-1097)."""

    def __init__(self, device):
        self.device = device

    def centering(self, K):
        """H K H without the dense H products (utils.py:1060-1065)."""
        return K - K.mean(0, keepdim=True) - K.mean(1, keepdim=True) + K.mean()

    def _sigma(self, X):
        from . import hsic as _h
        KX = _h.distmat(X)
        return math.sqrt(float(torch.median(KX[KX != 0])))            # utils.py:1070-1072

    def rbf(self, X, sigma=None):
        from . import _native as N
        if sigma is None:
            sigma = self._sigma(X)
        X = X.detach().to(torch.float32).contiguous()
        m, d = X.shape
        out = torch.empty(m, m, dtype=torch.float32, device=X.device)
        N.call("mcgra_pair_dense", N.ptr(X), d, m, N.ptr(X), m, 1, 0.5 / (sigma * sigma), N.ptr(out), N.stream_ptr())
        return out

    def kernel_HSIC(self, X, Y, sigma):
        from . import hsic as _h
        sx = sigma if sigma is not None else self._sigma(X)
        sy = sigma if sigma is not None else self._sigma(Y)
        tr, _ = _h._tr_khlh(X, Y, 0.5 / (sx * sx), 0.5 / (sy * sy))
        return tr.float()

    def linear_HSIC(self, X, Y):
        """sum (H XX^T H) . (H YY^T H) = ||(HX)^T (HY)||_F^2 (utils.py:1080-1084): weighted moments for narrow operands,
        the tcgen05 contraction (mcgra_gemm_nt) for wide ones (dense_measure.cross_frobenius)."""
        from .dense_measure import cross_frobenius
        return cross_frobenius(X, Y, center=True).float()

    def linear_CKA(self, X, Y):
        hsic = self.linear_HSIC(X, Y)
        return hsic / (torch.sqrt(self.linear_HSIC(X, X)) * torch.sqrt(self.linear_HSIC(Y, Y)))

    def kernel_CKA(self, X, Y, sigma=None):
        hsic = self.kernel_HSIC(X, Y, sigma)
        return hsic / (torch.sqrt(self.kernel_HSIC(X, X, sigma)) * torch.sqrt(self.kernel_HSIC(Y, Y, sigma)))
