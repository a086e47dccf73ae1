"""Math utilities with the reference's names (MC-GRA/utils.py hot subset)."""
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
