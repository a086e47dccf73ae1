"""Synthetic graphs of the shapes BASELINE.json names.

Host-side, numpy `RandomState` based so the same seed gives the same graph in the build container
and on the GPU box.  Nothing here is on the hot path.

* labels      y ~ U{0..c-1}
* adjacency   symmetric planted partition, P(edge)=p_in for equal labels else p_out, p_in = 8 p_out,
              mean degree `mean_deg` (4.5 = PubMed-like), zero diagonal.  Returned as an edge list
              (i > j), never dense.
* features    columns split into c blocks; X_ij ~ Bernoulli(0.02), Bernoulli(0.08) in the block of
              y_i; empty rows get one random 1; rows L1-normalised (bag-of-words like, in [0,1]).
"""
import numpy as np


def planted_partition_edges(labels, mean_deg=4.5, ratio=8.0, rng=None):
    """Edge list (E,2) int64 with i > j.  O(E) sampling (pair sampling with de-duplication)."""
    rng = rng or np.random.RandomState(15)
    n = labels.shape[0]
    classes, counts = np.unique(labels, return_counts=True)
    pairs_in = float(np.sum(counts.astype(np.float64) * (counts - 1) / 2))
    pairs_all = n * (n - 1) / 2.0
    pairs_out = pairs_all - pairs_in
    target_edges = mean_deg * n / 2.0
    p_out = target_edges / (ratio * pairs_in + pairs_out)
    p_in = ratio * p_out
    e_in = int(round(p_in * pairs_in))
    e_out = int(round(p_out * pairs_out))
    order = np.argsort(labels, kind="stable")
    starts = np.concatenate([[0], np.cumsum(counts)])
    # in-class pairs: pick a class proportional to its pair count, then two distinct members
    w = counts.astype(np.float64) * (counts - 1) / 2
    w = w / max(w.sum(), 1.0)
    cls = rng.choice(len(classes), size=int(e_in * 1.1) + 8, p=w)
    a = (rng.random_sample(cls.shape[0]) * counts[cls]).astype(np.int64)
    b = (rng.random_sample(cls.shape[0]) * counts[cls]).astype(np.int64)
    keep = a != b
    u = order[starts[cls[keep]] + a[keep]]
    v = order[starts[cls[keep]] + b[keep]]
    ein = np.stack([np.maximum(u, v), np.minimum(u, v)], 1)[:e_in]
    # out-of-class pairs: uniform pairs, reject equal labels
    u = rng.randint(0, n, size=int(e_out * 1.6) + 8)
    v = rng.randint(0, n, size=u.shape[0])
    keep = labels[u] != labels[v]
    eout = np.stack([np.maximum(u[keep], v[keep]), np.minimum(u[keep], v[keep])], 1)[:e_out]
    e = np.concatenate([ein, eout], 0).astype(np.int64)
    key = np.unique(e[:, 0] * n + e[:, 1])
    return np.stack([key // n, key % n], 1)


def block_features(labels, f, c, rng=None, p_bg=0.02, p_own=0.08):
    rng = rng or np.random.RandomState(16)
    n = labels.shape[0]
    X = np.zeros((n, f), dtype=np.float32)
    blk = np.minimum((np.arange(f) * c) // f, c - 1)
    chunk = 8192
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        u = rng.random_sample((e - s, f))
        own = blk[None, :] == labels[s:e, None]
        X[s:e] = (u < np.where(own, p_own, p_bg)).astype(np.float32)
    empty = np.where(X.sum(1) == 0)[0]
    if empty.size:
        X[empty, rng.randint(0, f, size=empty.size)] = 1.0
    X /= X.sum(1, keepdims=True)
    return X


def make_graph(n, f, c, seed=15, mean_deg=4.5):
    """Returns dict(labels[int64 n], edges[int64 E,2] with i>j, features[float32 n,f])."""
    rng = np.random.RandomState(seed)
    labels = rng.randint(0, c, size=n).astype(np.int64)
    edges = planted_partition_edges(labels, mean_deg=mean_deg, rng=rng)
    feats = block_features(labels, f, c, rng=rng)
    return {"labels": labels, "edges": edges, "features": feats}


def dense_adj(n, edges, dtype=np.float32):
    A = np.zeros((n, n), dtype=dtype)
    A[edges[:, 0], edges[:, 1]] = 1
    A[edges[:, 1], edges[:, 0]] = 1
    return A


def gcn_weights(f, h, c, seed=15, gain=1.0):
    """Seeded victim weights with the reference's init distributions
    (MC-GRA/models/gcn.py:28-33 uniform(+-1/sqrt(out)); nn.Linear default kaiming-uniform bound 1/sqrt(in))."""
    rng = np.random.RandomState(seed + 1000)
    s = gain / np.sqrt(h)
    W1 = rng.uniform(-s, s, size=(f, h)).astype(np.float32)
    b1 = rng.uniform(-s, s, size=(h,)).astype(np.float32)
    W2 = rng.uniform(-s, s, size=(h, h)).astype(np.float32)
    b2 = rng.uniform(-s, s, size=(h,)).astype(np.float32)
    Wl = rng.uniform(-s, s, size=(c, h)).astype(np.float32)   # nn.Linear weight is [out, in]
    bl = rng.uniform(-s, s, size=(c,)).astype(np.float32)
    return {"W1": W1, "b1": b1, "W2": W2, "b2": b2, "Wl": Wl, "bl": bl}
