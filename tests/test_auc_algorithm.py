"""CPU restatement of the ranking scheme of csrc/auc.cu (k_unique_pos, k_build_blocks, k_rank_negatives_tab, k_ap_finish)
in numpy, checked against the sklearn-semantics oracle in every table mode -- the GPU suite reaches the modes through the
data sizes (tests/test_gpu_metrics.py), this test through a small table capacity:

  (a) distinct keys U, start indices C and a private histogram fit         3 nU + 2 <= cap
  (b) U and C fit                                                          2 nU + 1 <= cap
  (c) sampled U + one block [step keys | step + 1 start indices] per sample, step <= 3
  (d) sampled U + search in the global array, step > 3

It pins the index arithmetic (segment bounds, block layout, tie handling, suffix sums of the histogram), not the kernel."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pgd_oracle as O  # noqa: E402


def fkey(f):
    """order-preserving float32 -> uint32 (auc.cu: fkey), -0.0 == +0.0"""
    u = (np.asarray(f, np.float32) + np.float32(0.0)).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def unique_pos(pos):
    """k_unique_pos: distinct keys and the index of their first occurrence, C[nU] = npos"""
    flag = np.r_[True, pos[1:] != pos[:-1]] if pos.size else np.zeros(0, bool)
    U = pos[flag]
    C = np.r_[np.nonzero(flag)[0], pos.size].astype(np.int64)
    return U, C


def tab_count(tab, k, le):
    """the uniform-trip-count halving search of the kernel: number of table entries < k (or <= k)"""
    ntab = tab.size
    if ntab == 0:
        return 0
    base, ln = 0, ntab
    cmp = (lambda v: v <= k) if le else (lambda v: v < k)
    while ln > 1:
        half = ln >> 1
        if cmp(tab[base + half - 1]):
            base += half
        ln -= half
    return base + (1 if cmp(tab[base]) else 0)


def rank_one(k, U, C, cap):
    """(lb, ub, mode) of one negative key, following k_rank_negatives_tab"""
    nU = U.size
    npos = int(C[-1])
    both = 2 * nU + 1 <= cap
    hsm = 3 * nU + 2 <= cap
    step = 1 if both else (nU + cap - 1) // cap
    ntab = nU if both else (nU + step - 1) // step
    tab = U[::step][:ntab]
    if both:
        u = tab_count(tab, k, le=False)
        tied = u < nU and tab[u] == k
        return int(C[u]), int(C[u + 1] if tied else C[u]), "a" if hsm else "b"
    if step <= 3:
        t = tab_count(tab, k, le=True) - 1
        if t < 0:
            return 0, 0, "c"
        w = np.full(8, 0xffffffff, np.uint64)                        # k_build_blocks
        for i in range(8):
            if i < step:
                idx = t * step + i
                if idx < nU:
                    w[i] = U[idx]
            elif i <= 2 * step:
                idx = t * step + (i - step)
                w[i] = C[min(idx, nU)]
        m = sum(1 for i in range(step) if w[i] < k)
        tied = any(w[i] == k for i in range(step))
        ci = step + m
        return int(w[ci]), int(w[ci + 1] if tied else w[ci]), "c"
    u = tab_count(tab, k, le=False)                                   # first SAMPLE not below k
    lo = 0 if u == 0 else (u - 1) * step + 1
    hi = nU if u == ntab else u * step
    ln = step
    while ln > 1:                                                     # entries beyond the segment count as +inf
        half = ln >> 1
        idx = lo + half - 1
        v = U[idx] if idx < hi else 0xffffffff
        if v < k:
            lo += half
        ln -= half
    v = U[lo] if lo < hi else 0xffffffff
    u = lo + (1 if v < k else 0)
    tied = u < nU and U[u] == k
    if nU == npos:
        return u, u + (1 if tied else 0), "d"
    return int(C[u]), int(C[u + 1] if tied else C[u]), "d"


def auc_ap(scores, labels, cap):
    keys = fkey(scores)
    pos = np.sort(keys[labels != 0])
    neg = keys[labels == 0]
    npos = pos.size
    U, C = unique_pos(pos)
    hist = np.zeros(npos + 1, np.int64)
    twice, modes = 0, set()
    for k in neg:
        lb, ub, mode = rank_one(int(k), U, C, cap)
        modes.add(mode)
        twice += 2 * (npos - ub) + (ub - lb)
        hist[ub] += 1
    suffix = np.cumsum(hist[::-1])[::-1]                              # k_ap_finish: hist[j] = sum_{q >= j}
    ap = 0.0
    for u in range(U.size):
        t, e = int(C[u]), int(C[u + 1])
        tp, fp = npos - t, int(suffix[t + 1]) if t + 1 <= npos else 0
        ap += (e - t) * tp / (tp + fp)
    return twice / (2.0 * npos * neg.size), ap / npos, modes


@pytest.mark.parametrize("n,frac,ties,cap,mode", [
    (3000, 0.05, False, 4096, "a"),      # ~150 distinct positives: everything in the table
    (3000, 0.30, False, 2048, "b"),      # ~900: 2 nU + 1 <= 2048 < 3 nU + 2
    (3000, 0.30, False, 1024, "c"),      # step 1: sampled == all keys, start indices from the blocks
    (3000, 0.30, False, 400, "c"),       # step 3
    (3000, 0.30, False, 128, "d"),       # step 8: search in the global array
    (3000, 0.30, True, 16, "d"),         # tied scores, few distinct keys, tiny table
    (3000, 0.30, True, 4096, "a"),
])
def test_ranking_scheme_matches_sklearn_semantics(n, frac, ties, cap, mode):
    rng = np.random.RandomState(n + cap)
    y = (rng.random_sample(n) < frac).astype(np.uint8)
    s = (rng.standard_normal(n) + 0.8 * y).astype(np.float32)
    if ties:
        s = (np.round(s * 20) / 20).astype(np.float32)
    s[:7] = [0.0, -0.0, 0.0, -0.0, 1.5, 1.5, -3.0]                    # signed zeros compare equal
    auc, ap, modes = auc_ap(s, y, cap)
    assert modes == {mode}
    assert abs(auc - O.roc_auc(y, s)) < 1e-12
    assert abs(ap - O.average_precision(y, s)) < 1e-12
