"""AUC / AP / ranking on the GPU (reference: main.metric_pool, MC-GRA/main.py:66-75; AP of
gcn_parameterized.metric, gcn_parameterized.py:55-65) through mcgra_auc_ap / mcgra_argsort_desc."""
import numpy as np
import torch

from . import _native as N
from ._native import call, ptr


def roc_auc_ap(scores, labels, npos_max=None):
    """scores: float32 device tensor (any shape); labels: same numel, non-zero = positive.
    Returns (auc, ap) as Python floats (sklearn roc_curve+auc / average_precision_score semantics)."""
    s = scores.detach().reshape(-1).to(torch.float32).contiguous()
    lab = labels.reshape(-1)
    if lab.dtype != torch.uint8:
        lab = (lab != 0).to(torch.uint8)
    lab = lab.contiguous()
    assert s.is_cuda and lab.is_cuda and s.numel() == lab.numel()
    total = s.numel()
    if npos_max is None:
        npos_max = int(lab.sum().item())
    npos_max = max(int(npos_max), 1)
    ws = torch.empty(N.lib().mcgra_auc_workspace_bytes(total, npos_max), dtype=torch.uint8, device=s.device)
    out = torch.zeros(4, dtype=torch.float64, device=s.device)
    call("mcgra_auc_ap", ptr(s), ptr(lab), total, npos_max, ptr(ws), ptr(out), N.stream_ptr())
    o = out.cpu().numpy()
    if int(o[2]) > npos_max:
        raise N.NativeError(f"positives {int(o[2])} exceed npos_max {npos_max}")
    return float(o[0]), float(o[1])


def auc_ap_from_edges(scores_dense, edges):
    """All n^2 ordered pairs incl. the diagonal (metric_pool semantics); positives = both orientations of `edges`."""
    n = scores_dense.shape[0]
    dev = scores_dense.device
    e = torch.as_tensor(edges, device=dev).long()
    lab = torch.zeros(n, n, dtype=torch.uint8, device=dev)
    lab[e[:, 0], e[:, 1]] = 1
    lab[e[:, 1], e[:, 0]] = 1
    return roc_auc_ap(scores_dense, lab, npos_max=2 * int(e.shape[0]))


def metric_pool(ori_adj, inference_adj, idx, index_delete=None, print_cfg=False):
    """Same signature / value as main.metric_pool: ROC-AUC over adj[idx][:,idx] (index_delete has no effect on
    the AUC in the reference either, main.py:70-72)."""
    dev = inference_adj.device
    idx_t = torch.as_tensor(np.asarray(idx), device=dev).long()
    if torch.is_tensor(ori_adj) and ori_adj.is_sparse:      # label matrix as uint8 on the device, never a dense host n x n
        a = ori_adj.to(dev).coalesce()
        lab = torch.zeros(a.shape, dtype=torch.uint8, device=dev)
        ij = a.indices()
        lab[ij[0], ij[1]] = (a.values() != 0).to(torch.uint8)
        real = lab[idx_t][:, idx_t]
    else:
        real = ori_adj.to(dev)[idx_t][:, idx_t]
    pred = inference_adj[idx_t][:, idx_t]
    auc, _ = roc_auc_ap(pred, real)
    if print_cfg:
        print(f"current auc={auc}")
    return auc


def argsort_desc(scores):
    """Stable descending arg-sort of a float32 device tensor (recovered-edge ranking)."""
    s = scores.detach().reshape(-1).to(torch.float32).contiguous()
    total = s.numel()
    ws = torch.empty(N.lib().mcgra_sort_workspace_bytes(total), dtype=torch.uint8, device=s.device)
    order = torch.empty(total, dtype=torch.int64, device=s.device)
    call("mcgra_argsort_desc", ptr(s), total, ptr(order), ptr(ws), N.stream_ptr())
    return order


def auc_ap_from_edges_sharded(scores_band, row0, n, edges, group=None):
    """metric_pool semantics (all n^2 ordered pairs) when every rank holds a ROW BAND of the scores (rows row0 ...): the
    positives' keys are all-gathered, every rank ranks its own negatives against the global positives (mcgra_auc_stage),
    the integer counts are all-reduced and every rank finishes with the same AUC / AP."""
    import torch.distributed as dist
    dev = scores_band.device
    rows = scores_band.shape[0]
    world = dist.get_world_size(group)
    e = torch.as_tensor(edges, device=dev).long()
    lab = torch.zeros(rows, n, dtype=torch.uint8, device=dev)
    for a, b in ((e[:, 0], e[:, 1]), (e[:, 1], e[:, 0])):
        m = (a >= row0) & (a < row0 + rows)
        lab[a[m] - row0, b[m]] = 1
    npos_max = max(2 * int(e.shape[0]), 1)
    s = scores_band.detach().reshape(-1).to(torch.float32).contiguous()
    lab = lab.reshape(-1)
    total = s.numel()
    ws = torch.zeros(N.lib().mcgra_auc_workspace_bytes(max(total, 1), npos_max), dtype=torch.uint8, device=dev)
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    st = N.stream_ptr()
    call("mcgra_auc_stage", 0, ptr(s), ptr(lab), total, npos_max, ptr(ws), ptr(out), st)
    counter = ws[:8].view(torch.int64)
    keys = ws[256:256 + 4 * npos_max].view(torch.int32)
    cnt = counter.clone()
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    allk = [torch.empty(npos_max, dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(allk, keys.clone(), group=group)
    counts = [int(c.item()) for c in counts]
    if sum(counts) > npos_max:
        raise N.NativeError(f"positives {sum(counts)} exceed npos_max {npos_max}")
    keys.fill_(-1)                       # 0xffffffff: unused slots sort to the top
    o = 0
    for c, k in zip(counts, allk):
        keys[o:o + c] = k[:c]
        o += c
    counter.fill_(o)
    call("mcgra_auc_stage", 1, ptr(s), ptr(lab), total, npos_max, ptr(ws), ptr(out), st)
    ho = int(N.lib().mcgra_auc_hist_offset(npos_max))
    hist = ws[ho:ho + 8 * (npos_max + 2)].view(torch.int64)
    sums = ws[8:24].view(torch.int64)
    dist.all_reduce(hist, group=group)
    dist.all_reduce(sums, group=group)
    call("mcgra_auc_stage", 2, ptr(s), ptr(lab), total, npos_max, ptr(ws), ptr(out), st)
    r = out.cpu().numpy()
    return float(r[0]), float(r[1])
