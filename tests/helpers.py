"""Shared helpers for the GPU parity tests: run the native attack on a golden fixture."""
import os
import tempfile
from copy import deepcopy

import numpy as np
import torch


class Args:
    pass


def make_models(d, device):
    from mcgra_b200.models.gcn import GCN, embedding_GCN
    f = d["X"].shape[1]
    c = int(d["Wl"].shape[0])
    victim = GCN(nfeat=f, nclass=c, nhid=16, nlayer=2, dropout=0.5, weight_decay=5e-4, device=device)
    with torch.no_grad():
        victim.gc[0].weight.copy_(torch.from_numpy(d["W1"]))
        victim.gc[0].bias.copy_(torch.from_numpy(d["b1"]))
        victim.gc[1].weight.copy_(torch.from_numpy(d["W2"]))
        victim.gc[1].bias.copy_(torch.from_numpy(d["b2"]))
        victim.linear1.weight.copy_(torch.from_numpy(d["Wl"]))
        victim.linear1.bias.copy_(torch.from_numpy(d["bl"]))
    victim = victim.to(device)
    for layer in victim.gc:
        layer.to(device)
    emb = embedding_GCN(nfeat=f, nhid=16, nlayer=2, device=device)
    emb.gc = deepcopy(victim.gc)
    victim.eval()
    emb.eval()
    return victim, emb


def make_args(d):
    a = Args()
    a.max_eval = 100
    a.lr = float(d["lr_exp"])
    a.eps = float(d["eps"])
    a.measure = str(d["measure"])
    a.dataset = str(d["dataset"])
    a.useH_A, a.useY_A, a.useY = [bool(u) for u in d["use"]]
    for k in range(1, 11):
        setattr(a, f"w{k}", float(d["weights"][k - 1]))
    return a


def run_native_case(d, device="cuda:0", epochs=None, trace=True):
    """Drive mcgra_b200.topology_attack.PGDAttack.attack exactly like main.objective does (main.py:298-307)."""
    from mcgra_b200.topology_attack import PGDAttack
    dev = torch.device(device)
    n = int(d["labels"].shape[0])
    victim, emb = make_models(d, dev)
    args = make_args(d)
    H_A = torch.from_numpy(d["H_A2"]).to(dev)
    Y_A = torch.from_numpy(d["Y_A"]).to(dev)
    model = PGDAttack(model=victim, embedding=emb, H_A=H_A, Y_A=Y_A, nnodes=n, loss_type="CE", device=dev).to(dev)
    if np.any(d["x0"] != 0):
        model.adj_changes.data = torch.from_numpy(d["x0"].copy()).to(dev)
    epochs = int(d["epochs"]) if epochs is None else epochs
    adj = torch.from_numpy(d["adj"].astype(np.float32))
    feature_adj = torch.from_numpy(d["feature_adj"])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)   # no ./saved_data here: label_adj comes from the labels (same matrix, main.py:440-450)
        try:
            model.attack(args, None, 10 ** float(d["lr_exp"]), 0, float(d["weight_sup"]), tuple(d["weights"]),
                         feature_adj, 0, 0, 0, None, None, np.arange(min(8, n)), adj, d["X"],
                         np.zeros((n, n), np.float32), d["labels"], d["idx_attack"], int(d["num_edges"]), 0,
                         epochs=epochs, _trace=trace)
        finally:
            os.chdir(cwd)
    torch.cuda.synchronize()
    L = model.engine.losses()
    return dict(loss=L["loss"], terms=L, x_iters=[x.cpu().numpy() for x in model._trace],
                x_final=model.adj_changes.data.cpu().numpy(), modified_adj=model.modified_adj.cpu().numpy(),
                model=model)
