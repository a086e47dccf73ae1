"""Golden fixtures from the UNMODIFIED defence repo (/root/reference/MC-GPB): the GraphMI attack with feature smoothing
(MC-GPB/topology_attack.py:36-87) and the MI-constrained GCN training penalties (MC-GPB/models/gcn.py:321-513).

Runs in its own process (MC-GPB's `utils`, `topology_attack`, `base_attack` shadow MC-GRA's module names):

    python tests/golden/make_golden_mcgpb.py          # writes tests/golden/mcgpb_*.npz

Import shims (third-party packages absent from this image; none of them is on the arithmetic recorded here):
`GCL.augmentors` (PyGCL EdgeRemoving, used only by utils.stochastic), `torch_geometric.utils` (same), `torchmetrics.AUROC`
(re-implemented with sklearn.roc_auc_score: same statistic).
"""
import importlib.util
import os
import random
import sys
import types
from copy import deepcopy

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/MC-GPB"


def install_shims():
    if not hasattr(np, "int"):
        np.int = int
    gcl = types.ModuleType("GCL")
    aug = types.ModuleType("GCL.augmentors")
    gcl.augmentors = aug
    sys.modules["GCL"], sys.modules["GCL.augmentors"] = gcl, aug
    tg = types.ModuleType("torch_geometric")
    tgu = types.ModuleType("torch_geometric.utils")
    tg.utils = tgu
    sys.modules["torch_geometric"], sys.modules["torch_geometric.utils"] = tg, tgu
    tm = types.ModuleType("torchmetrics")

    class AUROC:
        def __init__(self, task="binary"):
            pass

        def __call__(self, pred, real):
            from sklearn.metrics import roc_auc_score
            return torch.tensor(roc_auc_score(real.cpu().numpy().astype(int), pred.cpu().numpy()))
    tm.AUROC = AUROC
    sys.modules["torchmetrics"] = tm
    sys.path.insert(0, REF)


install_shims()
import topology_attack as R_ta  # noqa: E402
import utils as R_utils  # noqa: E402
from models import gcn as R_gcn  # noqa: E402

_spec = importlib.util.spec_from_file_location("mcgra_synth", os.path.join(ROOT, "mc-gra_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)


def build_victim(X, W, c):
    f = X.shape[1]
    victim = R_gcn.GCN(nfeat=f, nclass=c, nhid=16, nlayer=2, dropout=0.5, weight_decay=5e-4, device="cpu")
    with torch.no_grad():
        victim.gc[0].weight.copy_(torch.from_numpy(W["W1"])); victim.gc[0].bias.copy_(torch.from_numpy(W["b1"]))
        victim.gc[1].weight.copy_(torch.from_numpy(W["W2"])); victim.gc[1].bias.copy_(torch.from_numpy(W["b2"]))
        victim.linear1.weight.copy_(torch.from_numpy(W["Wl"])); victim.linear1.bias.copy_(torch.from_numpy(W["bl"]))
    emb = R_gcn.embedding_GCN(nfeat=f, nhid=16, nlayer=2, device="cpu")
    emb.gc = deepcopy(victim.gc)
    victim.eval()
    emb.eval()
    return victim, emb


def attack_case(name, n, f, c, epochs, seed=15, mean_deg=12.0, gain=3.0):
    g = synth.make_graph(n, f, c, seed=seed, mean_deg=mean_deg)
    X, labels = g["features"], g["labels"]
    A = synth.dense_adj(n, g["edges"])
    W = synth.gcn_weights(f, 16, c, seed=seed, gain=gain)
    victim, emb = build_victim(X, W, c)
    idx_attack = np.arange(n)                                           # MC-GPB/main.py:155
    num_edges = int(0.5 * 1.0 * A.sum() / n ** 2 * len(idx_attack) ** 2)  # --density 1.0 (main.py:120,156)
    model = R_ta.PGDAttack(model=victim, embedding=emb, nnodes=n, loss_type="CE", device="cpu")
    losses, xs = [], []
    orig_grad = torch.autograd.grad

    def rec_grad(loss, *a, **k):
        losses.append(float(loss.detach().double()))
        return orig_grad(loss, *a, **k)

    orig_proj = model.projection

    def rec_proj(ne):
        orig_proj(ne)
        xs.append(torch.clamp(model.adj_changes.detach().clone(), 0, 1).numpy())
    model.projection = rec_proj
    torch.autograd.grad = rec_grad
    try:
        out = model.attack(X, np.zeros((n, n), np.float32), labels, idx_attack, num_edges, epochs=epochs)
    finally:
        torch.autograd.grad = orig_grad
    keep = sorted(set([0, 1, 2, 49, 50, 51, epochs - 1]) & set(range(epochs)))
    rec = dict(X=X, adj=A.astype(np.uint8), labels=labels, idx_attack=idx_attack.astype(np.int64),
               num_edges=np.int64(num_edges), epochs=np.int64(epochs), loss=np.array(losses),
               x_keep_idx=np.array(keep), x_keep=np.stack([xs[k] for k in keep]).astype(np.float32),
               x_final=model.adj_changes.detach().numpy().astype(np.float32),
               modified_adj=model.modified_adj.numpy().astype(np.float32), output=out.numpy().astype(np.float32), **W)
    np.savez_compressed(os.path.join(HERE, f"mcgpb_attack_{name}.npz"), **rec)
    print(f"[golden] mcgpb_attack_{name}: n={n} epochs={epochs} loss[0]={losses[0]:.6f} loss[49]={losses[min(49, epochs - 1)]:.6f} "
          f"loss[-1]={losses[-1]:.6f} budget={num_edges} sum_x={xs[-1].sum():.2f}")


def penalty_fixtures():
    """The MI penalty functions of the defended training (MC-GPB/utils.py:774-797) on the operand shapes the training loop
    feeds them (models/gcn.py:400-417): a gram of embeddings against embeddings."""
    rng = np.random.RandomState(5)
    n, d, d2 = 210, 16, 7
    Z = rng.standard_normal((n, d)).astype(np.float32)
    Zn = (rng.standard_normal((n, d2)) * 0.5).astype(np.float32)
    out = dict(Z=Z, Znext=Zn)
    Zt, Znt = torch.from_numpy(Z).requires_grad_(True), torch.from_numpy(Zn).requires_grad_(True)
    for name in ("linear_HSIC", "linear_CKA", "DP"):
        fn = getattr(R_utils, name)
        v = fn(Znt @ Znt.T, Zt)
        gz, gzn = torch.autograd.grad(v, [Zt, Znt])
        out[f"{name}_value"] = v.detach().numpy()
        out[f"{name}_gZ"] = gz.numpy()
        out[f"{name}_gZnext"] = gzn.numpy()
    np.savez_compressed(os.path.join(HERE, "mcgpb_penalties.npz"), **out)
    print("[golden] mcgpb_penalties.npz", {k: float(v) for k, v in out.items() if k.endswith("_value")})


def fit_case(name, n, f, c, MI_type, epochs=3, seed=15):
    """MC-GPB GCN.fit with the MI constraints (models/gcn.py:184-277, 321-513), dropout 0 so that the run is a
    deterministic function of the seeded weights and node-pair draws; records the per-epoch losses and the final weights."""
    g = synth.make_graph(n, f, c, seed=seed, mean_deg=8.0)
    X, labels = g["features"], g["labels"]
    A = synth.dense_adj(n, g["edges"])
    rs = np.random.RandomState(seed)
    perm = rs.permutation(n)
    idx_train, idx_val, idx_test = perm[:n // 5], perm[n // 5:2 * n // 5], perm[2 * n // 5:]
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = R_gcn.GCN(nfeat=f, nclass=c, nhid=16, nlayer=2, dropout=0.0, weight_decay=5e-4, device="cpu")
    init = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    beta = {"layer-0": 0.05, "layer-1": 0.02, "layer-2": 0.01, "layer_inter-0": 0.03, "layer_inter-1": 0.01, "plain_acc": 0.7}
    draws = []
    orig_choice = np.random.choice

    def rec_choice(*a, **k):
        r = orig_choice(*a, **k)
        draws.append(np.asarray(r).copy())
        return r
    np.random.choice = rec_choice
    try:
        res = model.fit(torch.from_numpy(X), torch.from_numpy(A), torch.from_numpy(labels), idx_train, idx_val, idx_test,
                        train_iters=epochs, beta=beta, MI_type=MI_type, verbose=False, plain_acc=0.7)
    finally:
        np.random.choice = orig_choice
    final = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    out = dict(X=X, adj=A.astype(np.uint8), labels=labels, idx_train=idx_train, idx_val=idx_val, idx_test=idx_test,
               MI_type=np.array(MI_type), epochs=np.int64(epochs), pair_draws=np.stack(draws),
               full_losses=np.array(res["full_losses"], dtype=np.float64), IAZ=res["IAZ"].numpy(), IYZ=res["IYZ"].numpy(),
               final_layer_aucs=np.array(res["final_layer_aucs"], dtype=np.float64),
               beta_keys=np.array(list(beta.keys())), beta_vals=np.array(list(beta.values())),
               **{"init_" + k: v for k, v in init.items()}, **{"final_" + k: v for k, v in final.items()})
    np.savez_compressed(os.path.join(HERE, f"mcgpb_fit_{name}.npz"), **out)
    print(f"[golden] mcgpb_fit_{name}: MI_type={MI_type} losses(IYZ,IAZ,inter,mission) per epoch = {np.array(res['full_losses']).T.tolist()} "
          f"layer AUCs {res['final_layer_aucs']}")


if __name__ == "__main__":
    random.seed(15)
    np.random.seed(15)
    torch.manual_seed(15)
    which = sys.argv[1:] or ["attack", "penalties", "fit"]
    if "attack" in which:
        attack_case("n150", 150, 24, 4, epochs=56)
    if "penalties" in which:
        penalty_fixtures()
    if "fit" in which:
        for mi in ("linear_HSIC", "linear_CKA", "DP"):
            fit_case(f"{mi}_n300", 300, 24, 4, mi)
