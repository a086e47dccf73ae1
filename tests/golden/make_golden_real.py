"""Golden fixtures on the REAL datasets of BASELINE.json configs[0..1] (Cora, Citeseer, Polblogs), produced by running the
UNMODIFIED reference (/root/reference/MC-GRA: dataset.Dataset, utils.preprocess, models.gcn.GCN.fit, topology_attack.
PGDAttack.attack) on CPU with the README "all three priors" commands (MC-GRA/README.md:29, :59, :90).

    python tests/golden/make_golden_real.py cora [citeseer polblogs]      # writes tests/golden/real_<ds>.npz

main.py itself cannot be imported (it parses argv and trains at import time), so its driver sequence (main.py:141-248,
262-312) is replayed here step by step with the reference's own classes.  Recorded: the inputs in sparse form
(features CSR, undirected edge list, labels, the TRAINED victim weights, idx_attack), the per-iteration loss of a SHORT
run (5 iterations), ROC-AUC / AP of its final scores, a sample of final score entries, and the ROC-AUC of the full
100-iteration README run.  The n x n matrices themselves are not stored (29-44 MB each).
"""
import os
import random
import sys
import tempfile
from copy import deepcopy

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref = ref_shim.load()
import importlib  # noqa: E402

R_dataset = importlib.import_module("dataset")
R_utils, R_ta, R_gcn = ref["utils"], ref["topology_attack"], ref["models.gcn"]
from sklearn.metrics import auc, average_precision_score, roc_curve  # noqa: E402

README = {   # MC-GRA/README.md "all three priors" rows
    "cora": dict(measure="MSELoss", lr=-2.0, w={1: 0.01, 6: 10, 7: 10, 9: 10, 10: 1000}),
    "citeseer": dict(measure="KL", lr=-1.5, w={1: 100, 2: 0.0001, 6: 0.001, 9: 1000, 10: 0.001}),
    "polblogs": dict(measure="HSIC", lr=-2.5, w={1: 0.01, 2: 0.01, 6: 10000, 7: 100, 9: 0.001, 10: 1000}),
}


def feature_adj_of(features, dataset):            # main.dot_product_decode, main.py:44-55
    Z = features
    if dataset in ("cora", "citeseer", "AIDS"):
        Z = torch.matmul(Z, Z.t())
        return torch.sigmoid(torch.relu(Z - torch.eye(Z.shape[0])))
    Z = torch.nn.functional.normalize(Z, p=2, dim=1)
    Z = torch.matmul(Z, Z.t())
    return torch.relu(Z - torch.eye(Z.shape[0]))


def run(ds, short_epochs=5, full_epochs=100, seed=15):
    cfg = README[ds]
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    data = R_dataset.Dataset(root="/root/reference/MC-GRA/dataset", name=ds, setting="GCN")     # main.py:148
    adj, features, labels, init_adj = data.adj, data.features, data.labels, data.init_adj
    idx_train, idx_val, idx_test = data.idx_train, data.idx_val, data.idx_test
    n = adj.shape[0]
    random.sample(range(n), n)                                                                  # main.py:155 (first draw)
    adj_sp, feat_sp = adj.copy(), features.copy()
    adj, features, labels = R_utils.preprocess(adj, features, labels, preprocess_adj=False, onehot_feature=False)
    feature_adj = feature_adj_of(features, ds)
    init_adj = torch.FloatTensor(init_adj.todense())
    victim = R_gcn.GCN(nfeat=features.shape[1], nclass=labels.max().item() + 1, nhid=16, nlayer=2, dropout=0.5,
                       weight_decay=5e-4, device="cpu")
    victim.fit(features, adj, labels, idx_train, idx_val)                                       # main.py:183
    emb = R_gcn.embedding_GCN(nfeat=features.shape[1], nhid=16, nlayer=2, device="cpu")
    emb.gc = deepcopy(victim.gc)                                                                # main.py:190
    victim.eval()
    emb.eval()
    with torch.no_grad():
        emb.set_layers(2)
        H_A2 = emb(features, adj)
        Y_A = victim(features, adj)
    idx_attack = np.array(random.sample(range(n), n))                                           # main.py:244-245
    num_edges = int(0.5 * 1e7 * adj.sum() / n ** 2 * len(idx_attack) ** 2)                      # --density 1e7

    class Args:
        pass
    args = Args()
    args.max_eval, args.lr, args.eps, args.measure, args.dataset = 100, cfg["lr"], 0.0, cfg["measure"], ds
    args.useH_A = args.useY_A = args.useY = True
    for k in range(1, 11):
        setattr(args, f"w{k}", cfg["w"].get(k, 0.0))
    wp = tuple(cfg["w"].get(k, 0.0) for k in range(1, 11))
    labels_np = labels.numpy()
    real = adj.numpy().reshape(-1)

    def attack(epochs):
        torch.manual_seed(seed)
        model = R_ta.PGDAttack(model=victim, embedding=emb, H_A=H_A2, Y_A=Y_A, nnodes=n, loss_type="CE", device="cpu")
        losses = []
        orig_backward = torch.Tensor.backward

        def rec_backward(self, *a, **k):
            losses.append(float(self.detach().double()))
            return orig_backward(self, *a, **k)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "saved_data"))
            np.save(os.path.join(td, "saved_data", ds + ".npy"),
                    (labels_np[:, None] == labels_np[None, :]).astype(np.float32))              # main.py:440-450
            os.chdir(td)
            torch.Tensor.backward = rec_backward
            try:
                model.attack(args, None, 10 ** cfg["lr"], 0, 1.0, wp, feature_adj, 0, 0, 0, idx_train, idx_val, idx_test,
                             adj, features, init_adj, labels, idx_attack, num_edges, 0, epochs=epochs)
            finally:
                torch.Tensor.backward = orig_backward
                os.chdir(cwd)
        final = model.modified_adj.detach().numpy()
        pred = final.reshape(-1)
        fpr, tpr, _ = roc_curve(real, pred)                                                     # main.metric_pool
        return losses, final, float(auc(fpr, tpr)), float(average_precision_score(real, pred)), \
            model.adj_changes.detach().numpy()

    losses_s, final_s, auc_s, ap_s, x_s = attack(short_epochs)
    rs = np.random.RandomState(1)
    si, sj = rs.randint(0, n, 8192), rs.randint(0, n, 8192)
    losses_f, final_f, auc_f, ap_f, _ = attack(full_epochs)
    fcsr = sp.csr_matrix(feat_sp)
    ee = sp.triu(sp.csr_matrix(adj_sp), 1).tocoo()
    sd = {k: v.detach().numpy() for k, v in (("W1", victim.gc[0].weight), ("b1", victim.gc[0].bias),
                                              ("W2", victim.gc[1].weight), ("b2", victim.gc[1].bias),
                                              ("Wl", victim.linear1.weight), ("bl", victim.linear1.bias))}
    out = dict(n=np.int64(n), feat_data=fcsr.data.astype(np.float32), feat_indices=fcsr.indices.astype(np.int32),
               feat_indptr=fcsr.indptr.astype(np.int64), feat_shape=np.array(fcsr.shape, dtype=np.int64),
               edges=np.stack([ee.row, ee.col], 1).astype(np.int32), labels=labels_np.astype(np.int64),
               idx_attack=idx_attack.astype(np.int64), num_edges=np.int64(num_edges),
               H_A2=H_A2.numpy(), Y_A=Y_A.numpy(), measure=np.array(cfg["measure"]), dataset=np.array(ds),
               lr_exp=np.float64(cfg["lr"]), weights=np.array(wp, dtype=np.float64),
               short_epochs=np.int64(short_epochs), loss_short=np.array(losses_s), auc_short=np.float64(auc_s),
               ap_short=np.float64(ap_s), x_short_sum=np.float64(x_s.astype(np.float64).sum()),
               sample_i=si.astype(np.int64), sample_j=sj.astype(np.int64), sample_short=final_s[si, sj].astype(np.float32),
               full_epochs=np.int64(full_epochs), loss_full=np.array(losses_f), auc_full=np.float64(auc_f),
               ap_full=np.float64(ap_f), sample_full=final_f[si, sj].astype(np.float32), **sd)
    np.savez_compressed(os.path.join(HERE, f"real_{ds}.npz"), **out)
    print(f"[golden] real_{ds}: n={n} measure={cfg['measure']} loss_short={losses_s} auc_short={auc_s:.5f} "
          f"auc_full({full_epochs})={auc_f:.5f} ap_full={ap_f:.5f}")


if __name__ == "__main__":
    for ds in sys.argv[1:] or ["cora"]:
        run(ds)
