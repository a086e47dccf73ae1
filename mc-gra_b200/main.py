"""CLI driver with the reference's flags (MC-GRA/main.py:78-139) on the native attack.

    python -m mcgra_b200.main --dataset cora --w1=0.01 --w6=10 --w7=10 --w9=10 --w10=1000 --lr=-2 \
           --useH_A --useY_A --useY --measure=MSELoss            (MC-GRA/README.md:29)

Data: `./dataset/<name>.npz` in the reference's npz layout (adj_data/indices/indptr/shape, attr_*, labels), or
`--dataset synthetic:<n>:<f>:<c>`.  Only `--mode evaluate` and `--arch gcn` are on the north-star path."""
import argparse
import os
import random
from copy import deepcopy

import numpy as np
import scipy.sparse as sp
import torch

from . import synth
from .metrics import metric_pool
from .models.gcn import GCN, embedding_GCN
from .topology_attack import PGDAttack


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--seed', type=int, default=15)
    p.add_argument('--epochs', type=int, default=100)
    p.add_argument('--lr', type=float, default=0.01)
    p.add_argument('--weight_decay', type=float, default=5e-4)
    p.add_argument('--hidden', type=int, default=16)
    p.add_argument('--dropout', type=float, default=0.5)
    p.add_argument('--nlayers', type=int, default=2)
    p.add_argument('--arch', type=str, choices=["gcn", "gat", "sage"], default='gcn')
    p.add_argument('--dataset', type=str, default='cora')
    p.add_argument('--density', type=float, default=10000000.0)
    p.add_argument('--model', type=str, default='PGD', choices=['PGD', 'min-max'])
    p.add_argument('--nlabel', type=float, default=1.0)
    p.add_argument('--iter', type=int, default=1)
    p.add_argument('--max_eval', type=int, default=100)
    p.add_argument('--log_name', type=str, default="result.txt")
    p.add_argument("--mode", type=str, default="evaluate")
    p.add_argument("--measure", type=str, default="HSIC", choices=["HSIC", "MSELoss", "KL", "KDE", "CKA", "DP"])
    p.add_argument("--measure2", type=str, default="HSIC")
    p.add_argument("--nofeature", action='store_true')
    p.add_argument('--weight_aux', type=float, default=0)
    p.add_argument('--weight_sup', type=float, default=1)
    for k in range(1, 11):
        p.add_argument(f'--w{k}', type=float, default=0)
    p.add_argument('--eps', type=float, default=0)
    p.add_argument('--useH_A', action='store_true')
    p.add_argument('--useY_A', action='store_true')
    p.add_argument('--useY', action='store_true')
    p.add_argument('--ensemble', action='store_true')
    p.add_argument('--add_noise', action='store_true')
    p.add_argument('--defense', action='store_true')
    return p


def load_graph(name, root='./dataset'):
    """(adj csr, features dense float32, labels int64).  npz layout of the reference's datasets (dataset.py:340-361)."""
    if name.startswith("synthetic"):
        _, n, f, c = name.split(":")
        g = synth.make_graph(int(n), int(f), int(c))
        e = g["edges"]
        adj = sp.coo_matrix((np.ones(len(e), np.float32), (e[:, 0], e[:, 1])), shape=(int(n), int(n)))
        adj = (adj + adj.T).tocsr()
        return adj, g["features"], g["labels"]
    with np.load(os.path.join(root, name + '.npz'), allow_pickle=True) as z:
        adj = sp.csr_matrix((z['adj_data'], z['adj_indices'], z['adj_indptr']), shape=z['adj_shape'])
        if 'attr_data' in z:
            feats = sp.csr_matrix((z['attr_data'], z['attr_indices'], z['attr_indptr']), shape=z['attr_shape'])
            feats = np.asarray(feats.todense(), dtype=np.float32)
        else:
            feats = np.eye(adj.shape[0], dtype=np.float32)
        labels = np.asarray(z['labels'] if 'labels' in z else z['node_labels']).astype(np.int64)
    adj = adj + adj.T
    adj = adj.tolil()
    adj.setdiag(0)
    adj = adj.tocsr()
    adj.eliminate_zeros()
    adj.data[:] = 1
    return adj.astype(np.float32), feats, labels


def dot_product_decode(Z, dataset):
    """main.dot_product_decode (main.py:44-55) -> feature_adj."""
    eye = torch.eye(Z.shape[0], device=Z.device)
    if dataset in ('cora', 'citeseer', 'AIDS') or dataset.startswith("synthetic"):
        return torch.sigmoid(torch.relu(Z @ Z.t() - eye))
    Zn = torch.nn.functional.normalize(Z, p=2, dim=1)
    return torch.relu(Zn @ Zn.t() - eye)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.arch != "gcn" or args.mode != "evaluate":
        raise NotImplementedError("only --arch gcn --mode evaluate are on the B200 path (SURVEY.md 2)")
    device = torch.device("cuda:0")
    np.random.seed(args.seed); random.seed(args.seed); torch.manual_seed(args.seed); torch.cuda.manual_seed(args.seed)
    adj_sp, feats, labels = load_graph(args.dataset)
    if args.dataset.startswith("synthetic"):
        # the reference's dataset-keyed tables (feature similarity main.py:44-55, dot_product_decode2
        # topology_attack.py:421-467) have no synthetic entry: synthetic graphs use the Cora branches
        args.dataset = "cora"
    n = adj_sp.shape[0]
    rng = np.random.RandomState(args.seed)
    perm = rng.permutation(n)
    idx_train, idx_val, idx_test = perm[:max(n // 10, 1)], perm[n // 10:n // 5], perm[n // 5:]
    idx_attack = np.array(random.sample(range(n), int(n * args.nlabel)))
    num_edges = int(0.5 * args.density * adj_sp.sum() / n ** 2 * len(idx_attack) ** 2)
    # the true adjacency stays sparse (COO on the device): no dense n x n on the host (main.py:162 densifies it there)
    coo = adj_sp.tocoo()
    adj = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([coo.row, coo.col]).astype(np.int64)),
                                  torch.from_numpy(coo.data.astype(np.float32)), (n, n)).coalesce().to(device)
    features = torch.from_numpy(feats)
    labels_t = torch.from_numpy(labels)
    feature_adj = dot_product_decode(features.to(device), args.dataset)
    if args.nofeature:
        feature_adj = torch.eye(n, device=device)
    victim = GCN(nfeat=features.shape[1], nclass=int(labels.max()) + 1, nhid=16, nlayer=args.nlayers, dropout=0.5,
                 weight_decay=5e-4, device=device).to(device)
    for layer in victim.gc:
        layer.to(device)
    victim.fit(features, adj, labels_t, idx_train, idx_val)
    embedding = embedding_GCN(nfeat=features.shape[1], nhid=16, nlayer=args.nlayers, device=device)
    embedding.gc = deepcopy(victim.gc)
    victim.eval(); embedding.eval()
    with torch.no_grad():
        H_A2 = embedding(features.to(device), adj.to(device))
        Y_A = victim(features.to(device), adj.to(device))
    model = PGDAttack(model=victim, embedding=embedding, H_A=H_A2, Y_A=Y_A, nnodes=n, loss_type='CE', device=device)
    weight_param = tuple(getattr(args, f"w{k}") for k in range(1, 11))
    model.attack(args, None, 10 ** args.lr, 0, args.weight_sup, weight_param, feature_adj, 0, 0, 0, idx_train, idx_val,
                 idx_test, adj, features, torch.zeros(1), labels_t, idx_attack, num_edges, 0, epochs=args.epochs)
    inference_adj = model.modified_adj
    auc = metric_pool(adj, inference_adj, idx_attack, None)
    auc_train = metric_pool(adj, inference_adj, idx_train, None)
    auc_all = metric_pool(adj, inference_adj, np.arange(n), None, True)
    os.makedirs("./results/", exist_ok=True)
    with open(os.path.join("./results", args.log_name), "a") as f:
        f.write(f"current parameter: {args}\n")
        f.write(f"In attack graph: AUC={auc}\tIn train graph: AUC={auc_train}\tIn Whole Graph: AUC={auc_all}\n")
        f.write(f"current density: {float(inference_adj.mean())}\n")
    return auc_all


if __name__ == "__main__":
    main()
