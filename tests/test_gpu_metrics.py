"""GPU AUC / AP / arg-sort against sklearn-semantics oracle and the reference-generated fixture."""
import os

import numpy as np
import pytest
import torch

import pgd_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_auc_ap_matches_sklearn_fixture_with_ties():
    from mcgra_b200 import metrics
    fn = np.load(os.path.join(GOLDEN, "functions.npz"))
    s = torch.from_numpy(fn["auc_scores"]).cuda()
    y = torch.from_numpy(fn["auc_labels"]).cuda()
    auc, ap = metrics.roc_auc_ap(s, y)
    assert abs(auc - float(fn["auc_value"])) < 1e-12
    assert abs(ap - float(fn["ap_value"])) < 1e-12


@pytest.mark.parametrize("n,ties", [(1000, False), (300000, True), (2000003, False)])
def test_auc_ap_random(n, ties):
    from mcgra_b200 import metrics
    rng = np.random.RandomState(n % 1000)
    y = (rng.random_sample(n) < 0.03).astype(np.uint8)
    s = (rng.standard_normal(n) + 0.8 * y).astype(np.float32)
    if ties:
        s = np.round(s * 20) / 20
    auc, ap = metrics.roc_auc_ap(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda())
    assert abs(auc - O.roc_auc(y, s)) < 1e-10
    assert abs(ap - O.average_precision(y, s)) < 1e-10


def test_auc_edge_cases():
    from mcgra_b200 import metrics
    # all scores equal -> AUC 0.5; negative scores / negative zero / huge values keep their order
    y = torch.tensor([1, 0, 0, 1, 0], dtype=torch.uint8).cuda()
    auc, ap = metrics.roc_auc_ap(torch.zeros(5).cuda(), y)
    assert auc == 0.5 and abs(ap - 0.4) < 1e-12
    s = torch.tensor([-1e30, -3.0, -0.0, 0.0, 5e30]).cuda()
    yy = np.array([0, 1, 0, 1, 1], dtype=np.uint8)
    auc, ap = metrics.roc_auc_ap(s, torch.from_numpy(yy).cuda())
    assert abs(auc - O.roc_auc(yy, s.cpu().numpy())) < 1e-12
    assert abs(ap - O.average_precision(yy, s.cpu().numpy())) < 1e-12


@pytest.mark.parametrize("n", [5, 4096, 100001])
def test_argsort_desc_is_stable_and_bit_exact(n):
    from mcgra_b200 import metrics
    rng = np.random.RandomState(n)
    s = rng.standard_normal(n).astype(np.float32)
    s[::7] = s[0]                                  # ties: stable order must keep the lower index first
    order = metrics.argsort_desc(torch.from_numpy(s).cuda()).cpu().numpy()
    ref = np.argsort(-s.astype(np.float64), kind="stable")
    assert np.array_equal(order, ref)


def test_metric_pool_api():
    from mcgra_b200 import metrics
    d = np.load(os.path.join(GOLDEN, "attack_mse_A_n150.npz"))
    adj = torch.from_numpy(d["adj"].astype(np.float32)).cuda()
    inf = torch.from_numpy(d["modified_adj"]).cuda()
    auc = metrics.metric_pool(adj, inf, np.arange(150), None)
    assert abs(auc - float(d["auc"])) < 1e-9
    idx = np.random.RandomState(0).permutation(150)[:70]
    sub = metrics.metric_pool(adj, inf, idx, None)
    real = d["adj"][idx][:, idx].reshape(-1)
    pred = d["modified_adj"][idx][:, idx].reshape(-1)
    assert abs(sub - O.roc_auc(real, pred)) < 1e-9
