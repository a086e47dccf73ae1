"""GPU parity of the defended GCN training (mcgra_b200.mcgpb_gcn, BASELINE configs[2]) against fixtures produced by the
unmodified defence repo (tests/golden/make_golden_mcgpb.py): the MI penalties with their gradients through the native
moment kernels, the link AUROCs, and a replayed 3-epoch `fit` (dropout 0, recorded node-pair draws)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["linear_HSIC", "linear_CKA", "DP"])
def test_penalty_values_and_gradients(name):
    from mcgra_b200 import mcgpb_gcn as G
    d = np.load(os.path.join(GOLDEN, "mcgpb_penalties.npz"))
    dev = torch.device("cuda:0")
    Z = torch.from_numpy(d["Z"]).to(dev).requires_grad_(True)
    Zn = torch.from_numpy(d["Znext"]).to(dev).requires_grad_(True)
    v = getattr(G, name)(Zn, Z)
    gz, gzn = torch.autograd.grad(v, [Z, Zn])
    ref = float(d[f"{name}_value"])
    assert abs(float(v) - ref) <= 2e-4 * abs(ref)
    for g, key in ((gz, "gZ"), (gzn, "gZnext")):
        r = d[f"{name}_{key}"]
        assert np.max(np.abs(g.cpu().numpy() - r)) <= 2e-3 * np.max(np.abs(r))


@pytest.mark.parametrize("mi", ["linear_HSIC", "linear_CKA", "DP"])
def test_defended_fit_replays_reference(mi):
    from mcgra_b200.mcgpb_gcn import GCN
    d = np.load(os.path.join(GOLDEN, f"mcgpb_fit_{mi}_n300.npz"))
    dev = torch.device("cuda:0")
    f, c = d["X"].shape[1], int(d["labels"].max()) + 1
    model = GCN(nfeat=f, nclass=c, nhid=16, nlayer=2, dropout=0.0, weight_decay=5e-4, device=dev).to(dev)
    for layer in model.gc:
        layer.to(dev)
    sd = {k[5:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("init_")}
    model.load_state_dict(sd)
    with torch.no_grad():       # model.gc is a plain list whose first two layers alias gc1 / gc2 (models/gcn.py:140-149)
        assert model.gc[0] is model.gc1 and model.gc[1] is model.gc2
    beta = {str(k): float(v) for k, v in zip(d["beta_keys"], d["beta_vals"])}
    res = model.fit(torch.from_numpy(d["X"]), torch.from_numpy(d["adj"].astype(np.float32)), torch.from_numpy(d["labels"]),
                    d["idx_train"], d["idx_val"], d["idx_test"], train_iters=int(d["epochs"]), initialize=False, beta=beta,
                    MI_type=mi, plain_acc=0.7, pair_draws=d["pair_draws"])
    got = np.array(res["full_losses"])
    ref = d["full_losses"]
    rel = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12))
    print(f"[mcgpb-fit] {mi}: max rel err of (loss_IYZ, loss_IAZ, loss_inter, loss_mission) over {int(d['epochs'])} epochs {rel:.3e}")
    assert rel < 2e-3          # the reference evaluates every penalty through fp32 n^3 GEMMs
    np.testing.assert_allclose(res["IAZ"].numpy(), d["IAZ"], atol=1e-4)          # six link AUROCs per epoch
    np.testing.assert_allclose(res["IYZ"].numpy(), d["IYZ"], atol=1e-6)
    np.testing.assert_allclose(np.asarray(res["final_layer_aucs"], dtype=np.float64), d["final_layer_aucs"], atol=1e-4)
    for k in d.files:
        if k.startswith("final_") and k != "final_layer_aucs":
            w = model.state_dict()[k[6:]].cpu().numpy()
            # three Adam steps of lr 0.01: entries whose gradient sits at the fp32 noise floor may step differently
            bad = np.abs(w - d[k]) > 2e-4 * max(1.0, np.max(np.abs(d[k])))
            assert bad.mean() < 0.02 and np.max(np.abs(w - d[k])) <= 0.031, k
