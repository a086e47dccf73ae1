"""GPU HSIC / CKA surface against fixtures produced by the reference's own functions (functions.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fn():
    return np.load(os.path.join(GOLDEN, "functions.npz"))


def C(a):
    return torch.from_numpy(np.asarray(a)).cuda()


def close(a, b, rtol=2e-4, atol=1e-6):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def test_hsic_py_surface(fn):
    from mcgra_b200 import hsic as H
    X, Y = C(fn["hs_X"]), C(fn["hs_Y"])
    close(H.distmat(X), fn["hsic_distmat"], atol=2e-5)
    assert abs(H.sigma_estimation(X, Y[:, :7].repeat(1, 3)[:, :16]) - float(fn["hsic_sigma_est"])) < 1e-4 * float(fn["hsic_sigma_est"])
    close(H.kernelmat(X, 1.0), fn["hsic_kernelmat_s1"], atol=2e-6)
    close(H.kernelmat(X, None), fn["hsic_kernelmat_auto"], atol=2e-6)
    close(H.hsic_regular(X, Y, 1.0), fn["hsic_regular_s1"])
    close(H.hsic_regular(X, Y, None), fn["hsic_regular_auto"])
    close(H.hsic_normalized(X, Y, 1.0), fn["hsic_normalized_s1"])
    close(H.hsic_normalized(X, Y, None), fn["hsic_normalized_auto"])
    close(H.distcorr(X, 1.5), fn["hsic_distcorr"])
    close(H.compute_kernel(X, X[:20] * 0.5), fn["hsic_compute_kernel"])
    close(H.mmd(X, X * 0.7 + 0.1, 1.0), fn["hsic_mmd_s1"], atol=2e-6)
    close(H.mmd(X, X * 0.7 + 0.1, None), fn["hsic_mmd_auto"], atol=2e-6)
    close(H.mmd_pxpy_pxy(X, Y, 1.0), fn["hsic_mmd_pxpy_s1"], atol=1e-7)
    close(H.mmd_pxpy_pxy(X, Y, None), fn["hsic_mmd_pxpy_auto"], atol=1e-7)
    with pytest.raises(NotImplementedError):
        H.hsic_normalized_cca(X, Y)


def test_utils_cka_surface(fn):
    from mcgra_b200 import utils as U
    X, Y = C(fn["hs_X"]), C(fn["hs_Y"])
    cka = U.CudaCKA("cuda")
    close(cka.linear_HSIC(X, Y), fn["linear_HSIC"])
    close(cka.linear_CKA(X, Y), fn["linear_CKA"])
    close(cka.kernel_HSIC(X, Y, 2.0), fn["kernel_HSIC_s2"])
    close(cka.kernel_HSIC(X, Y, None), fn["kernel_HSIC_med"])
    close(cka.kernel_CKA(X, Y, 2.0), fn["kernel_CKA_s2"])
    close(cka.kernel_CKA(X, Y, None), fn["kernel_CKA_med"])
    close(cka.rbf(X, 2.0), fn["rbf_s2"], atol=2e-6)
    close(cka.centering(C(fn["norm_in"])), fn["centering"], atol=2e-6)
    close(U.HSIC(X, Y, 1, 1), fn["utils_HSIC_1_1"])
    close(U.HSIC(X, Y, 5, 3), fn["utils_HSIC_5_3"])
    close(U.normalize_adj_tensor(C(fn["norm_in"])), fn["norm_out"], rtol=1e-6, atol=1e-7)


def test_gauss_stats_large_m_matches_dense():
    from mcgra_b200 import hsic as H
    g = torch.Generator().manual_seed(3)
    X = torch.randn(1000, 16, generator=g).cuda()
    Y = (torch.randn(1000, 7, generator=g) + X[:, :7].cpu() * 0.5).cuda()
    got = float(H.hsic_regular(X, Y, 1.3))
    Kx = torch.exp(-torch.cdist(X.double(), X.double()) ** 2 / (2 * 1.3 ** 2))
    Ky = torch.exp(-torch.cdist(Y.double(), Y.double()) ** 2 / (2 * 1.3 ** 2))
    Hm = torch.eye(1000, dtype=torch.float64, device="cuda") - 1.0 / 1000
    ref = float(((Kx @ Hm) * (Ky @ Hm).t()).mean())
    assert abs(got - ref) < 1e-5 * abs(ref)


def test_gcn_parameterized_forward(fn):
    from copy import deepcopy
    from mcgra_b200.gcn_parameterized import PGDAttack as GP
    from mcgra_b200.models.gcn import GCN, embedding_GCN
    dev = torch.device("cuda")
    victim = GCN(nfeat=12, nclass=3, nhid=16, nlayer=2, device=dev)
    with torch.no_grad():
        victim.gc[0].weight.copy_(C(fn["gcn_W1"])); victim.gc[0].bias.copy_(C(fn["gcn_b1"]))
        victim.gc[1].weight.copy_(C(fn["gcn_W2"])); victim.gc[1].bias.copy_(C(fn["gcn_b2"]))
    for layer in victim.gc:
        layer.to(dev)
    emb = embedding_GCN(nfeat=12, nhid=16, nlayer=2, device=dev)
    emb.gc = deepcopy(victim.gc)
    gp = GP(features=C(fn["gcn_X"]), model=victim.to(dev), embedding=emb, nnodes=41, device=dev)
    close(gp.get_modified_adj(None), fn["gp_modified_adj"], rtol=1e-5, atol=1e-6)
    with pytest.raises(AttributeError):
        gp.attack()


def test_attack_helper_methods(fn):
    """Stand-alone PGDAttack helpers against the reference's own outputs."""
    from mcgra_b200.topology_attack import PGDAttack
    from helpers import Args
    dev = torch.device("cuda")
    atk = PGDAttack(model=None, embedding=None, nnodes=41, device=dev)
    atk.adj_changes.data = C(fn["pa_x"])
    close(atk.get_modified_adj(), fn["pa_expand"], rtol=0, atol=0)
    close(atk.dot_product_decode(C(fn["pa_Z"])), fn["pa_decode"], rtol=1e-5, atol=1e-6)
    for key in fn.files:
        if key.startswith("pa_decode2_"):
            _, _, ds, use = key.split("_")
            atk.args = Args()
            atk.args.dataset = ds
            atk.args.useH_A, atk.args.useY_A, atk.args.useY = use[0] == "1", use[1] == "1", use[2] == "1"
            close(atk.dot_product_decode2(C(fn["pa_Z"])), fn[key], rtol=2e-5, atol=2e-6)
    for budget, key in ((37, "pa_proj_out_37"), (100000, "pa_proj_out_big")):
        atk.adj_changes.data = C(fn["pa_proj_in"])
        atk.projection(budget)
        close(atk.adj_changes.data, fn[key], rtol=0, atol=2e-6)
