"""GPU parity at the BASELINE.json configuration sizes and beyond two tile rows.

  * real Cora / Citeseer / Polblogs (configs[0..1]): fixtures produced by the UNMODIFIED reference on the real datasets
    with the README all-three-priors commands (tests/golden/make_golden_real.py): trained victim, per-iteration loss
    of a 5-iteration run, final AUC / AP, sampled final scores, and the AUC of the full 100-iteration run;
  * multi-tile KL / HSIC / CKA / DP (n >= 700, 6 tile rows) against the CPU oracle run on the host;
  * the GraphMI baseline attack (MC-GRA/baseline.py) against its reference fixtures (budget-bound bisection).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import pgd_oracle as O
from helpers import run_native_case, synthetic_case, make_models

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _real_case(ds, epochs):
    """fixture dict in the attack_*.npz layout from tests/golden/real_<ds>.npz"""
    r = np.load(os.path.join(GOLDEN, f"real_{ds}.npz"))
    n = int(r["n"])
    X = sp.csr_matrix((r["feat_data"], r["feat_indices"], r["feat_indptr"]), shape=tuple(r["feat_shape"])).toarray()
    X = X.astype(np.float32)
    A = np.zeros((n, n), np.uint8)
    e = r["edges"]
    A[e[:, 0], e[:, 1]] = 1
    A[e[:, 1], e[:, 0]] = 1
    fa = O.feature_adj_of(torch.from_numpy(X), ds).numpy()
    d = dict(X=X, adj=A, labels=r["labels"], idx_attack=r["idx_attack"], feature_adj=fa, H_A2=r["H_A2"], Y_A=r["Y_A"],
             num_edges=r["num_edges"], epochs=np.int64(epochs), lr_exp=r["lr_exp"], eps=np.float64(0.0),
             weight_sup=np.float64(1.0), weights=r["weights"], measure=r["measure"], dataset=r["dataset"],
             use=np.array([True, True, True]), x0=np.zeros(n * (n - 1) // 2, np.float32),
             **{k: r[k] for k in ("W1", "b1", "W2", "b2", "Wl", "bl")})
    return r, d


# Polblogs + HSIC (README :90) starts from a degenerate point: identity features make feature_adj = 0 (c1 skipped) and at
# x = 0 every embedding row is the same bias vector, so M1 = 11^T - I and the exact gradient is +4000 for EVERY entry (no
# entry moves).  The reference's fp32 evaluation of its six n^3 centring GEMMs has errors larger than that: its own fp32
# gradient agrees in sign with its fp64 gradient on only 63 % of the entries (tests/test_oracle_golden.py::
# test_polblogs_hsic_reference_is_noise_dominated), and from iteration 1 on the loss is ~1e13 and chaotic.  Its fp32
# trajectory is therefore not a parity target; what is checked there is the loss at the native parameter against the fp64
# oracle, and that the final AUC is in the reference's range.
NOISE_DOMINATED = {"polblogs"}


@pytest.mark.parametrize("ds", ["cora", "citeseer", "polblogs"])
def test_real_dataset_short_run_matches_reference(ds):
    if not os.path.exists(os.path.join(GOLDEN, f"real_{ds}.npz")):
        pytest.skip("fixture not generated")
    r, d = _real_case(ds, int(np.load(os.path.join(GOLDEN, f"real_{ds}.npz"))["short_epochs"]))
    got = run_native_case(d, trace=(ds in NOISE_DOMINATED))
    rel = np.max(np.abs(np.asarray(got["loss"]) - r["loss_short"]) / np.abs(r["loss_short"]))
    print(f"[real] {ds} n={int(r['n'])} {str(r['measure'])}: max rel loss err over {len(got['loss'])} iterations {rel:.3e}")
    if ds in NOISE_DOMINATED:
        assert abs(got["loss"][0] - r["loss_short"][0]) <= 1e-4 * abs(r["loss_short"][0])      # same start: same loss
        prob, cfg = O.problem_from_npz(d, dtype=torch.float64)
        xs_prev = [d["x0"]] + got["x_iters"][:-1]
        forced = np.array([float(O.iteration_terms(torch.from_numpy(np.asarray(xp)).double(), prob, cfg)[0]) for xp in xs_prev])
        rel64 = np.max(np.abs(np.asarray(got["loss"]) - forced) / np.abs(forced))
        print(f"[real] {ds}: max rel loss err vs the fp64 oracle AT THE NATIVE PARAMETER {rel64:.3e}")
        assert rel64 <= 1e-4
        return
    assert rel <= (4e-4 if str(r["measure"]) == "KL" else 1e-4)
    real = d["adj"].reshape(-1).astype(np.float32)
    score = got["modified_adj"].reshape(-1)
    a, p = O.roc_auc(real, score), O.average_precision(real, score)
    print(f"[real] {ds}: AUC {a:.5f} (reference {float(r['auc_short']):.5f}), AP {p:.5f} (reference {float(r['ap_short']):.5f})")
    assert abs(a - float(r["auc_short"])) < 1e-3 and abs(p - float(r["ap_short"])) < 1e-3
    samp = got["modified_adj"][r["sample_i"], r["sample_j"]]
    np.testing.assert_allclose(samp, r["sample_short"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("ds", ["cora", "citeseer", "polblogs"])
def test_real_dataset_full_readme_run_auc(ds):
    """The README command end to end (100 iterations): final AUC / AP within 1e-3 of the reference's."""
    if not os.path.exists(os.path.join(GOLDEN, f"real_{ds}.npz")):
        pytest.skip("fixture not generated")
    r, d = _real_case(ds, int(np.load(os.path.join(GOLDEN, f"real_{ds}.npz"))["full_epochs"]))
    got = run_native_case(d, trace=False)
    real = d["adj"].reshape(-1).astype(np.float32)
    score = got["modified_adj"].reshape(-1)
    a, p = O.roc_auc(real, score), O.average_precision(real, score)
    rel_last = abs(got["loss"][-1] - r["loss_full"][-1]) / abs(r["loss_full"][-1])
    print(f"[real-full] {ds}: AUC {a:.5f} (reference {float(r['auc_full']):.5f}), AP {p:.5f} "
          f"(reference {float(r['ap_full']):.5f}), last-iteration loss rel err {rel_last:.2e}")
    if ds in NOISE_DOMINATED:       # chaotic trajectory (see above): the attack must still recover the graph as well
        assert a > float(r["auc_full"]) - 0.02
        return
    # 100 free-running Adam iterations are chaotic at the 1e-7 level: two runs of the SAME binary (float atomics in the
    # degree sums) end with 99.8 % of the entries differing by > 1e-4, AUC scattering by ~1e-3 and AP (a ~0.03 quantity
    # carried by a few hundred top-ranked pairs) by ~2e-3, while the loss agrees to 2e-5
    # (profiles/r02_cora100_run_to_run.txt).  The reference's value is one sample of that family, so the full run is held
    # to the loss (1e-4) and to 2.5e-3 / 4e-3 on AUC / AP; the 1e-3 bar is enforced on the 5-iteration run above.
    assert rel_last < 1e-4
    assert abs(a - float(r["auc_full"])) < 2.5e-3 and abs(p - float(r["ap_full"])) < 4e-3


MULTI = [
    ("KL", {1: 100, 2: 1e-4, 6: 1e-3, 9: 1000, 10: 1e-3}, -1.5, "citeseer"),                 # README Citeseer (Profile C)
    ("HSIC", {1: 0.01, 2: 0.01, 6: 10000, 7: 100, 9: 0.001, 10: 1000}, -2.5, "cora"),          # Profile B, c1 active
    ("CKA", {1: 0.01, 2: 0.01, 6: 100, 7: 1.0, 9: 1.0, 10: 1.0}, -2.0, "cora"),
    ("DP", {1: 1e-3, 2: 1e-3, 6: 10, 7: 1.0, 9: 0.1, 10: 1.0}, -2.0, "cora"),
    ("KDE", {1: 1000, 2: 500, 6: 0.01, 7: 1.0, 9: 50.0, 10: 20.0}, -2.0, "cora"),     # utils.MutualInformation on every term
]


@pytest.mark.parametrize("measure,weights,lr_exp,dataset", MULTI)
def test_multi_tile_measures_match_oracle(measure, weights, lr_exp, dataset):
    """6 tile rows (n = 700): the tiled KL row statistics and the dense contraction path against the oracle, with the
    loss of every iteration re-evaluated by the fp64 oracle at the native parameter when the fp32 trajectories separate."""
    d = synthetic_case(700, 40, 5, measure=measure, weights=weights, lr_exp=lr_exp, epochs=3, dataset=dataset, mean_deg=8.0)
    got = run_native_case(d)
    prob, cfg = O.problem_from_npz(d, dtype=torch.float64)
    xs_prev = [d["x0"]] + got["x_iters"][:-1]
    forced = np.array([float(O.iteration_terms(torch.from_numpy(np.asarray(xp)).double(), prob, cfg)[0]) for xp in xs_prev])
    rel = np.max(np.abs(np.asarray(got["loss"]) - forced) / np.abs(forced))
    print(f"[multi-tile] {measure} n=700: max rel loss err vs fp64 oracle at the native parameter {rel:.3e}")
    assert rel <= (4e-4 if measure == "KL" else 1e-4)
    # one full step against the fp32 oracle from the same start: parameter after the first iteration
    prob32, cfg32 = O.problem_from_npz(d)
    ref = O.attack(prob32, cfg32, 1, x0=torch.from_numpy(d["x0"]))
    dx = np.abs(got["x_iters"][0] - ref["x_iters"][0].numpy())
    print(f"[multi-tile] {measure}: after 1 iteration fraction |dx| > 2e-4 = {np.mean(dx > 2e-4):.2e}")
    # CKA's gradient is a difference of two normalised contractions: more entries sit at the fp32 noise floor where Adam's
    # first step (+-lr by the SIGN of the gradient) differs between any two fp32 evaluations
    assert np.mean(dx > 2e-4) < (0.03 if measure == "CKA" else 0.01)


@pytest.mark.parametrize("case", ["budget_n150", "free_n90"])
def test_graphmi_baseline_matches_reference(case):
    """mcgra_b200.baseline.PGDAttack against the unmodified reference's MC-GRA/baseline.py (fixtures baseline_*.npz)."""
    from mcgra_b200.baseline import PGDAttack
    d = np.load(os.path.join(GOLDEN, f"baseline_{case}.npz"))
    dev = torch.device("cuda:0")
    n = int(d["labels"].shape[0])
    victim, emb = make_models(d, dev)
    model = PGDAttack(model=victim, embedding=emb, nnodes=n, loss_type="CE", device=dev).to(dev)
    out = model.attack(None, float(d["lr"]), 0, 1.0, None, None, 0, 0, 0, None, None, None,
                       torch.from_numpy(d["adj"].astype(np.float32)), d["X"], np.zeros((n, n), np.float32), d["labels"],
                       d["idx_attack"], int(d["num_edges"]), 0, epochs=int(d["epochs"]), _trace=True)
    torch.cuda.synchronize()
    loss = model.engine.losses()["loss"]
    # ~2 % of the entries have |gradient| at the fp32 noise floor (~1e-8, Adam's eps): their first steps differ in sign
    # between any two fp32 evaluations (the trajectory then differs by ~5e-4 in the loss).  Per-iteration loss against
    # the fp64 oracle at the native parameter, x compared robustly -- the same protocol as tests/test_gpu_attack.py
    t64 = lambda a: torch.from_numpy(np.asarray(a)).double()
    prob = dict(n=n, X=t64(d["X"]), labels=torch.from_numpy(d["labels"]).long(),
                idx_attack=torch.from_numpy(d["idx_attack"]).long(), W={k: t64(d[k]) for k in ("W1", "b1", "W2", "b2", "Wl", "bl")})
    xs_prev = [np.zeros(n * (n - 1) // 2)] + [x.cpu().numpy() for x in model._trace[:-1]]
    forced = np.array([float(O.baseline_loss_at(t64(xp), prob)) for xp in xs_prev])
    np.testing.assert_allclose(loss, forced, rtol=1e-4)
    assert abs(loss[0] - d["loss"][0]) <= 1e-5 * abs(d["loss"][0])
    for k, xk in enumerate(model._trace):
        dx = np.abs(xk.cpu().numpy() - d["x_iters"][k])
        # (with a binding budget a noise-floor difference also moves the bisection root mu, i.e. every entry a little)
        thr = 1e-3 if case.startswith("budget") else 2e-4
        assert np.mean(dx > thr) < 0.03 and np.max(dx) <= 2.5 * float(d["lr"]) * (k + 1), f"x after iteration {k}"
    np.testing.assert_allclose(loss, d["loss"], rtol=2e-3)
    if case.startswith("budget"):
        assert abs(float(model._trace[-1].sum()) - float(d["num_edges"])) < 0.05 * float(d["num_edges"])
    assert np.mean(np.abs(model.adj_changes.data.cpu().numpy() - d["x_final"]) > 2e-3) < 0.01
    assert np.mean(np.abs(out.cpu().numpy() - d["output"]) > 2e-3) < 0.01
    A = torch.from_numpy(d["adj"].astype(np.float32)).to(dev)
    sm = float(model.feature_smoothing(A, torch.from_numpy(d["X"]).to(dev)))
    assert abs(sm - float(d["smooth_true_adj"])) <= 1e-4 * abs(float(d["smooth_true_adj"]))


def test_class_surface_methods():
    """bisection / adding_noise / calc_kl / dot_product / delete_eye / test (topology_attack.py:83-93, 397-412, 469-487)."""
    from mcgra_b200.topology_attack import PGDAttack
    fn = np.load(os.path.join(GOLDEN, "functions.npz"))
    dev = torch.device("cuda:0")
    atk = PGDAttack(model=None, embedding=None, nnodes=41, device=dev)
    x = torch.from_numpy(fn["pa_proj_in"]).to(dev)
    atk.adj_changes.data = x.clone()
    mu = atk.bisection(float((x - 1).min()), float(x.max()), 37, 1e-5)
    ref_mu = O.bisection(torch.from_numpy(fn["pa_proj_in"]), torch.from_numpy(fn["pa_proj_in"] - 1).min(),
                         torch.from_numpy(fn["pa_proj_in"]).max(), 37, 1e-5)
    assert abs(float(mu) - float(ref_mu)) < 1e-6
    np.testing.assert_allclose(torch.clamp(x - mu, 0, 1).cpu().numpy(), fn["pa_proj_out_37"], atol=2e-6)
    M = torch.from_numpy(fn["norm_in"]).to(dev)
    kl = atk.calc_kl(M, (M.t() * 0.5 + 0.1).contiguous())
    assert abs(float(kl) - float(fn["pa_kl"])) <= 2e-5 * abs(float(fn["pa_kl"])) + 1e-7
    Xh = torch.from_numpy(fn["hs_X"]).to(dev)
    dp = atk.dot_product(Xh, Xh * 0.3 + 1)
    assert abs(float(dp) - float(fn["pa_dp"])) <= 2e-5 * abs(float(fn["pa_dp"]))
    Wd = torch.rand(300, 200, device=dev)       # wide operands: the tcgen05 contraction path
    Yd = torch.rand(300, 180, device=dev)
    want = float(torch.norm(Yd.double().t() @ Wd.double()))
    assert abs(float(atk.dot_product(Wd, Yd)) - want) <= 2e-5 * want
    torch.manual_seed(3)
    Mn = torch.rand(41, 41, device=dev)
    keep = Mn.clone()
    torch.manual_seed(5)
    out = atk.adding_noise(Mn, 0.05)
    torch.manual_seed(5)
    want = torch.clamp(keep + torch.randn_like(keep) * 0.05, 0, 1)
    assert torch.allclose(out, want, atol=1e-7) and out.data_ptr() == Mn.data_ptr()
    assert atk.delete_eye(torch.ones(41, 41, device=dev)) is None


def test_mcgpb_graphmi_attack_matches_reference():
    """mcgra_b200.mcgpb_attack.PGDAttack against the unmodified defence repo's attack (MC-GPB/topology_attack.py:36-87):
    plain gradient descent, feature smoothing from iteration 50, un-normalised final decode."""
    from mcgra_b200.mcgpb_attack import PGDAttack
    d = np.load(os.path.join(GOLDEN, "mcgpb_attack_n150.npz"))
    dev = torch.device("cuda:0")
    n = int(d["labels"].shape[0])
    victim, emb = make_models(d, dev)
    model = PGDAttack(model=victim, embedding=emb, nnodes=n, loss_type="CE", device=dev).to(dev)
    out = model.attack(d["X"], np.zeros((n, n), np.float32), d["labels"], d["idx_attack"], int(d["num_edges"]),
                       epochs=int(d["epochs"]), _trace=True)
    torch.cuda.synchronize()
    loss = model.engine.losses()["loss"]
    rel = np.max(np.abs(loss - d["loss"]) / np.abs(d["loss"]))
    print(f"[mcgpb] max rel loss err vs the reference over {len(loss)} iterations (smoothing from 50): {rel:.3e}")
    # same protocol as the GraphMI baseline above: per-iteration loss against the fp64 oracle at the native parameter
    t64 = lambda a: torch.from_numpy(np.asarray(a)).double()
    prob = dict(n=n, X=t64(d["X"]), labels=torch.from_numpy(d["labels"]).long(),
                idx_attack=torch.from_numpy(d["idx_attack"]).long(), W={k: t64(d[k]) for k in ("W1", "b1", "W2", "b2", "Wl", "bl")})
    xs_prev = [np.zeros(n * (n - 1) // 2)] + [x.cpu().numpy() for x in model._trace[:-1]]
    forced = np.array([float(O.baseline_loss_at(t64(xp), prob, smooth_coef=1e-4 if t >= 50 else 0.0))
                       for t, xp in enumerate(xs_prev)])
    rel64 = np.max(np.abs(loss - forced) / np.abs(forced))
    print(f"[mcgpb] max rel loss err vs the fp64 oracle at the native parameter: {rel64:.3e}")
    assert rel64 < 1e-4
    assert rel < 2e-3
    for k, it in enumerate(d["x_keep_idx"]):
        dx = np.abs(model._trace[int(it)].cpu().numpy() - d["x_keep"][k])
        assert np.mean(dx > 2e-4) < 0.03, f"x after iteration {int(it)}"
    assert np.mean(np.abs(model.adj_changes.data.cpu().numpy() - d["x_final"]) > 2e-3) < 0.01
    assert np.mean(np.abs(out.cpu().numpy() - d["output"]) > 2e-3) < 0.01
