"""Probe: degenerate Polblogs-like HSIC start (identity features, x = 0): exact arithmetic moves nothing (g = +4000)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pgd_oracle as O
from helpers import run_native_case, synthetic_case
from mcgra_b200 import _native as N
from mcgra_b200 import synth
dev = torch.device("cuda:0")
W_B = {1: 0.01, 2: 0.01, 6: 10000, 7: 100, 9: 0.001, 10: 1000}
for n in (150, 700, 1490):
    for variant in ("all", "c2only", "no_c2"):
        w = dict(W_B)
        if variant == "c2only":
            w = {2: 0.01}
        if variant == "no_c2":
            w.pop(2)
        d = synthetic_case(n, 16, 2, measure="HSIC", weights=w, lr_exp=-2.5, epochs=1, dataset="polblogs", x0_scale=0.0, mean_deg=8.0)
        X = np.eye(n, dtype=np.float32)
        Wt = synth.gcn_weights(n, 16, 2, seed=15, gain=3.0)
        d["X"] = X
        d.update(Wt)
        At = torch.from_numpy(d["adj"].astype(np.float32)); Xt = torch.from_numpy(X)
        Wtt = {k: torch.from_numpy(v) for k, v in Wt.items()}
        d["feature_adj"] = O.feature_adj_of(Xt, "polblogs").numpy()
        d["H_A2"] = O.embed(Xt, At, Wtt, 2).numpy(); d["Y_A"] = O.victim(Xt, At, Wtt).numpy()
        got = run_native_case(d, trace=True)
        x1 = got["x_iters"][0]
        p64, c64 = O.problem_from_npz(d, dtype=torch.float64)
        x = torch.zeros(n * (n - 1) // 2, dtype=torch.float64, requires_grad=True)
        l, t, _ = O.iteration_terms(x, p64, c64)
        g, = torch.autograd.grad(l, x)
        g = g.numpy()
        moved = x1 > 0
        print(f"n={n} {variant}: native moved {moved.mean():.4f}; fp64 g<0 frac {(g < 0).mean():.4f}; sign mismatch {(moved != (g < 0)).mean():.4f}; "
              f"|g| min {np.abs(g).min():.3e} med {np.median(np.abs(g)):.3e}; loss native {got['loss'][0]:.6e} fp64 {float(l):.6e}")
