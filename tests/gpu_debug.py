"""Stage-by-stage comparison of the native path with the oracle on one golden case (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pgd_oracle as O
from helpers import run_native_case


def md(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))), float(np.max(np.abs(b)))


def main(case):
    d = np.load(os.path.join(ROOT, "tests", "golden", f"attack_{case}.npz"))
    prob, cfg = O.problem_from_npz(d)
    n = prob["n"]
    x0 = torch.from_numpy(d["x0"])
    # oracle: first iteration internals
    xr = x0.clone().requires_grad_(True)
    loss, terms, A_hat = O.iteration_terms(xr, prob, cfg)
    loss.backward()
    g = xr.grad.numpy()
    M = torch.clamp(O.expand(x0, n), 0, 1)
    deg = (M + torch.eye(n)).sum(1)
    r = deg.pow(-0.5)
    S1 = prob["X"] @ prob["W"]["W1"]
    got = run_native_case(d, epochs=1)
    eng = got["model"].engine
    print(f"== case {case} n={n} measure={cfg['measure']}")
    print("r        ", md(eng.r.cpu(), r))
    print("Y1[:, :16] vs M(r*S1)", md(eng.Y1[:, :16].cpu(), M @ (r[:, None] * S1)))
    print("Y1[:, 16:] vs M S1   ", md(eng.Y1[:, 16:].cpu(), M @ S1))
    H1 = torch.relu(A_hat.detach() @ S1 + prob["W"]["b1"])
    S2 = H1 @ prob["W"]["W2"]
    print("S2       ", md(eng.S2.cpu(), S2))
    H2 = torch.relu(A_hat.detach() @ S2 + prob["W"]["b2"])
    print("H2       ", md(eng.H2.cpu(), H2))
    em = O.embed(prob["X"], M, prob["W"], 2)
    zh = torch.nn.functional.normalize(em, p=2, dim=1)
    print("zhat     ", md(eng.zhat.cpu(), zh))
    L = got["terms"]
    for k in ("origin", "c1", "c2", "c6", "c7", "c9", "c10"):
        ref = float(terms[k]) if k in terms else 0.0
        sg = -1.0 if (cfg["measure"] == "HSIC" and k in ("c1", "c2")) else 1.0
        print(f"term {k:7s} native {L[k][0]: .8e}  oracle {ref: .8e}")
    print(f"loss native {L['loss'][0]:.8e} oracle {float(loss):.8e} golden {d['loss'][0]:.8e}")
    # gradient: after one Adam step from zero moments m = (1-beta1) g
    T = eng.T
    from mcgra_b200 import _native as N
    mp = torch.zeros(eng.P, device=eng.dev)
    N.call("mcgra_tiles_to_tril", N.ptr(eng.mt), n, 0, T, None, 1, N.ptr(mp), N.stream_ptr())
    gn = (mp / 0.1).cpu().numpy()
    print("grad     ", md(gn, g), " rel-to-max", md(gn, g)[0] / (np.abs(g).max() + 1e-30))
    worst = np.argsort(-np.abs(gn - g))[:5]
    print("  worst idx", worst, "native", gn[worst], "oracle", g[worst])
    print("x after 1", md(got["x_iters"][0], d["x_iters"][0]))


if __name__ == "__main__":
    for c in sys.argv[1:] or ["mse_A_n150"]:
        main(c)
