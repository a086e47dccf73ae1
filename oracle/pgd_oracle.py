"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MC-GRA PGD attack hot path.

A from-scratch restatement (torch CPU ops + autograd on dense n x n tensors, exactly the way the
reference shapes the work, including its dense-diagonal normalisation GEMMs and dense centring) of
`/root/reference/MC-GRA/topology_attack.py` `PGDAttack.attack` and the helpers it calls.  Every
function cites the reference lines it follows.  Parity is PINNED: tests/test_oracle_golden.py checks
this file against fixtures produced by the unmodified reference (tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (mc-gra_b200/) never does: it fails loudly without its CUDA library.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

ALIGN = {"c1": 100, "c2": 1000, "c6": 10, "c7": 10, "c9": 1, "c10": 1}      # utils.py:1100-1111


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def tril_pairs(n):
    """Row-major strict-lower-triangle index pairs: k = i(i-1)/2 + j (topology_attack.py:372-374)."""
    return torch.tril_indices(n, n, -1)


def expand(x, n):
    """get_modified_adj with ori_adj = 0 (topology_attack.py:365-379): symmetric, zero diagonal."""
    idx = tril_pairs(n)
    m = torch.zeros(n, n, dtype=x.dtype)
    m = m.index_put((idx[0], idx[1]), x)
    m = m + m.t()
    return (torch.ones(n, n, dtype=x.dtype) - torch.eye(n, dtype=x.dtype)) * m


def add_noise(M, eps, noise=None):
    """adding_noise (topology_attack.py:474-478).  `noise` lets a test inject the N(0,1) draw."""
    if noise is None:
        noise = torch.randn_like(M)
    return torch.clamp(M + noise * eps, min=0, max=1)


def normalize(M):
    """utils.normalize_adj_tensor dense branch (utils.py:221-229): D^-1/2 (M+I) D^-1/2 by two
    dense-diagonal matmuls, inf -> 0."""
    mx = M + torch.eye(M.shape[0], dtype=M.dtype)
    r = mx.sum(1).pow(-0.5).flatten()
    r = torch.where(torch.isinf(r), torch.zeros_like(r), r)
    R = torch.diag(r)
    return (R @ mx) @ R


def gc_layer(inp, adj, W, b):
    """GraphConvolution.forward (models/gcn.py:35-46)."""
    return adj @ (inp @ W) + b


def embed(X, adj, Wt, nlayer=2):
    """embedding_GCN.forward (models/gcn.py:71-76): relu(GC) stack of the victim's layers."""
    h = X
    layers = [(Wt["W1"], Wt["b1"]), (Wt["W2"], Wt["b2"])]
    for W, b in layers[:nlayer]:
        h = torch.relu(gc_layer(h, adj, W, b))
    return h


def victim(X, adj, Wt):
    """GCN.forward in eval mode (models/gcn.py:164-174): relu-GC x2, Linear, log_softmax."""
    h = embed(X, adj, Wt, 2)
    return F.log_softmax(h @ Wt["Wl"].t() + Wt["bl"], dim=1)


def decode_tril(Z):
    """PGDAttack.dot_product_decode (topology_attack.py:414-419)."""
    n = Z.shape[0]
    Zn = F.normalize(Z, p=2, dim=1)
    G = torch.relu(Zn @ Zn.t())
    idx = tril_pairs(n)
    return G[idx[0], idx[1]]


def decode2(Z, dataset, useH_A, useY_A, useY):
    """PGDAttack.dot_product_decode2 (topology_attack.py:421-467)."""
    n = Z.shape[0]
    eye = torch.eye(n, dtype=Z.dtype)
    if dataset in ("cora", "AIDS"):
        return torch.sigmoid(torch.relu(Z @ Z.t() - eye))
    if dataset == "citeseer":
        Zn = F.normalize(Z, p=2, dim=1)
        return torch.sigmoid(torch.relu(Zn @ Zn.t() - eye))
    if dataset == "brazil":
        return torch.relu(Z @ Z.t() - eye)
    if dataset in ("polblogs", "usair"):
        if dataset == "polblogs" and useH_A and useY_A and useY:
            G = F.normalize(Z @ Z.t(), p=2, dim=1)
        elif dataset == "usair" and useY and not useH_A and not useY_A:
            Zn = F.normalize(Z, p=3, dim=1)
            G = Zn @ Zn.t()
        elif dataset == "usair" and not useY and useH_A and useY_A:
            Zn = F.normalize(Z, p=2, dim=1)
            G = Zn @ Zn.t()
        elif dataset == "usair" and useY and useH_A and not useY_A:
            Zn = F.normalize(Z, p=5, dim=1)
            G = Zn @ Zn.t()
        else:
            G = F.normalize(Z @ Z.t(), p=2, dim=1)
        return torch.relu(G - eye)
    raise ValueError("dataset %r has no dot_product_decode2 branch in the reference" % dataset)


def feature_adj_of(X, dataset):
    """main.dot_product_decode (main.py:44-55)."""
    eye = torch.eye(X.shape[0], dtype=X.dtype)
    if dataset in ("cora", "citeseer", "AIDS"):
        return torch.sigmoid(torch.relu(X @ X.t() - eye))
    Xn = F.normalize(X, p=2, dim=1)
    return torch.relu(Xn @ Xn.t() - eye)


def info_entropy(p):
    """Info_entropy (topology_attack.py:44-47)."""
    q = torch.clamp(p, 1e-4, 1 - 1e-4)
    return -torch.mean(q * torch.log2(q))


# ----------------------------------------------------------------------------------------------
# measures (topology_attack.py:190-208)
# ----------------------------------------------------------------------------------------------
def centering(K):
    """CudaCKA.centering (utils.py:1060-1065): dense H K H."""
    n = K.shape[0]
    H = torch.eye(n, dtype=K.dtype) - torch.ones(n, n, dtype=K.dtype) / n
    return (H @ K) @ H


def linear_hsic(X, Y):
    """CudaCKA.linear_HSIC (utils.py:1080-1084)."""
    return torch.sum(centering(X @ X.t()) * centering(Y @ Y.t()))


def linear_cka(X, Y):
    """CudaCKA.linear_CKA (utils.py:1086-1091)."""
    return linear_hsic(X, Y) / (torch.sqrt(linear_hsic(X, X)) * torch.sqrt(linear_hsic(Y, Y)))


def rbf(X, sigma=None):
    """CudaCKA.rbf (utils.py:1067-1075): sigma = sqrt(median of the non-zero entries)."""
    G = X @ X.t()
    K = torch.diag(G) - G + (torch.diag(G) - G).t()
    if sigma is None:
        sigma = math.sqrt(torch.median(K[K != 0]).item())
    return torch.exp(K * (-0.5 / (sigma * sigma)))


def kernel_hsic(X, Y, sigma=None):
    """CudaCKA.kernel_HSIC (utils.py:1077-1078)."""
    return torch.sum(centering(rbf(X, sigma)) * centering(rbf(Y, sigma)))


def kernel_cka(X, Y, sigma=None):
    """CudaCKA.kernel_CKA (utils.py:1093-1097)."""
    return kernel_hsic(X, Y, sigma) / (torch.sqrt(kernel_hsic(X, X, sigma)) * torch.sqrt(kernel_hsic(Y, Y, sigma)))


def pairwise_sqdist(x):
    """utils.pairwise_distances (utils.py:803-806)."""
    nrm = torch.sum(x ** 2, -1).reshape(-1, 1)
    return -2 * (x @ x.t()) + nrm + nrm.t()


def gaussian_hsic(x, y, s_x=1, s_y=1):
    """utils.HSIC (utils.py:814-822): tr(L H K H)/(m-1)^2, K = exp(-dist/s)."""
    m = x.shape[0]
    K = torch.exp(-pairwise_sqdist(x) / s_x)
    L = torch.exp(-pairwise_sqdist(y) / s_y)
    H = torch.eye(m, dtype=x.dtype) - torch.ones(m, m, dtype=x.dtype) / m
    return torch.trace(L @ (H @ (K @ H))) / ((m - 1) ** 2)


def hs_distmat(X):
    """hsic.distmat (hsic.py:20-27)."""
    r = torch.sum(X * X, 1).view(-1, 1)
    a = X @ X.t()
    return r.expand_as(a) - 2 * a + r.t().expand_as(a)


def hs_sigma_estimation(X, Y):
    """hsic.sigma_estimation (hsic.py:5-17): median of the strict lower triangle of distmat([X;Y])."""
    D = hs_distmat(torch.cat([X, Y])).detach().cpu().numpy()
    tri = D[np.tril_indices(D.shape[0], -1)]
    med = np.median(tri)
    if med <= 0:
        med = np.mean(tri)
    if med < 1e-2:
        med = 1e-2
    return med


def hs_kernelmat(X, sigma):
    """hsic.kernelmat (hsic.py:30-47): exp(-D/(2 s^2)) @ H (column centring only)."""
    m = X.shape[0]
    H = torch.eye(m) - torch.ones(m, m) / m
    D = hs_distmat(X)
    s = sigma if sigma else hs_sigma_estimation(X, X)
    K = torch.exp(-D / (2.0 * s * s)).float()
    return K @ H


def hs_hsic_regular(x, y, sigma=None):
    """hsic.hsic_regular (hsic.py:117-124)."""
    return torch.mean(hs_kernelmat(x, sigma) * hs_kernelmat(y, sigma).t())


def hs_hsic_normalized(x, y, sigma=None):
    """hsic.hsic_normalized (hsic.py:127-135)."""
    return hs_hsic_regular(x, y, sigma) / (torch.sqrt(hs_hsic_regular(x, x, sigma)) * torch.sqrt(hs_hsic_regular(y, y, sigma)))


def hs_distcorr(X, sigma=1.0):
    """hsic.distcorr (hsic.py:50-53)."""
    return torch.mean(torch.exp(-hs_distmat(X) / (2.0 * sigma * sigma)))


def hs_compute_kernel(x, y):
    """hsic.compute_kernel (hsic.py:56-66)."""
    d = x.shape[1]
    diff = x.unsqueeze(1) - y.unsqueeze(0)
    return torch.exp(-diff.pow(2).mean(2) / float(d))


def hs_mmd(x, y, sigma=None):
    """hsic.mmd (hsic.py:69-90)."""
    Dxx, Dyy = hs_distmat(x), hs_distmat(y)
    if sigma:
        Kx, Ky, sxy = torch.exp(-Dxx / (2.0 * sigma * sigma)), torch.exp(-Dyy / (2.0 * sigma * sigma)), sigma
    else:
        sx, sy, sxy = hs_sigma_estimation(x, x), hs_sigma_estimation(y, y), hs_sigma_estimation(x, y)
        Kx, Ky = torch.exp(-Dxx / (2.0 * sx * sx)), torch.exp(-Dyy / (2.0 * sy * sy))
    Dxy = hs_distmat(torch.cat([x, y]))[: x.shape[0], x.shape[0]:]
    Kxy = torch.exp(-Dxy / (1.0 * sxy * sxy))
    return torch.mean(Kx) + torch.mean(Ky) - 2 * torch.mean(Kxy)


def hs_mmd_pxpy_pxy(x, y, sigma=None):
    """hsic.mmd_pxpy_pxy (hsic.py:93-114)."""
    Dxx, Dyy = hs_distmat(x), hs_distmat(y)
    if sigma:
        Kx, Ky = torch.exp(-Dxx / (2.0 * sigma * sigma)), torch.exp(-Dyy / (2.0 * sigma * sigma))
    else:
        sx, sy = hs_sigma_estimation(x, x), hs_sigma_estimation(y, y)
        Kx, Ky = torch.exp(-Dxx / (2.0 * sx * sx)), torch.exp(-Dyy / (2.0 * sy * sy))
    A = torch.mean(Kx * Ky)
    B = torch.mean(torch.mean(Kx, dim=0) * torch.mean(Ky, dim=0))
    C = torch.mean(Kx) * torch.mean(Ky)
    return A - 2 * B + C


def calc_kl(X, Y):
    """PGDAttack.calc_kl (topology_attack.py:483-487); implicit softmax dim of a 2-D tensor is 1."""
    return F.kl_div(F.log_softmax(Y, dim=1), F.softmax(X, dim=1), reduction="batchmean")


def dot_product(X, Y):
    """PGDAttack.dot_product (topology_attack.py:480-481)."""
    return torch.norm(Y.t() @ X, p=2)


def kde_mi(X, Y, num_bins, sigma=0.4):
    """utils.MutualInformation(sigma, num_bins, normalize=True).forward on 2-D inputs (utils.py:980-1053): the inputs
    broadcast to [1, m, B] against bins = linspace(0, B, B); returns the scalar the reference indexes with [0]."""
    eps = 1e-10
    s = 2 * sigma ** 2
    bins = torch.linspace(0, num_bins, num_bins, dtype=X.dtype)

    def marginal(v):
        kv = torch.exp(-0.5 * ((v - bins[None, None, :]) / s).pow(2))          # [1, m, B]
        pdf = kv.mean(dim=1)
        return pdf / (pdf.sum(dim=1).unsqueeze(1) + eps), kv
    p1, k1 = marginal(X)
    p2, k2 = marginal(Y)
    joint = torch.matmul(k1.transpose(1, 2), k2)
    pj = joint / (joint.sum(dim=(1, 2)).view(-1, 1, 1) + eps)
    H1 = -torch.sum(p1 * torch.log2(p1 + eps), dim=1)
    H2 = -torch.sum(p2 * torch.log2(p2 + eps), dim=1)
    H12 = -torch.sum(pj * torch.log2(pj + eps), dim=(1, 2))
    mi = H1 + H2 - H12
    return (2 * mi / (H1 + H2))[0]


def pick_measure(name):
    if name == "HSIC":
        return linear_hsic
    if name == "MSELoss":
        return lambda a, b: F.mse_loss(a, b)
    if name == "KL":
        return calc_kl
    if name == "CKA":
        return linear_cka
    if name == "DP":
        return dot_product
    if name == "KDE":            # topology_attack.py:199-201: num_bins = feature_adj.shape[0] for the n x n terms
        return lambda a, b: kde_mi(a, b, a.shape[0])
    raise NotImplementedError("measure %s" % name)


# ----------------------------------------------------------------------------------------------
# projection (topology_attack.py:338-347, 397-412)
# ----------------------------------------------------------------------------------------------
def bisection(x, a, b, budget, epsilon=1e-5):
    def func(mu):
        return torch.clamp(x - mu, 0, 1).sum() - budget

    miu = a
    while (b - a) >= epsilon:
        miu = (a + b) / 2
        if func(miu) == 0.0:
            break
        if func(miu) * func(a) < 0:
            b = miu
        else:
            a = miu
    return miu


def projection(x, budget):
    if torch.clamp(x, 0, 1).sum() > budget:
        left = (x - 1).min()
        right = x.max()
        miu = bisection(x, left, right, budget, 1e-5)
        return torch.clamp(x - miu, min=0, max=1)
    return torch.clamp(x, min=0, max=1)


# ----------------------------------------------------------------------------------------------
# the loop (topology_attack.py:95-324)
# ----------------------------------------------------------------------------------------------
def iteration_terms(x, prob, cfg, noise=None):
    """Forward of one PGD iteration, steps 2-15 of SURVEY 3.2 (topology_attack.py:164-272).
    Returns (loss, dict of terms, adj_norm)."""
    n = prob["n"]
    X, Wt, labels, idx = prob["X"], prob["W"], prob["labels"], prob["idx_attack"]
    adj_true, Fadj = prob["adj"], prob["feature_adj"]
    w = cfg["weights"]
    w1, w2, w6, w7, w9, w10 = w[0], w[1], w[5], w[6], w[8], w[9]
    measure = cfg["measure"]
    calc = pick_measure(measure)
    sgn = -1.0 if measure == "HSIC" else 1.0

    M = expand(x, n)                                                   # :164
    M = add_noise(M, cfg.get("eps", 0.0), noise) if (cfg.get("eps", 0.0) != 0 or noise is not None) \
        else torch.clamp(M, min=0, max=1)                              # :165 (eps=0: noise*0, clamp only)
    A_hat = normalize(M)                                               # :166
    output = victim(X, A_hat, Wt)                                      # :167
    origin = F.nll_loss(output[idx], labels[idx]) + torch.norm(x, p=2) * 0.001     # :172-173
    loss = cfg.get("weight_sup", 1.0) * origin                        # :175
    terms = {"origin": origin}

    H_A2 = embed(X, adj_true, Wt, 2)                                   # :179-180 (raw true adjacency)
    Y_A = victim(X, adj_true, Wt)                                      # :182
    em = embed(X, M, Wt, 2)                                            # :185 (raw M)
    after = decode_tril(em)                                            # :187
    M1 = expand(after, n)                                              # :188

    if w1 != 0 and Fadj.max() != Fadj.min():                           # :212-220
        c1 = w1 * calc(Fadj, A_hat) * 1000 * ALIGN["c1"]
        loss = loss + sgn * c1
        terms["c1"] = c1
    if w2 != 0:                                                        # :221-229
        c2 = w2 * calc(A_hat, M1) * 100 * ALIGN["c2"]
        loss = loss + sgn * c2
        terms["c2"] = c2
    if w6 != 0:                                                        # :230-232
        c6 = w6 * info_entropy(A_hat) * 100 * ALIGN["c6"]
        loss = loss + c6
        terms["c6"] = c6
    if w7 != 0:                                                        # :233-236
        c7 = w7 * info_entropy(M1) * ALIGN["c7"]
        loss = loss + c7
        terms["c7"] = c7
    if w9 != 0:                                                        # :237-258 (em_cur == em, H_A_cur == H_A2)
        calc9 = (lambda a, b: kde_mi(a, b, H_A2.shape[1])) if measure == "KDE" else calc       # :246-250
        c9 = sgn * w9 * calc9(H_A2[idx], em[idx]) * ALIGN["c9"]
        loss = loss + c9
        terms["c9"] = c9
    if w10 != 0:                                                       # :259-272
        output2 = F.log_softmax(em @ Wt["Wl"].t() + Wt["bl"], dim=1)   # victim(X, M): its hidden state == em
        calc10 = (lambda a, b: kde_mi(a, b, Y_A.shape[1])) if measure == "KDE" else calc      # :261-265
        c10 = sgn * w10 * calc10(Y_A[idx], torch.softmax(output2[idx], dim=1)) * ALIGN["c10"]
        loss = loss + c10
        terms["c10"] = c10
    return loss, terms, A_hat


def adam_update(x, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor step (topology_attack.py:121,279), t = 1-based step count."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    x.addcdiv_(m, denom, value=-(lr / bc1))


def finalize(prob, cfg, A_hat_last):
    """topology_attack.py:300-322."""
    n = prob["n"]
    X, Wt = prob["X"], prob["W"]
    ds = cfg["dataset"]
    use = cfg["use"]
    em = embed(X, A_hat_last, Wt, 2)
    x_final = decode_tril(em)
    mod = expand(x_final, n)
    H1 = embed(X, mod, Wt, 1)
    H2 = embed(X, mod, Wt, 2)
    Y2 = victim(X, mod, Wt)
    d2 = lambda Z: decode2(Z, ds, *use)
    cur = mod + d2(H1) + d2(H2) + prob["feature_adj"] + d2(Y2)
    if use[0]:
        cur = cur + d2(prob["H_A"])
    if use[1]:
        cur = cur + d2(prob["Y_A"])
    if use[2]:
        cur = cur + (prob["labels"][:, None] == prob["labels"][None, :]).to(cur.dtype)   # saved_data/<ds>.npy
    return x_final, cur


def attack(prob, cfg, epochs, x0=None, bookkeeping=False, record_terms=False):
    """Whole PGDAttack.attack.  prob: n, X, adj, labels, idx_attack, feature_adj, W(dict), H_A, Y_A.
    cfg: measure, weights(10), lr, eps, weight_sup, dataset, use(3), num_edges.
    bookkeeping=True additionally executes the reference's per-iteration bookkeeping (:285-296), which has
    no effect on results but is inside the reference's timed loop."""
    n = prob["n"]
    dt = prob["X"].dtype
    P = n * (n - 1) // 2
    x = torch.zeros(P, dtype=dt) if x0 is None else x0.clone().to(dt)
    m = torch.zeros_like(x)
    v = torch.zeros_like(x)
    losses, xs, terms_all = [], [], []
    A_hat = normalize(expand(x, n))                                    # :141-142
    for t in range(epochs):
        xr = x.clone().requires_grad_(True)
        loss, terms, A_hat = iteration_terms(xr, prob, cfg)
        loss.backward()
        losses.append(float(loss.detach().double()))
        if record_terms:
            terms_all.append({k: float(vv.detach().double()) for k, vv in terms.items()})
        adam_update(x, xr.grad, m, v, t + 1, cfg["lr"])
        x = projection(x, cfg["num_edges"])                            # :281
        x = torch.clamp(x, 0, 1)                                       # :282-283
        xs.append(x.clone())
        A_hat = A_hat.detach()
        if bookkeeping:                                                # :285-296
            with torch.no_grad():
                em = embed(prob["X"], A_hat, prob["W"], 2)
                decode_tril(em)
                Mb = expand(x, n)
                float(Mb.mean())
                out2 = victim(prob["X"], normalize(Mb), prob["W"])
                float((out2.argmax(1) == prob["labels"]).double().mean())
    x_final, cur = finalize(prob, cfg, A_hat)
    return {"loss": losses, "x_iters": xs, "x_final": x_final, "modified_adj": cur, "terms": terms_all}


# ----------------------------------------------------------------------------------------------
# GraphMI baseline attack (MC-GRA/baseline.py:36-86) and its feature-smoothing helper (:155-169)
# ----------------------------------------------------------------------------------------------
def feature_smoothing(adj, X):
    """baseline.PGDAttack.feature_smoothing (baseline.py:155-169): tr(X^T L~ X) through dense diagonal matmuls."""
    rowsum = adj.sum(1)
    r_inv = rowsum.flatten()
    L = torch.diag(r_inv) - adj
    r_inv = (r_inv + 1e-3).pow(-1 / 2).flatten()
    r_inv = torch.where(torch.isinf(r_inv), torch.zeros_like(r_inv), r_inv)
    R = torch.diag(r_inv)
    L = (R @ L) @ R
    return torch.trace((X.t() @ L) @ X)


def baseline_loss_at(x, prob, smooth_coef=0.0):
    """Loss of the GraphMI baselines at a given parameter (baseline.py:53-58; MC-GPB/topology_attack.py:51-61)."""
    n = prob["n"]
    M = expand(x, n)
    out = victim(prob["X"], normalize(M), prob["W"])
    loss = F.nll_loss(out[prob["idx_attack"]], prob["labels"][prob["idx_attack"]]) + torch.norm(x, p=2) * 0.001
    if smooth_coef:
        loss = loss + smooth_coef * feature_smoothing(M, prob["X"])
    return loss


def baseline_attack(prob, cfg, epochs, x0=None):
    """baseline.PGDAttack.attack: nll + 0.001 ||x|| only, Adam, budget projection, final decode (no ensemble).
    Returns loss per iteration, x after every projection, final x, modified_adj and the last victim output."""
    n = prob["n"]
    X, Wt, labels, idx = prob["X"], prob["W"], prob["labels"], prob["idx_attack"]
    P = n * (n - 1) // 2
    x = torch.zeros(P, dtype=X.dtype) if x0 is None else x0.clone().to(X.dtype)
    m, v = torch.zeros_like(x), torch.zeros_like(x)
    losses, xs = [], []
    output = A_hat = None
    for t in range(epochs):
        xr = x.clone().requires_grad_(True)
        A_hat = normalize(expand(xr, n))                               # :53-54 (no clamp in the forward)
        output = victim(X, A_hat, Wt)                                  # :55
        loss = F.nll_loss(output[idx], labels[idx]) + torch.norm(xr, p=2) * 0.001      # :57-58
        loss.backward()
        losses.append(float(loss.detach().double()))
        adam_update(x, xr.grad, m, v, t + 1, cfg["lr"])                # :75
        x = torch.clamp(projection(x, cfg["num_edges"]), 0, 1)         # :77-79
        xs.append(x.clone())
    em = embed(X, A_hat.detach(), Wt, 2)                               # :81
    x_final = decode_tril(em)                                          # :82
    return {"loss": losses, "x_iters": xs, "x_final": x_final, "modified_adj": expand(x_final, n),
            "output": output.detach()}


def mcgpb_attack(prob, cfg, epochs, smooth_from=50, lr=0.1):
    """MC-GPB/topology_attack.py PGDAttack.attack (:36-87): plain gradient descent (lr 0.1), from iteration 50 on the
    loss adds 1e-4 * feature_smoothing (:57-61); final decode relu(Z Z^T) without normalisation (:247-254)."""
    n = prob["n"]
    X, Wt, labels, idx = prob["X"], prob["W"], prob["labels"], prob["idx_attack"]
    P = n * (n - 1) // 2
    x = torch.zeros(P, dtype=X.dtype)
    losses, xs = [], []
    output = A_hat = None
    for t in range(epochs):
        xr = x.clone().requires_grad_(True)
        M = expand(xr, n)
        A_hat = normalize(M)
        output = victim(X, A_hat, Wt)
        loss = F.nll_loss(output[idx], labels[idx]) + torch.norm(xr, p=2) * 0.001
        if t >= smooth_from:
            loss = loss + 1e-4 * feature_smoothing(M, X)
        losses.append(float(loss.detach().double()))
        g = torch.autograd.grad(loss, xr)[0]
        x = x - lr * g                                                 # :66-70
        x = torch.clamp(projection(x, cfg["num_edges"]), 0, 1)         # :75-77
        xs.append(x.clone())
    em = embed(X, A_hat.detach(), Wt, 2)
    G = torch.relu(em @ em.t())
    ii = tril_pairs(n)
    x_final = G[ii[0], ii[1]]
    return {"loss": losses, "x_iters": xs, "x_final": x_final, "modified_adj": expand(x_final, n), "output": output.detach()}


# ----------------------------------------------------------------------------------------------
# AUC / AP with sklearn semantics (main.py:66-75; gcn_parameterized.py:55-65), numpy float64
# ----------------------------------------------------------------------------------------------
def roc_auc(labels, scores):
    """sklearn roc_curve + auc: stable descending sort, thresholds at distinct scores, trapezoid.
    Equivalent closed form: (sum over negatives of #pos above + 0.5 #pos tied) / (P*N)."""
    labels = np.asarray(labels).astype(np.float64).reshape(-1)
    scores = np.asarray(scores).reshape(-1)
    order = np.argsort(-scores, kind="stable")
    s, y = scores[order], labels[order]
    distinct = np.where(np.diff(s))[0]
    thr = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[thr]
    fps = 1 + thr - tps
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    tpr = tps / tps[-1]
    fpr = fps / fps[-1]
    return float(np.trapezoid(tpr, fpr))


def average_precision(labels, scores):
    """sklearn average_precision_score: sum_k (R_k - R_{k-1}) P_k over distinct thresholds."""
    labels = np.asarray(labels).astype(np.float64).reshape(-1)
    scores = np.asarray(scores).reshape(-1)
    order = np.argsort(-scores, kind="stable")
    s, y = scores[order], labels[order]
    distinct = np.where(np.diff(s))[0]
    thr = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[thr]
    fps = 1 + thr - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0, recall]) * precision))


def problem_from_npz(d, dtype=torch.float32):
    """Build (prob, cfg) from a tests/golden/attack_*.npz fixture."""
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    n = int(d["labels"].shape[0])
    prob = dict(n=n, X=t(d["X"]), adj=t(d["adj"].astype(np.float32)), labels=torch.from_numpy(d["labels"]).long(),
                idx_attack=torch.from_numpy(d["idx_attack"]).long(), feature_adj=t(d["feature_adj"]),
                W={k: t(d[k]) for k in ("W1", "b1", "W2", "b2", "Wl", "bl")}, H_A=t(d["H_A2"]), Y_A=t(d["Y_A"]))
    cfg = dict(measure=str(d["measure"]), weights=[float(v) for v in d["weights"]], lr=10 ** float(d["lr_exp"]),
               eps=float(d["eps"]), weight_sup=float(d["weight_sup"]), dataset=str(d["dataset"]),
               use=tuple(bool(u) for u in d["use"]), num_edges=int(d["num_edges"]))
    return prob, cfg
