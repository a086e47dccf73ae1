"""CPU-side guards on the contract around the hot path: no oracle on the product path, the L2-friendly tile order
is a bijection, the reference arm of bench.py prints the agreed JSON line."""
import ast
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(top):
    for dp, _, fs in os.walk(top):
        for f in fs:
            if f.endswith(".py"):
                yield os.path.join(dp, f)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mc-gra_b200/ may import it (or the reference shim)."""
    bad = []
    for path in _py_files(os.path.join(ROOT, "mc-gra_b200")):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for nm in names:
                if nm.split(".")[0] in ("oracle", "pgd_oracle", "ref_shim"):
                    bad.append((path, nm))
    assert not bad, bad


def _tri(i):
    return i * (i + 1) // 2


def _tile_coords(t):
    i = int((math.sqrt(8 * t + 1) - 1) / 2)
    while _tri(i + 1) <= t:
        i += 1
    while _tri(i) > t:
        i -= 1
    return i, t - _tri(i)


def _blocked(b, tr0, tr1, R=8):
    """Python restatement of tile_coords_blocked (csrc/common.cuh)."""
    Ib, _ = _tile_coords(_tri(tr0) + b)
    I0 = tr0 + ((Ib - tr0) // R) * R
    I1 = min(I0 + R, tr1)
    rows = I1 - I0
    rb = b - (_tri(I0) - _tri(tr0))
    full = (I0 + 1) * rows
    if rb < full:
        J, I = rb // rows, I0 + rb % rows
    else:
        rb -= full
        J = I0 + 1
        while rb >= I1 - J:
            rb -= I1 - J
            J += 1
        I = J + rb
    return I, J, _tri(I) + J - _tri(tr0)


def test_blocked_tile_order_is_a_bijection_with_column_reuse():
    from mcgra_b200.engine import shard_tile_rows
    shards = [(0, 1), (0, 9), (0, 37), (5, 23), (12, 13)] + list(shard_tile_rows(155, 8)) + list(shard_tile_rows(512, 2))
    for tr0, tr1 in shards:
        nt = _tri(tr1) - _tri(tr0)
        seen, prev = set(), None
        reuse = 0
        for b in range(nt):
            I, J, tix = _blocked(b, tr0, tr1)
            assert tr0 <= I < tr1 and 0 <= J <= I and 0 <= tix < nt
            seen.add(tix)
            reuse += prev is not None and prev == J
            prev = J
        assert len(seen) == nt
        if tr1 - tr0 >= 10:                            # consecutive CTAs mostly share the column operand (7 of 8)
            assert reuse > 0.75 * nt


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, MCGRA_BENCH_SAMPLE_N="96")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_attack_with_noise_or_other_depth_fails_loudly():
    """--eps != 0 and --nlayers != 2 are not built natively: the public call must raise, never fall back to a
    torch / CPU path.  Both guards sit in front of any device work, so this runs without a GPU."""
    import numpy as np
    import pytest
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from mcgra_b200.models.gcn import GCN, embedding_GCN
    from mcgra_b200.topology_attack import PGDAttack
    d = np.load(os.path.join(ROOT, "tests", "golden", "attack_mse_eps_n37.npz"))
    n = int(d["labels"].shape[0])
    victim, emb = helpers.make_models(d, torch.device("cpu"))
    args = helpers.make_args(d)
    assert args.eps != 0
    model = PGDAttack(model=victim, embedding=emb, H_A=torch.from_numpy(d["H_A2"]), Y_A=torch.from_numpy(d["Y_A"]),
                      nnodes=n, loss_type="CE", device="cpu")
    call = lambda: model.attack(args, None, 1e-2, 0, 1.0, tuple(d["weights"]), torch.from_numpy(d["feature_adj"]), 0, 0, 0,
                                None, None, None, torch.from_numpy(d["adj"].astype(np.float32)), d["X"],
                                np.zeros((n, n), np.float32), d["labels"], d["idx_attack"], 10 ** 9, 0, epochs=1)
    with pytest.raises(NotImplementedError, match="eps"):
        call()
    args.eps = 0.0
    f, c = d["X"].shape[1], int(d["Wl"].shape[0])
    deep = GCN(nfeat=f, nclass=c, nhid=16, nlayer=3, dropout=0.5, weight_decay=5e-4, device="cpu")
    model3 = PGDAttack(model=deep, embedding=embedding_GCN(nfeat=f, nhid=16, nlayer=3, device="cpu"),
                       H_A=torch.from_numpy(d["H_A2"]), Y_A=torch.from_numpy(d["Y_A"]), nnodes=n, loss_type="CE", device="cpu")
    model = model3
    with pytest.raises(NotImplementedError, match="layer"):
        call()
