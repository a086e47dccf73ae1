"""Where the time of mcgra_auc_ap goes at the bench size: the real n x n result of a short attack, ranked with the
histogram updates / the search switched off (developer knobs of mcgra_set_engine(6, 100 + bits)), and mcgra_ensemble timed
on the same state.  usage: python tools/auc_ab.py [workload] [epochs]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcgra_b200 import metrics  # noqa: E402
from mcgra_b200 import _native as N  # noqa: E402


def timed(fn, reps=3):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        out.append(round(a.elapsed_time(b), 2))
    return out, r


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "large"]
    dev = torch.device("cuda:0")
    prob = bench.build_problem(wl, dev, host_feature_adj=True)
    args = bench.make_args()
    atk, adj = bench.make_attack(prob, dev)
    N.TIMERS["on"] = {}
    atk.attack(args, None, 10 ** args.lr, 0, 1.0, bench.PROFILE_A, prob["feature_adj"], 0, 0, 0, None, None, None, adj,
               prob["X"], torch.zeros(1), prob["labels"], prob["idx_attack"], 10 ** 15, 0,
               epochs=int(sys.argv[2]) if len(sys.argv) > 2 else 2)
    torch.cuda.synchronize()
    kt = {k: round(sum(a.elapsed_time(b) for a, b in v), 2) for k, v in N.TIMERS["on"].items()
          if k in ("mcgra_ensemble", "mcgra_decode_to_tiles", "mcgra_tiles_to_tril", "mcgra_dense_to_tiles")}
    N.TIMERS["on"] = None
    print("finalisation entries, ms:", json.dumps(kt), flush=True)
    S = atk.modified_adj
    edges = prob["edges"]
    n = S.shape[0]
    e = torch.as_tensor(edges, device=dev).long()
    lab = torch.zeros(n, n, dtype=torch.uint8, device=dev)
    lab[e[:, 0], e[:, 1]] = 1
    lab[e[:, 1], e[:, 0]] = 1
    npos = int(lab.sum().item())
    ps = S[lab.bool()]
    lo, hi = float(ps.min()), float(ps.max())
    below = int((S < lo).sum().item())
    above = int((S > hi).sum().item())
    print(json.dumps({"npos": npos, "pos_min": lo, "pos_max": hi, "entries_below_all_pos": below, "entries_above_all_pos": above,
                      "distinct_pos_scores": int(torch.unique(ps).numel())}), flush=True)
    # (the logs under profiles/r02_auc_ablation.txt also list "no search" / "private hist w/o warp aggregation" variants:
    #  developer knobs of the intermediate kernels, removed again)
    for name, eng, dbg in (("tab (default)", 1, 0), ("tab, no histogram", 1, 1), ("1024-sample kernel", 0, 0)):
        N.lib().mcgra_set_engine(6, eng)
        N.lib().mcgra_set_engine(6, 100 + dbg)
        t, r = timed(lambda: metrics.roc_auc_ap(S, lab, npos_max=npos))
        print(f"{name:40s} ms {t}  auc/ap {r}", flush=True)
    N.lib().mcgra_set_engine(6, 1)
    N.lib().mcgra_set_engine(6, 100)


if __name__ == "__main__":
    main()
