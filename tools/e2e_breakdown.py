"""Wall-clock phases of the end-to-end public-API call bench.py times (PGDAttack.attack from pinned host tensors + AUC).
usage: python tools/e2e_breakdown.py [workload] [epochs]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcgra_b200 import metrics  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "large"]
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda:0")
    prob = bench.build_problem(wl, dev, host_feature_adj=True)
    args = bench.make_args()
    atk, adj = bench.make_attack(prob, dev)
    for rep in range(3):
        atk.adj_changes.data.zero_()
        torch.cuda.synchronize()
        if rep == 2:
            from mcgra_b200 import _native as N
            N.TIMERS["on"] = {}
        t0 = time.perf_counter()
        atk.attack(args, None, 10 ** args.lr, 0, 1.0, bench.PROFILE_A, prob["feature_adj"], 0, 0, 0, None, None, None, adj,
                   prob["X"], torch.zeros(1), prob["labels"], prob["idx_attack"], 10 ** 15, 0, epochs=K, _timing=True)
        t1 = time.perf_counter()
        auc, ap = metrics.auc_ap_from_edges(atk.modified_adj, prob["edges"])
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out = {k: round(v, 4) for k, v in atk._timing.items()}
        out.update({"attack_total": round(t1 - t0, 4), "auc_ap": round(t2 - t1, 4), "epochs": K})
        print(json.dumps(out), flush=True)
        if rep == 2:
            import numpy as np
            kt = {k: round(float(np.sum([a.elapsed_time(b) for a, b in v])), 2) for k, v in N.TIMERS["on"].items()}
            N.TIMERS["on"] = None
            print("per C entry, ms summed over the call:", json.dumps(dict(sorted(kt.items(), key=lambda kv: -kv[1]))), flush=True)


if __name__ == "__main__":
    main()
