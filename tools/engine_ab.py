"""Same-process A/B of the propagate engines (clocks drift between boxes / under the power cap, so only numbers taken
in one run are comparable).  usage: python tools/engine_ab.py [workload] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcgra_b200 import _native as N  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "pubmed"]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    prob = bench.build_problem(wl, dev, host_feature_adj=False)
    args = bench.make_args()
    atk, adj = bench.make_attack(prob, dev)
    atk.attack(args, None, 1e-2, 0, 1.0, bench.PROFILE_A, prob["feature_adj"], 0, 0, 0, None, None, None, adj, prob["X"],
               torch.zeros(1), prob["labels"], prob["idx_attack"], 10 ** 15, 0, epochs=0, _engine_epochs=400,
               _skip_finalize=True)
    eng = atk.engine
    for _ in range(3):
        eng.iterate()
    out = {}
    for rounds in range(1):
        for name, sel, split in (
                                 ("v4 tcgen05 both (T^T in TMEM)", 4, True),
                                 ("v5 tcgen05 fp16x2 single image", 5, True),
                                 ("v5 dbg: no flush", 5, 101), ("v5 dbg: no MMA", 5, 102), ("v5 dbg: neither", 5, 103)):
            N.lib().mcgra_set_engine(0, sel)
            N.lib().mcgra_set_engine(0, split if split is not True else 100)
            split = True
            eng.split_elem = split
            eng.iterate()
            torch.cuda.synchronize()
            N.TIMERS["on"] = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                eng.iterate()
            e1.record()
            torch.cuda.synchronize()
            kt = {k: round(float(np.mean([a.elapsed_time(b) for a, b in v])), 3) for k, v in N.TIMERS["on"].items()}
            N.TIMERS["on"] = None
            out[name] = {"ms_per_iter": round(e0.elapsed_time(e1) / steps, 3), **{k: v for k, v in kt.items() if v > 0.05}}
            print(name, json.dumps(out[name]), flush=True)
    N.lib().mcgra_set_engine(0, 100)
    N.lib().mcgra_set_engine(0, 5)


if __name__ == "__main__":
    main()
