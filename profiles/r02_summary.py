"""Round-2 summaries from gpurun_out/ (tools/profile_r2.sh): per-kernel shares of the launch lists, selected ncu metrics of
the fully captured kernels, and the dram traffic per launch that bench.py reports as roofline.traffic."""
import csv, io, json, os, re, subprocess, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
sys.path.insert(0, os.path.join(ROOT, "profiles"))
from ncu_summary import METRICS, short


def launches(csv_name, out, what):
    rows = [r for r in csv.reader(open(os.path.join(G, csv_name), errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        c = agg.setdefault(short(r[ki]), [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", out), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {what}\n# cold-cache, serialised: compare SHARES\n")
        f.write(f"{'kernel':62s}{'count':>6s}{'total_us':>13s}{'share':>8s}\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{k:62s}{c:6d}{t:13.1f}{100 * t / tot:7.1f}%\n")


def full(reps, out):
    traffic = {}
    with open(os.path.join(ROOT, "profiles", out), "w", newline="") as f:
        w = None
        for rep in reps:
            txt = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(txt)))
            hdr, units = rows[0], rows[1]
            cols = [hdr.index("Kernel Name")] + [next((i for i, h in enumerate(hdr) if h.endswith(m)), None) for m in METRICS]
            if w is None:
                w = csv.writer(f)
                w.writerow(["Kernel Name"] + METRICS)
                w.writerow([""] + [units[c] if c is not None else "" for c in cols[1:]])
            for r in rows[2:]:
                w.writerow([r[cols[0]]] + [(r[c] + " " + units[c]) if c is not None else "" for c in cols[1:]])
                sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                traffic[short(r[cols[0]])] = float(r[cols[2]]) * sc[units[cols[2]]] + float(r[cols[3]]) * sc[units[cols[3]]]
    return traffic


if __name__ == "__main__":
    launches("r02_large_A_launches.csv", "r02_large_A_launches.txt", "bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity (n = 65536, Profile A)")
    launches("r02_pubmed_B_launches.csv", "r02_pubmed_B_launches.txt", "bench.py --workload pubmed --profile B --steps 2 --warmup 1 ... (n = 19717, HSIC)")
    # r02_fold_rs / r02_elem_rs: the final streaming kernels (tools/profile_r2b.sh); r02b_fold_tc: the one-tile-per-CTA
    # engine they replaced, steady-state launch; pairs / gemm captures from tools/profile_r2.sh
    t = full(["r02_fold_rs.ncu-rep", "r02_elem_rs.ncu-rep", "r02_prop_h.ncu-rep", "r02b_fold_tc.ncu-rep", "r02_pairs_tc.ncu-rep", "r02_gemm3.ncu-rep"],
             "r02_ncu_full.csv")
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    allt = json.load(open(tp)) if os.path.exists(tp) else {}
    pick = lambda pre: next((v for k, v in t.items() if k.startswith(pre)), 0)
    allt["large_A"] = {"mcgra_fold_adam": round(pick("k_fold_rs")), "elem_stats": round(pick("k_elem_rs")),
                       "mcgra_pairs": round(pick("k_pairs_tc")),
                       "_note": "k_elem_rs: read 17.2 GB (= algorithmic), writes 17 MB; k_fold_rs: read 40.3 + write 25.8 GB"}
    allt["pubmed_B"] = {k: round(t.get("k_gemm3<2>", 0)) for k in ("gemm_grad", "gemm_c1", "gemm_c2_T")}
    allt["_source_r02"] = "profiles/r02_ncu_full.csv (ncu --set full --clock-control none; dram bytes read + written per launch)"
    json.dump(allt, open(tp, "w"), indent=1, sort_keys=True)
    print(json.dumps(t, indent=1))
