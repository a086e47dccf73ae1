"""Turn the artefacts of profiles/ncu_capture.sh (gpurun_out/launches_<W>_<TAG>.csv and prof_<W>_<TAG>.ncu-rep) into the
committed summaries:  profiles/<out>_launches.txt (per-kernel share of the step), profiles/<out>_ncu_full.csv (one row
per captured launch, selected metrics) and profiles/ncu_traffic.json (dram bytes per launch, read by bench.py).
usage: python profiles/ncu_summary.py <workload> <tag> <out-prefix>      (needs `ncu` on PATH to read the report)"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
           # (r2) tensor-core operand traffic from shared memory and the HMMA sub-pipe's busy cycles
           "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("at::", "")
    return name.strip()[:60]


def launches(workload, tag, out):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{workload}_{tag}.csv")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
        k = short(r[ki])
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{out}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --workload {workload} --steps 2 "
                f"--warmup 1 --no-e2e --no-cpu (problem setup + 3 iterations)\n# cold-cache, serialised: compare SHARES\n")
        f.write(f"{'kernel':62s}{'count':>6s}{'total_us':>13s}{'share':>8s}\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{k:62s}{c:6d}{t:13.1f}{100 * t / tot:7.1f}%\n")
    return agg


def full(workload, tag, out):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{workload}_{tag}.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [next((i for i, h in enumerate(hdr) if h.endswith(m)), None) for m in METRICS]
    traffic = {}
    with open(os.path.join(ROOT, "profiles", f"{out}_ncu_full.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + METRICS)
        w.writerow([""] + [units[c] if c is not None else "" for c in cols[1:]])
        for r in rows[2:]:
            w.writerow([r[c] if c is not None else "" for c in cols])
            name = short(r[cols[0]])
            rd, wr = float(r[cols[2]]), float(r[cols[3]])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[cols[2]]]
            traffic.setdefault(name, []).append((rd + wr) * scale)
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    allt = json.load(open(tp)) if os.path.exists(tp) else {}
    allt[workload] = {k.replace("<(int)", "<"): round(sum(v) / len(v)) for k, v in traffic.items()}
    allt["_source"] = f"profiles/{out}_ncu_full.csv (ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(allt, open(tp, "w"), indent=1, sort_keys=True)
    return allt[workload]


if __name__ == "__main__":
    wl, tag, out = sys.argv[1], sys.argv[2], sys.argv[3]
    launches(wl, tag, out)
    print(json.dumps(full(wl, tag, out), indent=1))
