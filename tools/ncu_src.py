"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS instructions.
usage: python tools/ncu_src.py file.csv [ntop]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[hdr.index("# Samples")].isdigit()]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: 0 for i in stall}
    T = 0
    for r in data:
        T += int(r[si])
        for i in stall:
            tot[hdr[i]] += int(r[i])
    print("total samples", T, "instructions", sum(int(r[ie]) for r in data))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {k:28s} {v:8d} {v / max(T, 1):.3f}")
    top = sorted([(int(r[si]), n) for n, r in enumerate(data)], reverse=True)[:ntop]
    for s, n in sorted(top, key=lambda t: t[1]):
        r = data[n]
        st = sorted(((hdr[i][6:], int(r[i])) for i in stall if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{n:5d} {s:6d} {r[ie]:>10s}  {r[src].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main()
