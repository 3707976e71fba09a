"""Per-source-line warp-stall summary from an ncu report captured with --import-source on (and -lineinfo builds):
    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv --kernel-name K > k.csv ; python tools/ncu_lines.py k.csv [top]
Each SASS row carries the CUDA line it maps to; samples are summed per (file, line)."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur, hdr = None, None
    agg = defaultdict(lambda: defaultdict(int))
    src = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or not r[0].strip().isdigit():
            continue
        key = (cur, int(r[0]))
        src.setdefault(key, r[1].strip())
        if len(r) > ix["# Samples"] and r[ix["# Samples"]].strip().isdigit():
            agg[key]["n"] += int(r[ix["# Samples"]])
            agg[key]["exec"] += int(r[ix["Instructions Executed"]] or 0)
            for s in stalls:
                v = r[ix[s]]
                agg[key][s] += int(v) if v.strip().isdigit() else 0
    tot = sum(a["n"] for a in agg.values())
    print("total samples", tot)
    allst = defaultdict(int)
    for a in agg.values():
        for k, v in a.items():
            if k.startswith("stall_"):
                allst[k] += v
    print({k: round(100 * v / tot, 1) for k, v in sorted(allst.items(), key=lambda x: -x[1])[:10]})
    for key, a in sorted(agg.items(), key=lambda x: -x[1]["n"])[:top]:
        st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_")), key=lambda x: -x[1])[:4]
        st = " ".join(f"{k}={100 * v / max(a['n'], 1):.0f}%" for k, v in st)
        print(f"{key[0]:18s} L{key[1]:<4d} {100 * a['n'] / tot:5.1f}% exec={a['exec']:>12d}  {src[key][:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
