"""Per-launch DRAM traffic table from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`:
    python tools/ncu_dram_table.py gpurun_out/spmm_dram_r02.csv profiles/spmm_dram_r02.json
bench.py reads the JSON to fill roofline_hbm.traffic (sum over the SpMM launches of one forward)."""
import csv
import json
import re
import sys


def main(src, dst):
    rows = list(csv.reader(l for l in open(src, newline="") if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv, iid, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    by = {}
    for r in rows[1:]:
        d = by.setdefault(int(r[iid]), {"kernel": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("dgdm::", ""), "dram_bytes": 0.0})
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if r[im].startswith("dram__bytes"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            d["dram_bytes"] += v * mult
            d["dram_read" if "read" in r[im] else "dram_write"] = v * mult
        elif r[im] == "gpu__time_duration.sum":
            d["time_us"] = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    out = {"source": src, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one forward "
                              "(100k-node graph) after 3 warm-up forwards", "launches": [dict(id=i, **d) for i, d in sorted(by.items())]}
    agg = {}
    for d in out["launches"]:
        a = agg.setdefault(d["kernel"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += d["dram_bytes"]; a[2] += d.get("time_us", 0.0)
    out["per_kernel"] = {k: {"launches": v[0], "dram_bytes": v[1], "time_us": v[2]} for k, v in agg.items()}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out["per_kernel"].items():
        print(f"{k:40s} launches {v['launches']:3d}  dram {v['dram_bytes'] / 1e9:8.3f} GB  {v['time_us'] / 1e3:8.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
