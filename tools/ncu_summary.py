"""Summarise ncu outputs brought back from the GPU box into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r01b.csv  > profiles/launches_r01b.txt
  python tools/ncu_summary.py report   gpurun_out/attn_r01b.ncu-rep  > profiles/attn_r01b.txt
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def launches(path):
    rows = list(csv.reader(l for l in open(path, newline="") if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    agg, order = OrderedDict(), []
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("dgdm::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        order.append((int(r[iid]), name, us))
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES) — {path}")
    print(f"# {len(order)} launches captured, total {tot/1e3:.3f} ms")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:8d} {us:12.1f} {100*us/tot:6.2f}% {us/n:10.1f}")


def report(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on — {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, zip(units, r)))
        print(f"\n## {d['Kernel Name'][1]}   grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}")
        for k in list(KEYS) + list(extra):
            if k in d:
                print(f"{k:90s} {d[k][1]:>16s} {d[k][0]}")
        stalls = []
        for h in hdr:
            m = STALL.match(h)
            if m and "per_issue_active" in h and "not_issued" not in h:
                try:
                    stalls.append((float(d[h][1].replace(",", "")), m.group(1) or m.group(2)))
                except ValueError:
                    pass
        if stalls:
            print("warp stall reasons (avg warps stalled per issue-active cycle), top 8:")
            for v, n in sorted(stalls, reverse=True)[:8]:
                print(f"    {n:40s} {v:8.3f}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3:])
