"""Per-kernel count of the Blackwell-specific SASS mnemonics in the in-tree library (evidence that the tensor-core kernels are
hand-written tcgen05 / TMEM / bulk-copy code):  python tools/sass_summary.py > profiles/sass_r02.txt
UTCHMMA = tcgen05.mma kind::f16, UTCQMMA/UTCIMMA other kinds, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (1-D TMA-engine copy), UTMALDG / UTMASTG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier ops,
MUFU.EX2 / MUFU.SQRT = the XU-pipe ops that bound the attention kernel, HMMA = legacy mma.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dgdm_histopath_lab_b200", "libdgdm_b200.so")
MNEM = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU.EX2", "MUFU.SQRT",
        "HMMA", "LDG", "STG", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, regs = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur).replace("dgdm::", "").replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEM:
                if op == k or op.startswith(k + ".") or (k.startswith("MUFU") and op.startswith(k)):
                    counts[cur][k] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — static instruction counts per kernel (sm_100a)")
    print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{k:>8s}" for k in MNEM))
    for name, c in counts.items():
        print(f"{name[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in MNEM))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"{'TOTAL':58s} {tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in MNEM))


if __name__ == "__main__":
    main()
