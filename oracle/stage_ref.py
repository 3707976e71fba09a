"""TEST INFRASTRUCTURE — stages the reference's Python package, unmodified, under oracle/_ref/ so that the GPU box (which has no
/root/reference) can run the REAL reference CPU forward as the baseline arm: `bench.py --impl reference` and `cpu_baseline`
then time the reference's own files (oracle/ref_loader.py applies the enumerated patch set P1-P7 as monkey-patches at import
time; the unpatched forward raises for every input).

oracle/_ref/ is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so it travels with the repo
snapshot like the built .so files. Run:  python -m oracle.stage_ref    (also called by __graft_entry__.build() when
/root/reference is present).
"""
import os
import shutil
import sys

SRC = os.environ.get("DGDM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def stage(force=False):
    src_pkg = os.path.join(SRC, "dgdm_histopath")
    dst_pkg = os.path.join(DST, "dgdm_histopath")
    if not os.path.isdir(src_pkg):
        return None
    if os.path.isdir(dst_pkg) and not force:
        return DST
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    n = 0
    for root, dirs, files in os.walk(src_pkg, followlinks=False):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, src_pkg)
        out = os.path.join(dst_pkg, rel) if rel != "." else dst_pkg
        os.makedirs(out, exist_ok=True)
        for f in files:
            if not f.endswith((".py", ".yaml", ".yml", ".json")):
                continue
            s = os.path.join(root, f)
            # utils/logging.py is a symlink to custom_logging.py in the reference: materialise it as a real file
            shutil.copyfile(os.path.realpath(s), os.path.join(out, f))
            n += 1
    with open(os.path.join(DST, "STAGED_FROM.txt"), "w") as fh:
        fh.write(f"unmodified copy of {src_pkg} ({n} files); patches P1-P7 are applied at import time by oracle/ref_loader.py\n")
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
