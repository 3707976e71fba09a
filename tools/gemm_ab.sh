#!/bin/bash
# Dev tool (GPU box): times tools/gemm_bench.py with every dgdm_histopath_lab_b200/build/ab/libg_*.so (A/B builds of gemm_tc.cu)
cd "$(dirname "$0")/../dgdm_histopath_lab_b200" || exit 1
cp libdgdm_b200.so /tmp/bak_g.so
(cd .. && timeout 120 python tools/gemm_bench.py 100000 5 > /dev/null 2>&1)
for R in 1 2; do for L in build/ab/libg_*.so; do cp $L libdgdm_b200.so; echo "== $L (round $R)"; (cd .. && timeout 120 python tools/gemm_bench.py ${1:-100000} ${2:-8} 2>&1 | tail -${3:-1}); done; done
cp /tmp/bak_g.so libdgdm_b200.so
