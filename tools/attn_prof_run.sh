#!/bin/bash
# Dev tool (GPU box): runs tools/attn_prof.py with every build/ab/lib_prof*.so (built with -DT2_PROF=1 by tools/attn_ab.sh)
# usage: tools/attn_prof_run.sh [nodes] [model]
cd "$(dirname "$0")/../dgdm_histopath_lab_b200" || exit 1
cp libdgdm_b200.so /tmp/bak.so
for L in build/ab/lib_prof*.so; do cp $L libdgdm_b200.so; echo "== $L"; (cd .. && timeout 200 python tools/attn_prof.py ${1:-100000} $2 2>&1 | tail -17); done
cp /tmp/bak.so libdgdm_b200.so
