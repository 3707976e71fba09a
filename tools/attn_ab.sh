#!/bin/bash
# Dev tool: A/B attention-kernel builds. Here (no GPU):  tools/attn_ab.sh build NAME "-DFLAG=..."   compiles csrc/attention_tc.cu with
# the flags and links dgdm_histopath_lab_b200/build/ab/lib_NAME.so from the other (already built) objects. On the GPU box:
# tools/attn_ab.sh run [nodes] [reps]   times every lib_*.so with tools/attn_time.py (the .so files travel with the snapshot).
cd "$(dirname "$0")/../dgdm_histopath_lab_b200" || exit 1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
if [ "$1" = "build" ]; then
  mkdir -p build/ab
  nvcc $F $3 -c csrc/attention_tc.cu -o build/ab/attention_tc_$2.o || exit 1
  OBJS=$(ls build/*.o | grep -v attention_tc.o)
  nvcc -shared -o build/ab/lib_$2.so $OBJS build/ab/attention_tc_$2.o -lcudart && echo "built lib_$2.so ($3)"
elif [ "$1" = "run" ]; then
  cp libdgdm_b200.so /tmp/libdgdm_b200.so.bak
  (cd .. && timeout 120 python tools/attn_time.py ${2:-100000} 10 > /dev/null 2>&1)      # throw-away: the first process on a fresh box runs ~1 ms slow
  for ROUND in 1; do
  for L in build/ab/lib_*.so; do
    case "$L" in *prof*) continue;; esac
    cp "$L" libdgdm_b200.so
    echo "== $L (round $ROUND)"; (cd .. && timeout 120 python tools/attn_time.py ${2:-100000} ${3:-5} 2>&1 | tail -1)
  done
  done
  cp /tmp/libdgdm_b200.so.bak libdgdm_b200.so
fi
