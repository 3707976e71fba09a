#!/bin/bash
# Dev tool: time the tcgen05 attention kernel for several TC_POLY_EVERY settings (fraction of exp2 moved off the MUFU pipe).
cd dgdm_histopath_lab_b200
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
cp libdgdm_b200.so /tmp/libdgdm_b200.so.bak
for V in "$@"; do
  nvcc $F -DTC_POLY_EVERY=$V -c csrc/attention_tc.cu -o build/attention_tc_exp.o 2>/dev/null
  nvcc -shared -o libdgdm_b200.so build/gemm_tc.o build/rowops.o build/graph.o build/attention.o build/attention_tc_exp.o build/attention_simt.o build/builder.o -lcudart
  echo "TC_POLY_EVERY=$V"; (cd .. && timeout 120 python tools/attn_only.py 100000 1 4 2>&1 | tail -2)
done
cp /tmp/libdgdm_b200.so.bak libdgdm_b200.so
