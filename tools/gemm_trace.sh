#!/bin/bash
# Dev tool: rebuild the library with the GEMM role-wait instrumentation, run tools/gemm_trace.py, restore the normal build.
set -e
cd dgdm_histopath_lab_b200
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
nvcc $F -DDGDM_GEMM_TRACE -c csrc/gemm_tc.cu -o build/gemm_tc_trace.o
cp libdgdm_b200.so /tmp/libdgdm_b200.so.bak
nvcc -shared -o libdgdm_b200.so build/gemm_tc_trace.o build/rowops.o build/graph.o build/attention.o build/attention_tc.o build/attention_simt.o build/builder.o -lcudart
cd ..
python tools/gemm_trace.py || true
cp /tmp/libdgdm_b200.so.bak dgdm_histopath_lab_b200/libdgdm_b200.so
