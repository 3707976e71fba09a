#!/bin/bash
# Round-2 evidence run on one GPU: bench (+ reference arm), launch list, full-counter captures of the BIG launches
# (attention, spmm<4>/<2>, the N=512 / N=256 encoder GEMMs at 100k rows), per-launch DRAM table of all SpMM + GEMM launches.
# Usage: bash tools/gpu_bench_r02.sh [tag] [what]      what = all | bench | ncu
TAG=${1:-r02a}; WHAT=${2:-all}
mkdir -p gpurun_out
if [ "$WHAT" != "ncu" ]; then
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench_$TAG.json; tail -n 3 gpurun_out/bench_$TAG.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
echo "reference arm exit $?"; tail -c 900 gpurun_out/bench_ref_$TAG.json
fi
if [ "$WHAT" != "bench" ]; then
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches exit $?"
# 3 warm-up forwards x 24 SpMM launches are skipped; the next 24 are one whole forward
timeout -k 10 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"spmm_gcn|gemm_f16x2" -s 225 -c 75 --csv --log-file gpurun_out/spmm_dram_$TAG.csv $CMD > gpurun_out/ncu_dram_$TAG.log 2>&1
echo "ncu dram table exit $?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:spatial_attn_tc -s 3 -c 1 -f -o gpurun_out/attn_tc_$TAG $CMD > gpurun_out/ncu_attn_$TAG.log 2>&1
echo "ncu attn exit $?"
# the big launches: first 4 SpMM launches of a forward = encoder layer 0 (C=512, <4>) x2 and layer 1 (C=256, <2>) x2
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:spmm_gcn -s 72 -c 4 -f -o gpurun_out/spmm_big_$TAG $CMD > gpurun_out/ncu_spmm_$TAG.log 2>&1
echo "ncu spmm exit $?"
# first 12 GEMM launches of a forward: FeatureEncoder (768->512 x2, 512->512) and encoder layers 0/1 (N=512, N=256) at 100k rows
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x2 -s 153 -c 12 -f -o gpurun_out/gemm_big_$TAG $CMD > gpurun_out/ncu_gemm_$TAG.log 2>&1
echo "ncu gemm exit $?"
fi
ls -la gpurun_out | tail -8
