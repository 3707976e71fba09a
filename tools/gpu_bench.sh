#!/bin/bash
# bench + ncu evidence for one GPU. Usage: bash tools/gpu_bench.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 3500 gpurun_out/bench_$TAG.json; tail -n 5 gpurun_out/bench_$TAG.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
echo "reference arm exit $?"; tail -c 1200 gpurun_out/bench_ref_$TAG.json
timeout -k 10 600 python bench.py --workload batch4x50k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_batch_$TAG.json 2>> gpurun_out/bench_$TAG.err
echo "batch4x50k exit $?"; head -c 400 gpurun_out/bench_batch_$TAG.json
kill $SMI
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout -k 10 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout -k 10 600 $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:spatial_attn_tc -s 3 -c 1 -f -o gpurun_out/attn_tc_$TAG $CMD > gpurun_out/ncu_attn_$TAG.log 2>&1
echo "ncu attn exit $?"
timeout -k 10 600 $CMD > gpurun_out/plain3_$TAG.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_f16x2|spmm_gcn" -s 60 -c 12 -f -o gpurun_out/gemm_spmm_$TAG $CMD > gpurun_out/ncu_gemm_$TAG.log 2>&1
echo "ncu gemm/spmm exit $?"
ls -la gpurun_out | tail -12
