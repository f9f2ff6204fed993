#!/bin/bash
# GPU session 1: parity suite, A/B of the split recurrence, default bench, ncu captures
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests.log
echo "== tests done" >&2
for so in "" meteor_demod_b200/ab_pipe0.so; do
  echo "### LRPT_SO=$so" >> gpurun_out/r2_ab_pipe.log
  LRPT_SO=$so timeout 600 python tools/quick_perf.py --ws --spec --cfg=1:1048576,32:262144,148:262144,1024:262144,4736:262144 >> gpurun_out/r2_ab_pipe.log 2>&1
done
echo "== ab done" >&2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "== bench done rc=$?" >&2
# ncu: single-stream ws kernel (source-level) and the lane kernel at the bench workload
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_ws_kernel -c 1 -o gpurun_out/r2_ws_single python tools/prof_one.py 1 262144 32 5 0 ws > gpurun_out/r2_ncu_ws.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_spec_kernel -c 1 -o gpurun_out/r2_spec_4096 python tools/prof_one.py 4096 131072 32 5 0 spec > gpurun_out/r2_ncu_spec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_lane_kernel -c 1 -o gpurun_out/r2_lane_c1 python tools/prof_one.py 75776 32768 32 5 0 lane > gpurun_out/r2_ncu_lane.log 2>&1
echo "== ncu done" >&2
tail -3 gpurun_out/r2_gpu_tests.log; tail -12 gpurun_out/r2_ab_pipe.log; head -c 1500 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
