#!/bin/bash
# GPU session 2: two-warp recurrence -- parity of ws/spec, A/B against the one-warp build, FIR stage tests + bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fir_stage.py -m gpu -x -q -k "ws or spec or grid or fir_stage or ragged or batch_of" 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests2.log
echo "== tests done" >&2
rm -f gpurun_out/r2_ab_split.log
for so in "" meteor_demod_b200/ab_split0.so; do
  echo "### LRPT_SO=$so" >> gpurun_out/r2_ab_split.log
  LRPT_SO=$so timeout 600 python tools/quick_perf.py --ws --spec --all --cfg=1:1048576,32:262144,148:262144,1184:262144,2368:262144,4736:262144 >> gpurun_out/r2_ab_split.log 2>&1
done
echo "== ab done" >&2
timeout 600 python tools/fir_stage_bench.py > gpurun_out/r2_fir_stage.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_ws_kernel -c 1 -o gpurun_out/r2_ws_split_single python tools/prof_one.py 1 262144 32 5 0 ws > gpurun_out/r2_ncu_ws2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_stage_kernel -c 1 -o gpurun_out/r2_fir_stage python tools/fir_stage_bench.py > gpurun_out/r2_ncu_fir.log 2>&1
tail -4 gpurun_out/r2_gpu_tests2.log; cat gpurun_out/r2_ab_split.log | grep -v "period gen"; cat gpurun_out/r2_fir_stage.log | cut -c1-400
