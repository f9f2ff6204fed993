#!/bin/bash
# GPU session 3: two-warp recurrence v2 (loop warp scales and mixes)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ws or spec or grid or ragged or batch_of" 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests3.log
echo "== tests done" >&2
rm -f gpurun_out/r2_ab_split2.log
timeout 600 python tools/quick_perf.py --ws --spec --all --cfg=1:1048576,32:262144,1184:262144,2368:262144,4736:262144 >> gpurun_out/r2_ab_split2.log 2>&1
echo "== perf done" >&2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_ws_kernel -c 1 -o gpurun_out/r2_ws_split2_single python tools/prof_one.py 1 262144 32 5 0 ws > gpurun_out/r2_ncu_ws3.log 2>&1
tail -4 gpurun_out/r2_gpu_tests3.log; cat gpurun_out/r2_ab_split2.log | grep -v "period gen"
