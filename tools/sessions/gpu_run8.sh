#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_sharded.py -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r2_gpu_tests8.log
tail -25 gpurun_out/r2_gpu_tests8.log
