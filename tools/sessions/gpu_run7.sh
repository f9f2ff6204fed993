#!/bin/bash
# GPU session 7 (2 GPUs): the unified sharded runner (1-GPU tests + C-level 2-GPU), CLI tests, front-end tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_sharded.py tests/test_cli.py tests/test_frontend.py tests/test_relay.py -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r2_gpu_tests7.log
echo "== tests done" >&2
tail -25 gpurun_out/r2_gpu_tests7.log
