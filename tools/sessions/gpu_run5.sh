#!/bin/bash
# GPU session 5 (2 GPUs): bench --gpus 2 incl. the c4 sub-record over two ranks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "== bench2 rc=$?" >&2
tail -5 gpurun_out/r2_bench_n2.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().split('\n')[-1])
for k in ('value','e2e','value_locked','single_stream'):
    print(k, json.dumps(d.get(k))[:400])
print('c4', json.dumps(d['c4'])[:3000])
P
