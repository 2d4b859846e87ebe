#!/bin/bash
# last multi-GPU regression on the final commit (2 GPUs): the whole multi-GPU test file and the bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2fm_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2fm_bench.err | tail -1 > gpurun_out/r2fm_bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r2fm_bench_n2.json')); print('n2 value', round(d['value'],1), 'parity', d.get('parity_ok'), 'e2e', round(d['e2e']['value'],1), 'weak5', round(d['config5']['weak']['value'],1))"
