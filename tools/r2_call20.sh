#!/bin/bash
# round 2, call 20 (the last 45 GPU-seconds): config 3 with the task table cut for 12 warps per SM
# again, and the division / axisymmetric parity tests on that build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 25 python benchmarks/configs.py --configs 3 --steps 200 --warmup 20 2>gpurun_out/r2h_cfg.err | tee gpurun_out/r2h_config3.jsonl | cut -c1-330
timeout 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "division or axisymmetric" 2>&1 | tail -2 | tee gpurun_out/r2h_pytest.log
