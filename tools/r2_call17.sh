#!/bin/bash
# GPU call 17: one-process slabs after fds_step_prepare; $1 = GPUs on the box, $2 = pytest -k selection
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "$2" > gpurun_out/r2c17_pytest_n$1.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r2c17_pytest_n$1.log
