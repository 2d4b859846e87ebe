#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_coupled_golden.py tests/test_coupling_api.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c12_pytest.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/r2c12_pytest.log
