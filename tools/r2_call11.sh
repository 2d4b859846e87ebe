#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coupled_golden.py tests/test_coupling_api.py tests/test_gpu_parity.py -x -q -m gpu -k "coupl or thermo or 1d or material or boundary" > gpurun_out/r2c11_pytest.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2c11_pytest.log
for env in "X=1" "FDS_NO_OVERLAP=1"; do
  env $env timeout 200 python benchmarks/next_rows.py --rows coupled 2>gpurun_out/r2c11.err | tee -a gpurun_out/r2c11_coupled.jsonl
done
