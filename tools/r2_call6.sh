#!/bin/bash
# round 2, GPU call 6 (1 GPU): coupled fields on the device (goldens, rate), full GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coupled_golden.py tests/test_coupling_api.py -x -q -m gpu > gpurun_out/r2c6_pytest_coupled.log 2>&1
echo "coupled rc=$?"; tail -15 gpurun_out/r2c6_pytest_coupled.log
timeout 200 python benchmarks/next_rows.py --rows coupled > gpurun_out/r2c6_coupled.jsonl 2> gpurun_out/r2c6_coupled.err
cat gpurun_out/r2c6_coupled.jsonl; tail -3 gpurun_out/r2c6_coupled.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c6_pytest_all.log 2>&1
echo "all rc=$?"; tail -6 gpurun_out/r2c6_pytest_all.log
