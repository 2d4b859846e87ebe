#!/bin/bash
# round 2, GPU call 10 (1 GPU): host-resolved 1-D entries -- full suite, coupled rate, config 1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c10_pytest_all.log 2>&1
echo "all rc=$?"; tail -4 gpurun_out/r2c10_pytest_all.log
timeout 200 python benchmarks/next_rows.py --rows coupled > gpurun_out/r2c10_coupled.jsonl 2> gpurun_out/r2c10_coupled.err
cat gpurun_out/r2c10_coupled.jsonl; tail -3 gpurun_out/r2c10_coupled.err
timeout 200 python benchmarks/configs.py --configs 1 > gpurun_out/r2c10_config1.jsonl 2>&1; cat gpurun_out/r2c10_config1.jsonl
