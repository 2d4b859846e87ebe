#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2c14_pytest.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/r2c14_pytest.log
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitize_memcheck_pipeline_coupled.log \
    python tools/sanitize_case.py pipeline coupled > gpurun_out/r2_sanitize_memcheck_pipeline_coupled.out 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitize_memcheck_pipeline_coupled.out; tail -2 gpurun_out/r2_sanitize_memcheck_pipeline_coupled.log
bash tools/r2_ncu16k.sh
