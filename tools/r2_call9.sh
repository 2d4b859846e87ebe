#!/bin/bash
# round 2, GPU call 9 (2 GPUs): coupled pair kernel, sanitizer on the 2-slab peer path, multi-GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_coupled_golden.py tests/test_coupling_api.py -x -q -m gpu > gpurun_out/r2c9_pytest_coupled.log 2>&1
echo "coupled rc=$?"; tail -4 gpurun_out/r2c9_pytest_coupled.log
timeout 200 python benchmarks/next_rows.py --rows coupled > gpurun_out/r2c9_coupled.jsonl 2> gpurun_out/r2c9_coupled.err
FDS_NO_PAIR_KERNEL=1 timeout 200 python benchmarks/next_rows.py --rows coupled >> gpurun_out/r2c9_coupled.jsonl 2>> gpurun_out/r2c9_coupled.err
cat gpurun_out/r2c9_coupled.jsonl; tail -3 gpurun_out/r2c9_coupled.err
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2_sanitize_memcheck_peer2_%p.log \
    python tools/sanitize_case.py peer2 > gpurun_out/r2_sanitize_peer2.out 2>&1
echo "peer2 memcheck rc=$?"; tail -3 gpurun_out/r2_sanitize_peer2.out; tail -qn 2 gpurun_out/r2_sanitize_memcheck_peer2_*.log
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2c9_pytest_multi.log 2>&1
echo "multi rc=$?"; tail -4 gpurun_out/r2c9_pytest_multi.log
