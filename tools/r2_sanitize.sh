#!/bin/bash
# compute-sanitizer logs for profiles/ (SURVEY.md 5.2). $1 = number of GPUs on the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitize_memcheck.log \
    python tools/sanitize_case.py stream2d streamv axi thermal line1d > gpurun_out/r2_sanitize_memcheck.out 2>&1
echo "memcheck rc=$?"; cat gpurun_out/r2_sanitize_memcheck.out | tail -6; tail -3 gpurun_out/r2_sanitize_memcheck.log
timeout 1200 $S --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2_sanitize_racecheck.log \
    python tools/sanitize_case.py stream2d streamv > gpurun_out/r2_sanitize_racecheck.out 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r2_sanitize_racecheck.out; tail -3 gpurun_out/r2_sanitize_racecheck.log
timeout 600 $S --tool synccheck --error-exitcode 9 --log-file gpurun_out/r2_sanitize_synccheck.log \
    python tools/sanitize_case.py stream2d streamv > gpurun_out/r2_sanitize_synccheck.out 2>&1
echo "synccheck rc=$?"; tail -2 gpurun_out/r2_sanitize_synccheck.log
if [ "${1:-1}" -ge 2 ]; then
  timeout 900 $S --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2_sanitize_memcheck_peer2_%p.log \
      python tools/sanitize_case.py peer2 > gpurun_out/r2_sanitize_peer2.out 2>&1
  echo "peer2 memcheck rc=$?"; tail -3 gpurun_out/r2_sanitize_peer2.out; tail -n 2 gpurun_out/r2_sanitize_memcheck_peer2_*.log
fi
