#!/bin/bash
# round 2, GPU call 7 (1 GPU): pipelined fds_simulate -- full GPU suite, e2e A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c7_pytest_all.log 2>&1
echo "all rc=$?"; tail -6 gpurun_out/r2c7_pytest_all.log
ab() {
  local label="$1"; shift
  out=$(env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --only main,e2e --no-cpu-baseline 2>gpurun_out/r2c7_ab.err | tail -1)
  echo "{\"label\": \"$label\", \"line\": $out}" >> gpurun_out/r2c7_ab.jsonl
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(round(d["value"],1), "e2e", round(e.get("value",0),2), e.get("seconds"), e.get("phases"), e.get("error"))' 2>/dev/null)"
}
rm -f gpurun_out/r2c7_ab.jsonl
ab nopipe FDS_NO_PIPELINE=1
ab pipe8 X=1
ab pipe4 FDS_PIPELINE_BANDS=4
ab pipe16 FDS_PIPELINE_BANDS=16
ab pipe12 FDS_PIPELINE_BANDS=12
ab pipe8_again X=1
