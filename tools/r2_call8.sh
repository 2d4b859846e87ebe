#!/bin/bash
# round 2, GPU call 8 (1 GPU): pipelined fds_simulate with the plan arena; reference stub test; sanitizers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_stub.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r2c8_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2c8_pytest.log
ab() {
  local label="$1"; shift
  out=$(env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --only main,e2e --no-cpu-baseline 2>gpurun_out/r2c8_ab.err | tail -1)
  echo "{\"label\": \"$label\", \"line\": $out}" >> gpurun_out/r2c8_ab.jsonl
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(round(d["value"],1), "e2e", round(e.get("value",0),2), e.get("seconds"), e.get("phases"), e.get("error"))' 2>/dev/null)"
}
rm -f gpurun_out/r2c8_ab.jsonl
ab nopipe FDS_NO_PIPELINE=1
ab pipe8 X=1
ab pipe4 FDS_PIPELINE_BANDS=4
ab pipe6 FDS_PIPELINE_BANDS=6
ab pipe12 FDS_PIPELINE_BANDS=12
ab pipe16 FDS_PIPELINE_BANDS=16
ab pipe8_again X=1
bash tools/r2_sanitize.sh 1
