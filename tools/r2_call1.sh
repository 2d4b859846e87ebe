#!/bin/bash
# round 2, GPU call 1 (1 GPU): parity suite with overlapped sweeps, A/B of overlap / task height /
# warps per CTA at 4096^2 and 16384^2, then the full default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt
free -g >> gpurun_out/r2c1_smi.txt; nproc >> gpurun_out/r2c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -3 gpurun_out/r2c1_pytest.log
ab() {  # label, env assignments...
  local label="$1"; shift
  for size in 4096; do
    out=$(env "$@" timeout 300 python bench.py --steps 400 --warmup 40 --size $size --only main --no-cpu-baseline 2>gpurun_out/r2c1_ab.err | tail -1)
    echo "{\"label\": \"$label\", \"size\": $size, \"line\": $out}" >> gpurun_out/r2c1_ab.jsonl
    echo "$label $size $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), d["repeats"]["value_min"], d["repeats"]["value_max"])' 2>/dev/null)"
  done
}
rm -f gpurun_out/r2c1_ab.jsonl
ab base_nooverlap FDS_NO_OVERLAP=1
ab overlap X=1
ab overlap_rows128 FDS_TARGET_ROWS=128
ab overlap_rows96 FDS_TARGET_ROWS=96
ab overlap_rows64 FDS_TARGET_ROWS=64
ab overlap_rows48 FDS_TARGET_ROWS=48
ab nooverlap_rows96 FDS_NO_OVERLAP=1 FDS_TARGET_ROWS=96
W1=$PWD/pyfds_b200/libfdsb200_w1.so
ab w1_nooverlap FDS_LIBRARY_PATH=$W1 FDS_NO_OVERLAP=1
ab w1_overlap FDS_LIBRARY_PATH=$W1
ab w1_overlap_rows128 FDS_LIBRARY_PATH=$W1 FDS_TARGET_ROWS=128
ab w1_overlap_rows96 FDS_LIBRARY_PATH=$W1 FDS_TARGET_ROWS=96
ab w1_overlap_rows64 FDS_LIBRARY_PATH=$W1 FDS_TARGET_ROWS=64
ab w1_overlap_rows48 FDS_LIBRARY_PATH=$W1 FDS_TARGET_ROWS=48
# the full default line as the driver runs it (wall time matters)
/usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench_default.json 2> gpurun_out/r2c1_bench_default.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2c1_bench_default.json
grep -E "Elapsed|Maximum resident" gpurun_out/r2c1_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c1_bench_reference.json 2> gpurun_out/r2c1_bench_reference.err
echo "ref rc=$?"
