#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:stream2d_kernel -s 2 \
    -o gpurun_out/r2f_stream2d_16384 python bench.py --size 16384 --steps 12 --warmup 4 --only main --no-cpu-baseline --repeats 1 \
    > gpurun_out/r2f_ncu16k.log 2>&1
tail -3 gpurun_out/r2f_ncu16k.log
python tools/ncu_summary.py gpurun_out/r2f_stream2d_16384.ncu-rep -o gpurun_out/r2f_stream2d_16384.md > /dev/null
rm -f gpurun_out/r2f_stream2d_16384.ncu-rep
head -30 gpurun_out/r2f_stream2d_16384.md
