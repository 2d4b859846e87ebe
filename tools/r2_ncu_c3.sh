#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:streamv_kernel -s 3 \
    -o gpurun_out/r2_streamv_c3 python benchmarks/configs.py --configs 3 --steps 20 --warmup 10 \
    > gpurun_out/r2_ncu_c3.log 2>&1
tail -2 gpurun_out/r2_ncu_c3.log
python tools/ncu_summary.py gpurun_out/r2_streamv_c3.ncu-rep -o gpurun_out/prof_r2_streamv_c3.md > /dev/null
ncu -i gpurun_out/r2_streamv_c3.ncu-rep --page source --csv 2>/dev/null | head -3 | cut -c1-400
rm -f gpurun_out/r2_streamv_c3.ncu-rep
head -45 gpurun_out/prof_r2_streamv_c3.md | tail -38
