#!/bin/bash
# Round-end evidence run on ONE B200 (under gpurun): next-row benchmarks, launch list of the bench,
# ncu --set full captures of the shipped kernels, summarised on the box (the reports are too large to
# travel: only the summaries and two reports come back in gpurun_out/).
S="python tools/ncu_summary.py"
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 200 python benchmarks/next_rows.py --steps 1000 > gpurun_out/f_next_rows.jsonl 2>gpurun_out/f_next.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/f_launches.csv python bench.py --steps 40 --warmup 8 --no-cpu-baseline \
    > gpurun_out/f_launches.log 2>&1
timeout 200 $NCU -k regex:stream2d_kernel -s 6 -o /tmp/f_stream2d_4096 \
    python bench.py --steps 40 --warmup 8 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu1.log 2>&1
$S /tmp/f_stream2d_4096.ncu-rep -o gpurun_out/f_stream2d_4096.md > /dev/null
for c in 6 3 7; do
    timeout 200 $NCU -k regex:stream -s 6 -o /tmp/f_c$c \
        python benchmarks/configs.py --configs $c --steps 20 --warmup 10 > gpurun_out/f_ncu_c$c.log 2>&1
    $S /tmp/f_c$c.ncu-rep -o gpurun_out/f_c$c.md > /dev/null
done
timeout 200 $NCU -k regex:line1d_kernel -s 20 -o /tmp/f_line1d \
    python benchmarks/configs.py --configs 1 > gpurun_out/f_ncu_line1d.log 2>&1
$S /tmp/f_line1d.ncu-rep -o gpurun_out/f_line1d.md > /dev/null
cp /tmp/f_c6.ncu-rep /tmp/f_stream2d_4096.ncu-rep gpurun_out/
ls -la gpurun_out
