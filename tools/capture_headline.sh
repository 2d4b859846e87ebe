#!/bin/bash
# Headline evidence on ONE B200 (under gpurun): GPU tests, the bench lines, the launch list of the bench
# and one ncu --set full capture of the dominant kernel (summarised on the box).
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/h_pytest.log
timeout 300 python bench.py > gpurun_out/h_bench_default.json 2>gpurun_out/h_def.err
timeout 300 python bench.py --size 16384 --steps 400 --warmup 40 --no-e2e --no-cpu-baseline \
    > gpurun_out/h_bench_16384.json 2>gpurun_out/h_16k.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/h_launches.csv python bench.py --steps 40 --warmup 8 --no-cpu-baseline \
    > gpurun_out/h_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:stream2d_kernel -s 6 \
    -o /tmp/h_stream2d_4096 python bench.py --steps 40 --warmup 8 --no-e2e --no-cpu-baseline \
    > gpurun_out/h_ncu.log 2>&1
python tools/ncu_summary.py /tmp/h_stream2d_4096.ncu-rep -o gpurun_out/h_stream2d_4096.md > /dev/null
timeout 100 python benchmarks/configs.py --configs 2,4,5,7,8 > gpurun_out/h_configs.jsonl 2>/dev/null
cat gpurun_out/h_pytest.log; cut -c1-120 gpurun_out/h_bench_default.json; cut -c1-120 gpurun_out/h_bench_16384.json
