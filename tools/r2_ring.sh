#!/bin/bash
cd "$(dirname "$0")/.."
for v in "" ring4 ring8; do
  lib=$PWD/pyfds_b200/libfdsb200${v:+_$v}.so
  FDS_LIBRARY_PATH=$lib timeout 300 python benchmarks/configs.py --configs 2,3,6,4 --steps 200 --warmup 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('${v:-ring6}', d['config'], d['kernel'], round(d['gcell_updates_per_s'],1))"
done
