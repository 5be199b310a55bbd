#!/bin/bash
# whole GPU suite, then bench.py with the CTA-pair row threshold (NNR_TC_PAIR_MIN_M) at each value given, per-shape GEMM table kept
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for V in "$@"; do
  env NNR_TC_PAIR_MIN_M=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --gemm-detail 2>gpurun_out/pair_detail_$V.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PAIR_MIN_M=$V', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
