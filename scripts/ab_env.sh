#!/bin/bash
# usage: ab_env.sh VAR v1 v2 ...   -- bench.py (short) with VAR set to each value on the same box, then the timeline with the last
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VAR=$1; shift
for V in "$@"; do
  env $VAR=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$V', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
env $VAR=$V python scripts/step_timeline.py gpurun_out/step_timeline.csv 2>&1 | tail -1
