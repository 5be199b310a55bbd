#!/bin/bash
# A/B of the lane schedule (engine.Lanes): parity tests, then bench.py with NNR_LANES=0 / 1 on the same box, then the
# kernel timeline of one replay.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/lanes_tests.log
cat gpurun_out/lanes_tests.log
for L in ${LANES_SEQ:-0 1 1}; do
  NNR_LANES=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile > gpurun_out/lanes_bench_$L.json 2> gpurun_out/lanes_bench_$L.err
  python - <<P
import json
d=json.loads(open('gpurun_out/lanes_bench_$L.json').read().strip().splitlines()[-1])
print('NNR_LANES=$L', d['value'], d['ms_per_step'], d['e2e']['value'])
P
done
python scripts/step_timeline.py gpurun_out/step_timeline.csv 2>&1 | tail -1
