#!/bin/bash
# parity tests of the model-level schedule, the short bench, and the kernel timeline of one replay
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "${TESTS:-two_lane or cuda_graph or deterministic or train_loss_and_gradients or scorer_graph or baseline_shape_parity_against}" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
python scripts/step_timeline.py gpurun_out/step_timeline.csv 2>&1 | tail -1
