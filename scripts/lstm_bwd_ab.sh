#!/bin/bash
# A/B of the BPTT kernel with the W slice in tensor memory (NNR_LSTM_BWD_TM=1, default) against shared memory (=0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm" 2>&1 | tail -8
for TM in 0 1; do
  echo "== NNR_LSTM_BWD_TM=$TM"
  NNR_LSTM_BWD_TM=$TM NNR_LSTM_DEBUG=1 timeout 300 python scripts/lstm_bench.py 2>&1 | grep -E "bwd|nnr lstm"
done
for TM in 0 1; do
  NNR_LSTM_BWD_TM=$TM timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NNR_LSTM_BWD_TM=$TM', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
