#!/bin/bash
# rebuild seq_embed.cu with different EB_CHUNK values on the GPU box and print the embedding-scatter time of the bench step
for c in 64 128 256; do
  touch nnr_b200/csrc/seq_embed.cu
  NVCC_EXTRA=-DEB_CHUNK=$c bash nnr_b200/csrc/build.sh > /dev/null 2>&1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('EB_CHUNK=$c', 'imp/s %.0f' % d['value'], 'embed_gather_bwd %.3f ms' % d['breakdown_ms_per_step']['embed_gather_bwd']['ms'])"
done
