cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "embed" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>gpurun_out/_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/_err.txt; }
b NNR_TC_CHAIN_K=2048
b NNR_TC_CHAIN_K=4096
python scripts/step_timeline.py gpurun_out/step_timeline_new2.csv 2>&1 | tail -1
