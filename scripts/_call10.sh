cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
b() { env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>gpurun_out/_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/_err.txt; }
b NNR_X=1
b NNR_TC_STAGE_PENALTY=1.15
b NNR_TC_STAGE_PENALTY=1.0
