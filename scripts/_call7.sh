cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>gpurun_out/_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/_err.txt; }
b NNR_TC_CHAIN_K=1024
b NNR_TC_CHAIN_K=2048
NNR_TC_CHAIN_K=2048 timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
