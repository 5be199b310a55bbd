cd /root/repo; mkdir -p gpurun_out
b() { env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-profile 2>gpurun_out/_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/_err.txt; }
b NNR_TC_PAIR_MIN_M=1024
b NNR_TC_PAIR_MIN_M=37888
python scripts/step_timeline.py gpurun_out/step_timeline_new.csv 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --gemm-detail 2>gpurun_out/gemm_detail_new.txt | tail -1 > gpurun_out/bench_new.json
