#!/bin/bash
# Round-end evidence of the final build on one B200: the default bench line, the bf16 line, the ncu launch list of host-launched
# steps, per-launch counters of every gemm_tc_kernel launch, one `--set full` capture of the selective-gate GEMM.
# Outputs under gpurun_out/final/ (summarised into profiles/ by hand).  The multi-metric pass over every gemm_tc_kernel launch of the
# run (659 launches x 5 metrics) takes ~10 minutes of box time; `-c 134` after the warm-up launches is enough for one step.
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
timeout 900 python bench.py --gemm-detail > $O/bench_default.json 2> $O/bench_default_detail.txt; tail -c 600 $O/bench_default.json
timeout 300 python bench.py --dtype bf16 --no-cpu-baseline --no-extras --no-profile > $O/bench_bf16.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-profile --no-cpu-baseline --no-extras --no-graph > $O/launches_bench.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
  --clock-control none -k regex:gemm_tc_kernel --csv --log-file $O/gemm_tc_step_metrics.csv \
  python bench.py --steps 2 --warmup 1 --no-profile --no-cpu-baseline --no-extras --no-graph > $O/gemm_metrics_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o $O/gate_gemm_after \
  python scripts/gemm_gate_probe.py 84000 1 > $O/gate_probe.log 2>&1
ls -la $O
