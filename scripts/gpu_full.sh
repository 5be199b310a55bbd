#!/bin/bash
# the whole GPU suite, then an env A/B of bench.py (usage: gpu_full.sh [VAR v1 v2 ...]) and the timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
if [ -n "$1" ]; then bash scripts/ab_env.sh "$@"; fi
