#!/bin/bash
# round 2, GPU call I: radix select with thresholds from a sample (cold calls), tests + timing.
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_next_rows.py tests/test_gpu_reference_own_checks.py -m gpu -q -x 2>&1 | tail -25 | cut -c1-300
echo "== select A/B"; timeout 300 python tools/select_ab.py > gpurun_out/r2i_select_ab.txt 2>&1; tail -8 gpurun_out/r2i_select_ab.txt
echo "== kbench quantile"; timeout 300 python tools/kbench.py --only quantile --reps 20 > gpurun_out/r2i_kbench_quantile.txt 2>&1; cat gpurun_out/r2i_kbench_quantile.txt
