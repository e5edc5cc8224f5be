#!/bin/bash
# round 2, GPU call K: select finish / init with branch-free counting -- tests of every quantile path, then timing.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -x 2>&1 | tail -2
timeout 100 python tools/kbench.py --only quantile --reps 20 > gpurun_out/r2k_kbench_quantile.txt 2>&1; cat gpurun_out/r2k_kbench_quantile.txt
