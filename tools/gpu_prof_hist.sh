#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$" | cut -c1-300 | tail -25
ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 3 -c 1 -f -o gpurun_out/prof_hist2 python tools/kbench.py --only hist --reps 1 > gpurun_out/ncu_hist2.log 2>&1; tail -2 gpurun_out/ncu_hist2.log
