#!/bin/bash
# ncu captures of the top kernels (one GPU; never multi-rank).  Outputs -> gpurun_out/.
mkdir -p gpurun_out
python tools/kbench.py --reps 10 > gpurun_out/kbench.txt 2>&1; cat gpurun_out/kbench.txt
ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 3 -c 2 -f -o gpurun_out/prof_hist python tools/kbench.py --only hist --reps 1 > gpurun_out/ncu_hist.log 2>&1; tail -2 gpurun_out/ncu_hist.log
ncu --set full --clock-control none --import-source on -k regex:ew_tensor_kernel -s 3 -c 1 -f -o gpurun_out/prof_lt python tools/kbench.py --only lt --reps 1 > gpurun_out/ncu_lt.log 2>&1; tail -2 gpurun_out/ncu_lt.log
ls -la gpurun_out/
