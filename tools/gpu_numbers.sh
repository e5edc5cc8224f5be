#!/bin/bash
# Collect the numbers quoted in DESIGN.md / profiles (lean: no ncu).
mkdir -p gpurun_out
timeout 300 python tools/kbench.py --reps 20 > gpurun_out/kbench.txt 2>&1; cat gpurun_out/kbench.txt
timeout 300 python tools/compare_ref_cuda.py 2>&1 | tail -9
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
