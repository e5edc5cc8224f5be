#!/bin/bash
# Lean ncu pass: launch list of the timed region (NVTX-filtered) + full captures of the three kernels of a calibration step + LT.
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --rotate 2 --no-e2e --no-cpu-baseline"
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
for k in multi_histogram multi_minmax multi_channel; do
  timeout 300 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:$k -c 1 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ew_tensor_kernel -s 3 -c 1 -f -o gpurun_out/prof_lt2 python tools/kbench.py --only lt --reps 1 > gpurun_out/ncu_lt2.log 2>&1; tail -1 gpurun_out/ncu_lt2.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 3 -c 1 -f -o gpurun_out/prof_hist3 python tools/kbench.py --only hist --reps 1 > gpurun_out/ncu_hist3.log 2>&1; tail -1 gpurun_out/ncu_hist3.log | cut -c1-200
ls -la gpurun_out | grep -E "prof_|launches"
