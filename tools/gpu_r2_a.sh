#!/bin/bash
# round 2, GPU call A: the whole GPU suite (new: graph-level parity with the real reference package, 2-pass select, arena percentile / mse),
# smoke, quantile micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | grep -v "^  \|^$\|Warning\|warn" | cut -c1-400 | tail -80 > gpurun_out/r2a_tests.log; tail -60 gpurun_out/r2a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/kbench.py --only quantile,kl --reps 20 > gpurun_out/r2a_kbench.txt 2>&1; cat gpurun_out/r2a_kbench.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; cut -c1-6000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget 40 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref exit $?"; cut -c1-2500 gpurun_out/r2a_bench_ref.json; tail -3 gpurun_out/r2a_bench_ref.err
