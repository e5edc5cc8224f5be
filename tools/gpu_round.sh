#!/bin/bash
# End-of-iteration GPU check: tests, smoke, bench (both arms), ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$" | cut -c1-300 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-600 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:multi_histogram -s 4 -c 1 -f -o gpurun_out/prof_multi_hist timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_mh.log 2>&1; tail -1 gpurun_out/ncu_mh.log
ncu --set full --clock-control none --import-source on -k regex:ew_tensor_kernel -s 3 -c 1 -f -o gpurun_out/prof_lt python tools/kbench.py --only lt --reps 1 > gpurun_out/ncu_lt.log 2>&1; tail -1 gpurun_out/ncu_lt.log
ls -la gpurun_out | head -30
