#!/bin/bash
# round 2, GPU call J (final build): whole GPU suite, smoke, the driver's bench command, quantile / vs-reference lines.
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | grep -v "^  \|^$" | grep -v "^tests/.*\]$" | grep -v "Warning\|_descriptor\|ppl_caffe" | cut -c1-400 | tail -40 > gpurun_out/r2j_tests.log; tail -12 gpurun_out/r2j_tests.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench ours (driver command)"; ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench.err ); echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2j_bench_n1.json').read().strip().split(chr(10))[-1]); e=d['e2e']; print('value', d['value'], 'roofline', d['roofline']['frac'], 'e2e', e['value'], e['step_ms'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])" 2>&1 | cut -c1-1500; tail -3 gpurun_out/r2j_bench.err
echo "== vs reference kernels"; timeout 400 python tools/compare_ref_cuda.py 2>&1 | tail -9; cp gpurun_out/vs_reference_kernels.md gpurun_out/r2j_vs_reference_kernels.md 2>/dev/null
echo "== kbench quantile"; timeout 300 python tools/kbench.py --only quantile --reps 20 > gpurun_out/r2j_kbench_quantile.txt 2>&1; cat gpurun_out/r2j_kbench_quantile.txt
ncu --set full --clock-control none --import-source on -k regex:"select_" -s 14 -c 7 -f -o gpurun_out/r2_prof_select timeout 300 python tools/kbench.py --only quantile --reps 1 > gpurun_out/r2_ncu_select.log 2>&1; tail -1 gpurun_out/r2_ncu_select.log
