#!/bin/bash
# round 2, GPU call D (final build): GPU suite, smoke, all kernel micro-benchmarks, ours-vs-reference kernels, both bench arms, yolov5s, the
# NVTX-filtered ncu launch list of the timed region and full ncu captures of the kernels that changed this round.  Everything -> gpurun_out/r2d_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2d_bench_ref_n1.json 2> gpurun_out/r2d_bench_ref.err; echo "ref exit $?"; cut -c1-600 gpurun_out/r2d_bench_ref_n1.json; tail -3 gpurun_out/r2d_bench_ref.err
echo "== bench ours (driver command)"; ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench.err ); echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'roofline', d['roofline']['frac'], 'e2e', json.dumps(d['e2e']), 'cpu', json.dumps(d['cpu_baseline'])[:300], 'sweep', json.dumps(d.get('fakequant'))[:1500])" 2>&1 | cut -c1-5000; tail -5 gpurun_out/r2d_bench.err
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | grep -v "^  \|^$" | grep -v "^tests/.*\]$" | cut -c1-400 | tail -60 > gpurun_out/r2d_tests.log; tail -12 gpurun_out/r2d_tests.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== kbench"; timeout 600 python tools/kbench.py --only minmax,hist,lt,lc,ft,quantile,kl,multi --reps 20 > gpurun_out/r2d_kbench.txt 2>&1; cat gpurun_out/r2d_kbench.txt
echo "== kbench kl variant 1 (round-1 serial candidates)"; timeout 200 python - <<'PY' > gpurun_out/r2d_kbench_kl_old.txt 2>&1
import sys, os
sys.path.insert(0, os.getcwd())
from ppq_b200.ffi import extension
ext = extension(); ext.set_variant('kl_search', 1)
sys.argv = ['kbench', '--only', 'kl', '--reps', '20']
import runpy; runpy.run_path('tools/kbench.py', run_name='__main__')
PY
cat gpurun_out/r2d_kbench_kl_old.txt
echo "== vs reference kernels"; timeout 400 python tools/compare_ref_cuda.py 2>&1 | tail -10; cp gpurun_out/vs_reference_kernels.md gpurun_out/r2d_vs_reference_kernels.md 2>/dev/null
echo "== bench yolov5s"; timeout 600 python bench.py --workload yolov5s --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2d_bench_yolo_n1.json 2> gpurun_out/r2d_bench_yolo.err; echo "yolo exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_yolo_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'e2e', json.dumps(d['e2e']))" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2d_bench_yolo.err
echo "== ncu launch list (timed NVTX range only)"; ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches.csv timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-sweep --no-cpu-baseline > gpurun_out/r2d_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2d_launches.csv; tail -2 gpurun_out/r2d_bench_under_ncu.log | cut -c1-300
echo "== ncu captures"
ncu --set full --clock-control none --import-source on -k regex:select_pass -s 8 -c 3 -f -o gpurun_out/r2_prof_select timeout 300 python tools/kbench.py --only quantile --reps 1 > gpurun_out/r2_ncu_select.log 2>&1; tail -1 gpurun_out/r2_ncu_select.log
ncu --set full --clock-control none --import-source on -k regex:multi_select_pass0_spec -s 3 -c 1 -f -o gpurun_out/r2_prof_select_spec timeout 300 python tools/kbench.py --only quantile --reps 1 > gpurun_out/r2_ncu_select_spec.log 2>&1; tail -1 gpurun_out/r2_ncu_select_spec.log
ncu --set full --clock-control none --import-source on -k regex:ew_channel_table -s 2 -c 4 -f -o gpurun_out/r2_prof_lc_table timeout 300 python tools/kbench.py --only lc --reps 1 > gpurun_out/r2_ncu_lc.log 2>&1; tail -1 gpurun_out/r2_ncu_lc.log
ncu --set full --clock-control none --import-source on -k regex:kl_search -s 1 -c 1 -f -o gpurun_out/r2_prof_kl timeout 300 python tools/kbench.py --only kl --reps 1 > gpurun_out/r2_ncu_kl.log 2>&1; tail -1 gpurun_out/r2_ncu_kl.log
ls -la gpurun_out | grep "r2d_\|r2_prof"
