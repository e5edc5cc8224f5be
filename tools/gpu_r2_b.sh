#!/bin/bash
# round 2, GPU call B: full GPU suite again (3 fixes + new YOLOv5s / Concat / graphwise tests), select + histogram micro-benchmarks after the
# integer-pipe fixes, default bench (e2e with the persistent descriptor stager), NHWC A/B, yolov5s with its e2e arm.
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | grep -v "^  \|^$" | grep -v "^tests/.*\]$" | cut -c1-400 | tail -60 > gpurun_out/r2b_tests.log; tail -40 gpurun_out/r2b_tests.log
echo "== kbench"; timeout 300 python tools/kbench.py --only hist,quantile,lc,kl --reps 20 > gpurun_out/r2b_kbench.txt 2>&1; cat gpurun_out/r2b_kbench.txt
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'e2e', json.dumps(d['e2e']), 'cpu', json.dumps(d['cpu_baseline'])[:300])" 2>&1 | cut -c1-2500; tail -5 gpurun_out/r2b_bench.err
echo "== bench e2e NHWC A/B"; timeout 400 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline --e2e-channels-last > gpurun_out/r2b_bench_e2e_nhwc.json 2> gpurun_out/r2b_bench_e2e_nhwc.err; echo "nhwc exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_e2e_nhwc.json').read().strip().split(chr(10))[-1]); print(json.dumps(d['e2e']))" 2>&1 | cut -c1-1500; tail -3 gpurun_out/r2b_bench_e2e_nhwc.err
echo "== bench yolov5s"; timeout 900 python bench.py --workload yolov5s --steps 10 --warmup 3 --no-sweep > gpurun_out/r2b_bench_yolo_n1.json 2> gpurun_out/r2b_bench_yolo.err; echo "yolo exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_yolo_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'e2e', json.dumps(d['e2e']), 'cpu', json.dumps(d['cpu_baseline'])[:300])" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2b_bench_yolo.err
echo "== yolov5s reference arm"; timeout 500 python bench.py --impl reference --workload yolov5s --steps 6 --warmup 1 --ref-budget 120 > gpurun_out/r2b_bench_yolo_ref.json 2> gpurun_out/r2b_bench_yolo_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/r2b_bench_yolo_ref.json; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_yolo_ref.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], json.dumps(d['arm'])[:600])"
echo "== ncu launch list (timed NVTX range only)"; ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-sweep --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches.csv
ncu --set full --clock-control none --import-source on -k regex:select_pass -s 8 -c 3 -f -o gpurun_out/r2_prof_select timeout 300 python tools/kbench.py --only quantile --reps 1 > gpurun_out/r2_ncu_select.log 2>&1; tail -1 gpurun_out/r2_ncu_select.log
ncu --set full --clock-control none --import-source on -k regex:ew_channel_table -s 2 -c 2 -f -o gpurun_out/r2_prof_lc_table timeout 300 python tools/kbench.py --only lc --reps 1 > gpurun_out/r2_ncu_lc.log 2>&1; tail -1 gpurun_out/r2_ncu_lc.log
