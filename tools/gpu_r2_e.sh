#!/bin/bash
# round 2, GPU call E: what call D lost to a NameError in the e2e trace (both bench lines), the new e2e-arm test, select after the U=8 revert.
mkdir -p gpurun_out
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_next_rows.py -m gpu -q 2>&1 | tail -3
echo "== bench ours (driver command)"; ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench.err ); echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'roofline', d['roofline']['frac'], 'e2e', json.dumps(d['e2e']), 'cpu', json.dumps(d['cpu_baseline'])[:300], 'sweep', json.dumps(d.get('fakequant'))[:2500])" 2>&1 | cut -c1-6000; tail -5 gpurun_out/r2d_bench.err
echo "== kbench quantile"; timeout 300 python tools/kbench.py --only quantile --reps 20 > gpurun_out/r2e_kbench_quantile.txt 2>&1; cat gpurun_out/r2e_kbench_quantile.txt
echo "== bench yolov5s"; timeout 600 python bench.py --workload yolov5s --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2d_bench_yolo_n1.json 2> gpurun_out/r2d_bench_yolo.err; echo "yolo exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_yolo_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'e2e', json.dumps(d['e2e']))" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2d_bench_yolo.err
echo "== e2e again, 10 timed calibrations (spread)"; timeout 300 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline --e2e-steps 10 > gpurun_out/r2e_bench_e2e10.json 2> gpurun_out/r2e_bench_e2e10.err; python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_e2e10.json').read().strip().split(chr(10))[-1]); print(json.dumps(d['e2e']))" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2e_bench_e2e10.err
ncu --set full --clock-control none --import-source on -k regex:select_pass -s 8 -c 3 -f -o gpurun_out/r2_prof_select timeout 300 python tools/kbench.py --only quantile --reps 1 > gpurun_out/r2_ncu_select.log 2>&1; tail -1 gpurun_out/r2_ncu_select.log
