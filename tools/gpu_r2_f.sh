#!/bin/bash
# round 2, GPU call F: e2e through the persistent device ring + allocator / CPU-time diagnostics of the slowest calibration.
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -3
for i in 1 2; do
echo "== e2e, 12 timed calibrations, run $i"; timeout 300 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline --e2e-steps 12 > gpurun_out/r2f_bench_e2e_$i.json 2> gpurun_out/r2f_bench_e2e_$i.err; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_e2e_$i.json').read().strip().split(chr(10))[-1]); print(json.dumps(d['e2e']))" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2f_bench_e2e_$i.err
done
echo "== driver command"; ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench.err ); python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], json.dumps(d['e2e']))" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2f_bench.err
echo "== yolov5s"; timeout 600 python bench.py --workload yolov5s --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2f_bench_yolo_n1.json 2> gpurun_out/r2f_bench_yolo.err; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_yolo_n1.json').read().strip().split(chr(10))[-1]); print('value', d['value'], 'e2e', json.dumps(d['e2e']))" 2>&1 | cut -c1-2500; tail -3 gpurun_out/r2f_bench_yolo.err
