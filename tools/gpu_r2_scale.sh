#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus 8): weak-scaling table of the replay at N = 1, 2, 4, 8 for both workloads (ResNet-50 = BASELINE config 2,
# YOLOv5s 4096 x 3x640x640 over 8 GPUs = config 5), per-rank timing tables, the real-NCCL two-rank parity test and the two-devices-one-process test.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
run() {  # N workload extra...
  local n=$1 wl=$2; shift 2
  if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 --workload $wl --steps 20 --warmup 5 "$@"
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --workload $wl --steps 20 --warmup 5 "$@"; fi
}
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
for n in 1 2 4 8; do
  echo "== resnet50 N=$n"; run $n resnet50 --no-sweep --no-cpu-baseline > gpurun_out/r2_scale_resnet50_n$n.json 2> gpurun_out/r2_scale_resnet50_n$n.err; echo "exit $?"; cut -c1-400 gpurun_out/r2_scale_resnet50_n$n.json; tail -2 gpurun_out/r2_scale_resnet50_n$n.err | cut -c1-300
done
for n in 1 8; do
  echo "== yolov5s N=$n"; run $n yolov5s --no-sweep --no-cpu-baseline > gpurun_out/r2_scale_yolov5s_n$n.json 2> gpurun_out/r2_scale_yolov5s_n$n.err; echo "exit $?"; cut -c1-400 gpurun_out/r2_scale_yolov5s_n$n.json; tail -2 gpurun_out/r2_scale_yolov5s_n$n.err | cut -c1-300
done
python - <<'PY'
import json, glob
for wl in ('resnet50', 'yolov5s'):
    base = None
    for n in (1, 2, 4, 8):
        try: d = json.loads(open(f'gpurun_out/r2_scale_{wl}_n{n}.json').read().strip().split('\n')[-1])
        except Exception as e: continue
        if base is None: base = d['value'] / d['n_gpus']
        e2e = d.get('e2e') or {}
        print(wl, 'N', n, 'value', d['value'], 'eff', round(d['value'] / (n * base), 4), 'ms/step', d['ms_per_step'], 'exchange', d.get('exchange_ms_per_step'), 'e2e', e2e.get('value'),
              'slowest/fastest rank total', max(r['total_ms'] for r in d['per_rank']), min(r['total_ms'] for r in d['per_rank']))
PY
