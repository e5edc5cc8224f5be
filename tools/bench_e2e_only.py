import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppq_b200.executor import e2e_calibration_benchmark
for g in (False, True):
    print(json.dumps(e2e_calibration_benchmark(batch=32, steps=8, warmup=1, device=torch.device('cuda', 0), graphs=g)))
