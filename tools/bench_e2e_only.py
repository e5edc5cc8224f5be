import sys, os, json, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchvision
from ppq_b200.executor import e2e_calibration_benchmark, fuse_conv_bn
dev = torch.device('cuda', 0)
torch.backends.cudnn.benchmark = True
m = fuse_conv_bn(torchvision.models.resnet50(weights=None)).to(dev)
x = torch.rand(32, 3, 224, 224, device=dev)
with torch.no_grad():
    for _ in range(5): m(x)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): m(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
print(f'plain torch fp32 forward, batch 32: {dt*1e3:.2f} ms -> {32/dt:.0f} imgs/s per pass, {16/dt:.0f} imgs/s for two passes')
for g in (False,):
    print(json.dumps(e2e_calibration_benchmark(batch=32, steps=16, warmup=1, device=dev, graphs=g)))
