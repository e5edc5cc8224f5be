"""Where does the end-to-end calibration step go?  Host enqueue time vs device time per forward, with / without H2D prefetch."""
import sys, os, json, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchvision
from ppq_b200.executor import TorchExecutor, calibrate_arena
from ppq_b200.core import QuantizationStates
dev = torch.device('cuda', 0)
torch.backends.cudnn.benchmark = True
model = torchvision.models.resnet50(weights=None).eval()
ex = TorchExecutor(model.to(dev), torch.zeros(2, 3, 224, 224, device=dev))
ex.quantize_parameters()
host = [torch.rand(32, 3, 224, 224).pin_memory() for _ in range(16)]
cfgs = ex.observed_configs()
x = host[0].to(dev)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    cpu = (time.perf_counter() - t) / reps
    b.record(); torch.cuda.synchronize()
    return cpu * 1e3, a.elapsed_time(b) / reps
from ppq_b200.executor import fuse_conv_bn
plain = fuse_conv_bn(torchvision.models.resnet50(weights=None)).to(dev)
with torch.no_grad():
    print('plain model forward            host %.2f ms  device %.2f ms' % timed(lambda: plain(x)))
print('executor forward (weights q)   host %.2f ms  device %.2f ms' % timed(lambda: ex.forward(x)))
print('executor forward + collect     host %.2f ms  device %.2f ms' % timed(lambda: ex.forward(x, collect=True)))
print('H2D copy of one batch          host %.2f ms  device %.2f ms' % timed(lambda: host[1].to(dev, non_blocking=True)))
for pf in (False, True):
    for d in (False, 'auto'):
        def run():
            for c in cfgs: c.state = QuantizationStates.INITIAL
            cal = calibrate_arena(ex, host, method='kl', to_device=lambda t: t.to(dev, non_blocking=True), deferred=d, prefetch=pf)
            return cal.scale.cpu()
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f'calibrate_arena prefetch={pf} deferred={d}: {ms:.1f} ms -> {16 * 32 / ms * 1e3:.0f} imgs/s')
