"""The reference's USING_CUDA_KERNEL=False calibration pipeline, end to end on the host cores -- TEST / BASELINE INFRASTRUCTURE.

Used only by bench.py (`--impl reference`, and as a cross-check in tests).  It restates, with torch CPU ops, what
RuntimeCalibrationPass does on the CPU path for a torchvision network (the reference itself needs an ONNX graph, which cannot be
built here: no `onnx`):
  * BatchNorm folded into the convolutions (ppq/core/common.py:41),
  * per-channel INT8 weight fake-quant on EVERY forward, torch formulation (ppq/quantization/qfunction/linear.py:73-81),
    scales from per-channel min-max (observer/range.py:93-98, 120-135),
  * phase 1: value.min() / value.max() appended per batch (range.py:85-100), rendered with minmax_to_scale_offset,
  * phase 2: torch.histc(abs(x), 4096, 0, hist_scale * 4096) accumulated (range.py:183), KL search on the CPU (range.py:190-282),
  * conv -> relu fusion: the conv output is not observed (QuantizeFusionPass), same observed set as ppq_b200.executor.
"""
import time

import torch

from . import kl_search, minmax_to_scale_offset, torch_cpu_hist_sym, torch_cpu_linear_quant_c, torch_cpu_minmax

BINS = 4096
_KINDS = (torch.nn.Conv2d, torch.nn.Linear, torch.nn.ReLU, torch.nn.MaxPool2d, torch.nn.AdaptiveAvgPool2d)


def _fuse_bn(model):
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    model.eval()

    def walk(mod):
        prev_name, prev = None, None
        for name, child in list(mod.named_children()):
            if isinstance(child, torch.nn.BatchNorm2d) and isinstance(prev, torch.nn.Conv2d):
                setattr(mod, prev_name, fuse_conv_bn_eval(prev, child)); setattr(mod, name, torch.nn.Identity())
                prev_name, prev = None, None
                continue
            walk(child)
            prev_name, prev = name, child
    walk(model)
    return model


class CpuCalibrator:
    def __init__(self, model, example):
        self.model = _fuse_bn(model)
        self.phase, self.k, self.trace = 0, 0, True
        self.skip, self.last = set(), None
        self.mins, self.maxs, self.hists, self.hist_scale = {}, {}, {}, {}
        self.wparams = {}
        for m in self.model.modules():
            if isinstance(m, _KINDS):
                m.register_forward_pre_hook(self._pre); m.register_forward_hook(self._post)
            if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)):
                w = m.weight.data
                flat = w.reshape(w.shape[0], -1)
                sc = []
                for lo, hi in zip(flat.min(dim=1)[0].tolist(), flat.max(dim=1)[0].tolist()):   # per-channel python loop, as upstream
                    sc.append(minmax_to_scale_offset(lo, hi, -128, 127, True)[0])
                self.wparams[id(m)] = (torch.tensor(sc, dtype=torch.float32), torch.zeros(len(sc)))
        with torch.no_grad():
            self.k = 0; self.model(example)
        self.trace = False

    def _observe(self, t):
        k = self.k; self.k += 1
        if self.trace: return
        if self.phase == 1:
            lo, hi = torch_cpu_minmax(t)
            self.mins.setdefault(k, []).append(lo.reshape(1)); self.maxs.setdefault(k, []).append(hi.reshape(1))
        elif self.phase == 2:
            h = torch_cpu_hist_sym(t, self.hist_scale[k], BINS)
            if k in self.hists: self.hists[k] += h
            else: self.hists[k] = h

    def _pre(self, m, args):
        x = args[0]
        if self.trace:
            if isinstance(m, torch.nn.ReLU) and self.last is not None and self.last[1] is x and x._version == self.last[2] and self.last[3]:
                self.skip.add(self.last[0])
            return None
        if self.first_op is m and self.first_pending:
            self.first_pending = False
            self._observe(x)
        if id(m) in self.wparams:
            s, o = self.wparams[id(m)]
            m.__dict__['_fp32_w'] = m.weight.data
            m.weight.data = torch_cpu_linear_quant_c(m.weight.data, s, o, 0, -128, 127)

    def _post(self, m, args, out):
        if self.trace:
            if not hasattr(self, 'first_op'): self.first_op = m
            key = (id(m), self.k)
            self.last = (key, out, out._version, isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)))
            self.order = getattr(self, 'order', []) + [key]
            self.k += 1
            return None
        w = m.__dict__.pop('_fp32_w', None)
        if w is not None: m.weight.data = w
        pos = self.pos; self.pos += 1
        if self.order[pos] not in self.skip: self._observe(out)

    @torch.no_grad()
    def forward(self, x):
        self.k, self.pos, self.first_pending = 0, 0, True
        return self.model(x)

    @torch.no_grad()
    def calibrate(self, batches):
        self.phase = 1
        for x in batches: self.forward(x)
        for k in self.mins:
            lo = torch.min(torch.cat(self.mins[k])).item(); hi = torch.max(torch.cat(self.maxs[k])).item()
            self.hist_scale[k] = float(max(abs(hi), abs(lo))) / BINS
        self.phase = 2
        for x in batches: self.forward(x)
        return {k: kl_search(self.hists[k], self.hist_scale[k], 8)[0] for k in sorted(self.hists)}


def resnet50_cpu_calibration(batch: int, steps: int, seed: int = 0, threads: int = None):
    """Returns (imgs/s, seconds, number of observed tensors) for `steps` calibration batches (both phases + KL search)."""
    import torchvision
    if threads: torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model = torchvision.models.resnet50(weights=None)
    cal = CpuCalibrator(model, torch.zeros(1, 3, 224, 224))
    g = torch.Generator().manual_seed(seed + 1)
    data = [torch.rand(batch, 3, 224, 224, generator=g) for _ in range(steps)]
    t0 = time.perf_counter()
    scales = cal.calibrate(data)
    secs = time.perf_counter() - t0
    return steps * batch / secs, secs, len(scales)
