#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, rotating buffers > L2) used to choose variants and to feed ncu.

    python tools/kbench.py [--only hist,lt,lc,ft,minmax,multi,quantile,kl,ltshape] [--reps 20]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppq_b200.ffi import extension  # noqa: E402

PEAK = 6575.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timeit(fn, reps, nbuf):
    for i in range(min(nbuf, 3)): fn(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(reps): fn(i % nbuf)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='hist,lt,lc,ft,minmax')
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--n', type=int, default=32 * 12 * 512 * 512 // 2)     # 50 M elements = 201 MB
    args = ap.parse_args()
    ext = extension()
    dev = torch.device('cuda')
    n, nbuf = args.n, 6
    xs = [torch.randn(n, device=dev) for _ in range(nbuf)]
    xr = [torch.relu(x) for x in xs[:3]] + xs[:3]
    one, zero = torch.tensor([0.05], device=dev), torch.tensor([0.0], device=dev)
    rows = []

    def report(name, secs, bytes_per_elem):
        gbs = bytes_per_elem * n / secs / 1e9
        rows.append((name, secs * 1e6, gbs, gbs / PEAK))
        print(f'{name:48s} {secs*1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK:6.1%} of {PEAK:.0f}', flush=True)

    only = set(args.only.split(','))
    if 'minmax' in only:
        mm = torch.empty(2, device=dev); ext.MinMax_Init(mm[0:1], mm[1:2])
        report('minmax_t', timeit(lambda i: ext.MinMax_T(xs[i], mm), args.reps, nbuf), 4)
    if 'hist' in only:
        h = torch.zeros(4096, dtype=torch.int32, device=dev)
        hs = float(xs[0].abs().max().item()) / 4096
        for var in (0, 7, 8, 5):
            ext.set_variant('histogram', var)
            report(f'histogram_t var{var} randn', timeit(lambda i: ext.Histogram_T(xs[i], hs, True, h), args.reps, nbuf), 4)
            report(f'histogram_t var{var} relu/randn mix', timeit(lambda i: ext.Histogram_T(xr[i], hs, True, h), args.reps, nbuf), 4)
            report(f'histogram_t var{var} relu', timeit(lambda i: ext.Histogram_T(xr[i % 3], hs, True, h), args.reps, 3), 4)
        ext.set_variant('histogram', 0)
        report('histogram_asym_t randn', timeit(lambda i: ext.Histogram_Asymmetric_T(-5.0, 5.0, xs[i], True, h), args.reps, nbuf), 4)
        h2 = torch.zeros(2048, dtype=torch.int32, device=dev)
        report('histogram_t 2048 bins randn', timeit(lambda i: ext.Histogram_T(xs[i], hs * 2, True, h2), args.reps, nbuf), 4)
    if 'ltshape' in only:
        for shape_n in (8388608, 12582912, 802816):
            xs2 = [torch.randn(shape_n, device=dev) for _ in range(max(2, int(1.5e9 // (8 * shape_n))))]
            for var in (0, 2, 3):
                ext.set_variant('linear_quant_t', var)
                secs = timeit(lambda i: ext.QuantizeTensor_LT(xs2[i], one, zero, -128, 127, 0), 100, len(xs2))
                print(f'LT n={shape_n} var{var}: {secs*1e6:8.2f} us  {8*shape_n/secs/1e9:8.1f} GB/s  {8*shape_n/secs/1e9/PEAK:6.1%}', flush=True)
            ext.set_variant('linear_quant_t', 0)
            del xs2
    if 'lt' in only:
        for var in (0, 1, 2, 3):
            ext.set_variant('linear_quant_t', var)
            report(f'linear_quant_t var{var}', timeit(lambda i: ext.QuantizeTensor_LT(xs[i], one, zero, -128, 127, 0), args.reps, nbuf), 8)
        ext.set_variant('linear_quant_t', 0)
        report('linear_quant_t ROUND_HALF_UP (compile-time mode)', timeit(lambda i: ext.QuantizeTensor_LT(xs[i], one, zero, -128, 127, 1), args.reps, nbuf), 8)
        report('linear_quant_t ROUND_HALF_FAR_FROM_ZERO (run-time mode)', timeit(lambda i: ext.QuantizeTensor_LT(xs[i], one, zero, -128, 127, 4), args.reps, nbuf), 8)
        report('linear_quant_t toInt8', timeit(lambda i: ext.QuantizeTensor_toInt(xs[i], one, zero, -128, 127, -1000, 0, 8), args.reps, nbuf), 5)
    if 'quantile' in only:
        # algorithmic traffic of the percentile observer is 4 B/element; the select streams the tensor twice
        report('quantile_t q=0.9999 randn (2 passes + finish)', timeit(lambda i: ext.Quantile_T(xs[i], 0.9999), args.reps, nbuf), 4)
        report('quantile_t q=0.9999 relu (zero bucket refined)', timeit(lambda i: ext.Quantile_T(xr[i % 3], 0.9999), args.reps, 3), 4)
        sizes = [n // 16] * 16
        parts = [[x[k * (n // 16):(k + 1) * (n // 16)] for k in range(16)] for x in xs]
        descs = [torch.tensor([[t.data_ptr(), t.numel(), k] for k, t in enumerate(p)], dtype=torch.int64, device=dev) for p in parts]
        cap = 1 << 16
        ws = torch.empty(ext.Multi_Quantile_Workspace_Bytes(16, cap), dtype=torch.uint8, device=dev)
        out = torch.zeros(16, 2, device=dev)
        report('multi_quantile_t 16 tensors of n/16 (one table)', timeit(lambda i: ext.Multi_Quantile_T(descs[i], n // 16, 0.9999, out, 2, ws, cap, None), args.reps, nbuf), 4)
        guess = ext.Quantile_Guess_Init(16, out)
        report('  same, speculative (thresholds from the previous call)', timeit(lambda i: ext.Multi_Quantile_T(descs[i], n // 16, 0.9999, out, 2, ws, cap, guess), args.reps, nbuf), 4)
    if 'lc' in only:
        for shape, axis in (((n // 4608, 4608), 0), ((n // 512, 512), 0), ((n // 64, 64), 0), ((8, n // 8 // 3136, 3136), 1), ((n // 9, 9), 0), ((n // 768, 768), 1)):
            C = shape[axis]
            s = torch.rand(C, device=dev) * 0.1 + 0.01; o = torch.zeros(C, device=dev)
            m = 1
            for d in shape: m *= d
            vs = [x[:m].view(shape) for x in xs]
            secs = timeit(lambda i: ext.QuantizeTensor_LC(vs[i], s, o, -128, 127, axis, 0), args.reps, nbuf)
            gbs = 8 * m / secs / 1e9
            print(f'{"linear_quant_c " + str(shape) + " axis " + str(axis):48s} {secs*1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK:6.1%}', flush=True)
    if 'kl' in only:
        for T in (1, 106):
            hist = torch.poisson(torch.full((T, 4096), 50.0, device=dev) * torch.linspace(2, 0.01, 4096, device=dev)).to(torch.int32)
            hs = torch.full((T,), 0.001, device=dev)
            secs = timeit(lambda i: ext.KL_Search(hist, 4096, hs, None, 8, False, 1e-8), args.reps, 1)
            print(f'{"kl_search " + str(T) + " x 4096-bin histograms":48s} {secs*1e6:9.1f} us', flush=True)
    if 'multi' in only:
        # every Conv/Linear weight of a network in one launch; copies of the table rotate so that the weights come from HBM, not L2
        import torchvision
        from ppq_b200.calibration import MultiWeightQuantizer
        for name, ctor, copies in (('resnet50', torchvision.models.resnet50, 3), ('mobilenet_v2', torchvision.models.mobilenet_v2, 12)):
            shapes = [tuple(m.weight.shape) for m in ctor(weights=None).modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
            qs = []
            for _ in range(copies):
                ws = [torch.randn(sh, device=dev) * 0.05 for sh in shapes]
                sc = [w.flatten(1).abs().amax(1) / 127 for w in ws]
                qs.append(MultiWeightQuantizer(ws, sc, [torch.zeros_like(v) for v in sc], channel_axis=0))
            m = sum(w.numel() for w in qs[0].weights)
            secs = timeit(lambda i: qs[i](), args.reps, copies)
            gbs = 8 * m / secs / 1e9
            print(f'{"multi_linear_quant_c " + name + f" {len(shapes)} tensors {m/1e6:.1f} M el":48s} {secs*1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK:6.1%}', flush=True)
            one_by_one = timeit(lambda i: [ext.QuantizeTensor_LC(w, s_, o_, -128, 127, 0, 0) for w, s_, o_ in zip(qs[i].weights, qs[i].scales, qs[i].offsets)], max(args.reps // 4, 2), copies)
            print(f'{"  same, one QuantizeTensor_LC launch per tensor":48s} {one_by_one*1e6:9.1f} us', flush=True)
    if 'ft' in only:
        report('float_quant_t E4M3', timeit(lambda i: ext.QuantizeTensor_FT(xs[i], torch.ones(1, device=dev), zero, 4, 3, -448.0, 448.0, 0), args.reps, nbuf), 8)
        report('float_quant_t E4M3 mode1', timeit(lambda i: ext.QuantizeTensor_FT(xs[i], torch.ones(1, device=dev), zero, 4, 3, -448.0, 448.0, 1), args.reps, nbuf), 8)
        C = 768
        v = [x[:(n // C) * C].view(-1, C) for x in xs]
        s = torch.ones(C, device=dev); o = torch.zeros(C, device=dev)
        report('float_quant_c E4M3 [*,768] axis 1', timeit(lambda i: ext.QuantizeTensor_FC(v[i], s, o, 4, 3, -448.0, 448.0, 1, 0), args.reps, nbuf), 8)


if __name__ == '__main__':
    main()
