#!/usr/bin/env python
"""A/B of the single-tensor radix select's launch shapes (variant key "select", variants.h): 50 M-element tensors, q = 0.9999, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from kbench import timeit  # noqa: E402
from ppq_b200.ffi import extension  # noqa: E402

ext = extension()
n = 32 * 12 * 512 * 512 // 2
xs = [torch.randn(n, device='cuda') for _ in range(6)]
xr = [torch.relu(x) for x in xs[:3]]
want = [ext.Quantile_T(x, 0.9999).clone() for x in xs]
for var in (0, 64, 64 + 3, 64 + 24, 0):                                # 64 = no thresholds from a sample (the plain two-pass route)
    ext.set_variant('select', var)
    ok = all(torch.equal(ext.Quantile_T(x, 0.9999), w) for x, w in zip(xs, want))
    t = timeit(lambda i: ext.Quantile_T(xs[i], 0.9999), 30, 6)
    tr = timeit(lambda i: ext.Quantile_T(xr[i % 3], 0.9999), 30, 3)
    print(f'select variant {var:2d} (sample {"off" if var & 64 else "on"}, pass0 {var & 7}, pass1 {(var >> 3) & 7}): randn {t * 1e6:7.1f} us   relu {tr * 1e6:7.1f} us   same result: {ok}', flush=True)
ext.set_variant('select', 0)
