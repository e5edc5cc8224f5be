"""world_size-2 gloo tests (CPU) of the multi-GPU exchange logic: sample sharding, the MAX all-reduce of {-min, max} and the SUM
all-reduce of the histogram arena reproduce the single-process statistics bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ppq_b200.calibration import allreduce_hist, allreduce_minmax, shard_indices
    T, bins, samples = 5, 64, 12
    g = torch.Generator().manual_seed(123)
    data = [[torch.randn(257, generator=g) * (t + 1) for t in range(T)] for _ in range(samples)]     # identical on every rank
    mine = list(shard_indices(samples, rank, world))
    minmax = torch.empty(T, 2); minmax[:, 0] = float('inf'); minmax[:, 1] = float('-inf')
    for i in mine:
        for t in range(T):
            minmax[t, 0] = min(minmax[t, 0], data[i][t].min()); minmax[t, 1] = max(minmax[t, 1], data[i][t].max())
    allreduce_minmax(minmax)
    hs = torch.maximum(minmax[:, 0].abs(), minmax[:, 1].abs()) / bins
    hist = torch.zeros(T, bins, dtype=torch.int32)
    for i in mine:
        for t in range(T):
            b = torch.floor(data[i][t].abs() / hs[t]).long()
            b = b[b <= bins - 1]
            hist[t] += torch.bincount(b, minlength=bins).int()
    allreduce_hist(hist)
    from ppq_b200.calibration import gather_in_sample_order
    pairs = torch.stack([torch.stack([torch.stack([data[i][t].max(), data[i][t].min()]) for i in mine]) for t in range(T)])   # [T, n_local, 2]
    ordered = gather_in_sample_order(pairs)
    torch.save({'minmax': minmax, 'hist': hist, 'mine': mine, 'ordered': ordered}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_equal_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / 'r0.pt'), torch.load(tmp_path / 'r1.pt')
    assert sorted(r0['mine'] + r1['mine']) == list(range(12)) and not set(r0['mine']) & set(r1['mine'])
    assert torch.equal(r0['minmax'], r1['minmax']) and torch.equal(r0['hist'], r1['hist'])
    # single-process recomputation
    T, bins = 5, 64
    g = torch.Generator().manual_seed(123)
    data = [[torch.randn(257, generator=g) * (t + 1) for t in range(T)] for _ in range(12)]
    mm = torch.tensor([[min(d[t].min() for d in data), max(d[t].max() for d in data)] for t in range(T)])
    assert torch.equal(r0['minmax'], mm)
    hs = torch.maximum(mm[:, 0].abs(), mm[:, 1].abs()) / bins
    hist = torch.zeros(T, bins, dtype=torch.int32)
    for d in data:
        for t in range(T):
            b = torch.floor(d[t].abs() / hs[t]).long(); b = b[b <= bins - 1]
            hist[t] += torch.bincount(b, minlength=bins).int()
    assert torch.equal(r0['hist'], hist)
    want = torch.stack([torch.stack([torch.stack([d[t].max(), d[t].min()]) for d in data]) for t in range(T)])
    assert torch.equal(r0['ordered'], want) and torch.equal(r1['ordered'], want)        # global sample order restored on every rank
    assert torch.equal(r0['ordered'].mean(dim=1), want.mean(dim=1))                     # hence the fp32 mean is bit-identical


def test_pack_unpack_minmax_and_shards():
    sys.path.insert(0, ROOT)
    from ppq_b200.calibration import pack_minmax_for_max_reduce, shard_indices, unpack_minmax_after_max_reduce
    mm = torch.tensor([[-1.5, 2.0], [0.0, 0.0], [3.0, 7.0], [float('inf'), float('-inf')]])
    packed = pack_minmax_for_max_reduce(mm)
    assert torch.equal(packed[:, 0], -mm[:, 0]) and torch.equal(packed[:, 1], mm[:, 1])
    other = torch.tensor([[-4.0, 1.0], [-0.5, 0.25], [2.0, 9.0], [5.0, 6.0]])
    red = torch.maximum(packed, pack_minmax_for_max_reduce(other))
    out = unpack_minmax_after_max_reduce(red, torch.empty_like(mm))
    assert torch.equal(out, torch.tensor([[-4.0, 2.0], [-0.5, 0.25], [2.0, 9.0], [5.0, 6.0]]))
    for n, w in ((4096, 8), (13, 4), (8, 8), (3, 8)):
        got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
        assert got == list(range(n))
