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


def _reduce_worker(rank, world, port, out_dir, starve_rank):
    """RuntimeCalibrationPass._reduce on observers with hand-made statistics (the collectors themselves need a GPU; the exchange does not)."""
    import types
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ppq_b200 import LinearQuantizationConfig
    from ppq_b200.calibration import RuntimeCalibrationPass
    from ppq_b200.observer import TorchHistObserver, TorchMinMaxObserver, TorchPercentileObserver
    g = torch.Generator().manual_seed(7 + rank)

    def slot(channels=None, bins=0):
        lo, hi = -torch.rand(1, generator=g) * (rank + 1), torch.rand(1, generator=g) * (3 - rank)
        s = types.SimpleNamespace(minmax=torch.cat([lo, hi]), cmins=None, cmaxs=None, hist=None)
        if channels: s.cmins, s.cmaxs = -torch.rand(channels, generator=g) * (rank + 1), torch.rand(channels, generator=g) * (2 - rank)
        if bins: s.hist = torch.randint(0, 50, (bins,), generator=g, dtype=torch.int32)
        return s

    def make(cls, cfg, s, observed=3):
        ob = object.__new__(cls)
        ob._quant_cfg, ob._slot, ob._observed = cfg, s, (0 if rank == starve_rank else observed)
        return ob
    t_obs = make(TorchMinMaxObserver, LinearQuantizationConfig(), slot())
    c_obs = make(TorchMinMaxObserver, LinearQuantizationConfig(channel_axis=1), slot(channels=6))
    h_obs = make(TorchHistObserver, LinearQuantizationConfig(calibration='kl'), slot(bins=32))
    p_obs = object.__new__(TorchPercentileObserver)
    p_obs._quant_cfg = LinearQuantizationConfig(calibration='percentile')
    p_obs._percentile_collector = [] if rank == starve_rank else [torch.tensor([[10.0 * rank + j, -(10.0 * rank + j)]]) for j in range(2)]
    table = {1: t_obs, 2: c_obs, 3: h_obs, 4: p_obs}
    pas = RuntimeCalibrationPass(method=None)
    pas._observers = {'op': types.SimpleNamespace(hook=types.SimpleNamespace(_observer_table=table))}
    before = {'t': t_obs._slot.minmax.clone(), 'cmin': c_obs._slot.cmins.clone(), 'cmax': c_obs._slot.cmaxs.clone(),
              'h_mm': h_obs._slot.minmax.clone(), 'hist': h_obs._slot.hist.clone()}
    err = None
    try:
        pas._reduce(1)
        pas._reduce(2)
    except RuntimeError as e:
        err = str(e)
    torch.save({'before': before, 'err': err, 't': t_obs._slot.minmax, 'cmin': c_obs._slot.cmins, 'cmax': c_obs._slot.cmaxs,
                'h_mm': h_obs._slot.minmax, 'hist': h_obs._slot.hist,
                'pct': torch.cat(p_obs._percentile_collector) if p_obs._percentile_collector else None}, os.path.join(out_dir, f'q{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_covers_per_channel_statistics_and_percentile_pairs(tmp_path):
    """per-channel min / max vectors must be exchanged too, in the same packed MAX all-reduce."""
    world, port = 2, _free_port()
    mp.spawn(_reduce_worker, args=(world, port, str(tmp_path), -1), nprocs=world, join=True)
    r = [torch.load(tmp_path / f'q{k}.pt') for k in range(world)]
    assert r[0]['err'] is None and r[1]['err'] is None
    for key, fn in (('cmin', torch.minimum), ('cmax', torch.maximum)):
        want = fn(r[0]['before'][key], r[1]['before'][key])
        assert torch.equal(r[0][key], want) and torch.equal(r[1][key], want)
    for key in ('t', 'h_mm'):
        want = torch.stack([torch.minimum(r[0]['before'][key][0], r[1]['before'][key][0]), torch.maximum(r[0]['before'][key][1], r[1]['before'][key][1])])
        assert torch.equal(r[0][key], want) and torch.equal(r[1][key], want)
    assert torch.equal(r[0]['hist'], r[0]['before']['hist'] + r[1]['before']['hist']) and torch.equal(r[1]['hist'], r[0]['hist'])
    # percentile pairs come back in global sample order: rank 0 batch 0, rank 1 batch 0, rank 0 batch 1, rank 1 batch 1
    assert r[0]['pct'][:, 0].tolist() == [0.0, 10.0, 1.0, 11.0] and torch.equal(r[0]['pct'], r[1]['pct'])


def test_reduce_fails_on_every_rank_when_one_rank_saw_no_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_reduce_worker, args=(world, port, str(tmp_path), 1), nprocs=world, join=True)
    r = [torch.load(tmp_path / f'q{k}.pt') for k in range(world)]
    assert all(x['err'] is not None and 'observed no calibration batch' in x['err'] for x in r)


def test_descriptor_stager_cache_is_lru_and_never_stops_caching():
    """the descriptor cache silently stopped caching after 64 entries; it is an LRU now (host-side logic, no GPU needed)."""
    sys.path.insert(0, ROOT)
    from ppq_b200.calibration import DescriptorStager
    st = DescriptorStager('cpu', 3, cache=4)
    keys = [tuple((1000 * k + i, 10 + i, i) for i in range(3)) for k in range(6)]
    first = [st.get(k) for k in keys[:4]]
    assert all(torch.equal(t, torch.tensor(k, dtype=torch.int64)) for t, k in zip(first, keys))
    assert st.get(keys[0]) is first[0]                                   # hit: the same tensor, and keys[0] becomes the most recent
    st.get(keys[4])                                                      # evicts the least recently used = keys[1]
    assert keys[1] not in st._cache and keys[0] in st._cache and len(st._cache) == 4
    st.get(keys[5])
    assert keys[2] not in st._cache and len(st._cache) == 4
    assert torch.equal(st.get(keys[1]), torch.tensor(keys[1], dtype=torch.int64))       # a miss after eviction is rebuilt and cached again
    assert keys[1] in st._cache
