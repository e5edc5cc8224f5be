"""BASELINE.json configs 3-5 as parity cases (config 1 is in test_gpu_parity.py, config 2 is bench.py's workload), plus the host-side
pieces of the path that sit above the kernels: FP8 observers, dynamic quantisation, the module executor and the calibration drivers."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ext():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from ppq_b200.ffi import extension
    return extension()


def bits(t):
    return t.detach().cpu().numpy().view(np.uint32)


def test_config3_mobilenetv2_per_channel_weights(ext, oracle):
    """Config 3: per-channel INT8 ChannelwiseLinearQuant over the Conv/Linear weights of MobileNetV2 (53 tensors, 3.47 M elements, 17 depth-wise
    layers with epc = 9): scales from the on-device per-channel min-max search, fake-quant by single launches AND by the one-launch multi-tensor
    kernel, both bit-exact vs the oracle."""
    import torchvision
    from ppq_b200 import LinearQuantizationConfig
    from ppq_b200.calibration import MultiWeightQuantizer
    from ppq_b200.observer import Observer
    from ppq_b200.qfunction import PPQLinearQuant_toInt, PPQuantFunction
    torch.manual_seed(0)
    m = torchvision.models.mobilenet_v2(weights=None)
    ws = [p.weight.data.cuda() for p in m.modules() if isinstance(p, (torch.nn.Conv2d, torch.nn.Linear))]
    assert len(ws) == 53 and sum(w.numel() for w in ws) == 3469760 and sum(1 for w in ws if w.shape[1:] == (1, 3, 3)) == 17
    cfgs = []
    for w in ws:
        cfg = LinearQuantizationConfig(symmetrical=True, quant_min=-128, quant_max=127, channel_axis=0, calibration='minmax')
        ob = Observer(cfg); ob.observe(w); ob.render_quantization_config()
        lo, hi = oracle.minmax_c(w.cpu().numpy(), 0)
        want = np.float32([oracle.minmax_to_scale_offset(float(a), float(b), -128, 127, True)[0] for a, b in zip(lo, hi)])
        assert np.array_equal(cfg.scale.cpu().numpy(), want), tuple(w.shape)
        y = PPQuantFunction(w, cfg)
        yo, qo = oracle.linear_quant_c(w.cpu().numpy(), want, np.zeros_like(want), 0, -128, 127, 0, return_int=True)
        assert np.array_equal(bits(y), yo.view(np.uint32)), tuple(w.shape)
        q = PPQLinearQuant_toInt(w, cfg)
        assert q.dtype == torch.int8 and np.array_equal(q.cpu().numpy(), qo.astype(np.int8)), tuple(w.shape)
        cfgs.append(cfg)
    outs = MultiWeightQuantizer(ws, [c.scale for c in cfgs], [c.offset for c in cfgs])()
    for w, c, y in zip(ws, cfgs, outs):
        assert torch.equal(y, PPQuantFunction(w, c)), tuple(w.shape)


def test_config4_fp8_bert_activations(ext, oracle):
    """Config 4: FP8 E4M3 FloatingQuant on BERT-base activation shapes (bs 32, seq 512): hidden / FFN / attention probabilities, the observer's
    power-of-two scale candidates, tie-rich (bf16-representable) values; oracle on a strided sample, full-size properties on the rest."""
    from ppq_b200 import FloatingQuantizationConfig, QuantizationStates
    from ppq_b200.observer import Observer
    from ppq_b200.qfunction import PPQuantFunction
    g = torch.Generator(device='cuda').manual_seed(4)
    for shape, sigma in (((32, 512, 768), 1.0), ((32, 512, 3072), 8.0), ((32, 12, 512, 512), 1.0)):
        x = torch.randn(shape, device='cuda', generator=g) * sigma
        x.view(-1)[::5] = x.view(-1)[::5].bfloat16().float()                   # exact-tie candidates
        for s in (1.0, 0.125, 4.0):
            cfg = FloatingQuantizationConfig(exponent=4, mantissa=3, quant_min=-448.0, quant_max=448.0, power_of_2=True)
            cfg.scale, cfg.offset, cfg.state = torch.tensor(s, device='cuda'), torch.tensor(0.0, device='cuda'), QuantizationStates.ACTIVATED
            y = PPQuantFunction(x, cfg)
            assert torch.equal(PPQuantFunction(y, cfg), y)                     # idempotent
            assert float((y / s).abs().max()) <= 448.0
            idx = torch.arange(0, x.numel(), 997, device='cuda')
            want = oracle.float_quant_t(x.view(-1)[idx].cpu().numpy(), s, 0.0)
            assert np.array_equal(bits(y.view(-1)[idx]), want.view(np.uint32)), (shape, s)
    # the FP8 observers: constant (scale 1) and direct-MSE (7 power-of-two candidates)
    x = torch.randn(32, 512, 768, device='cuda', generator=g) * 50
    cfg = FloatingQuantizationConfig(calibration='constant')
    ob = Observer(cfg); ob.observe(x); ob.render_quantization_config()
    assert cfg.state == QuantizationStates.ACTIVATED and cfg.scale.item() == 1.0 and cfg.offset.item() == 0.0
    cfg = FloatingQuantizationConfig(calibration='floating')
    torch.manual_seed(1)
    ob = Observer(cfg); ob.observe(x); ob.render_quantization_config()
    assert cfg.scale.item() in ob.SCALE_CANDIDATES
    # the chosen candidate minimises the fake-quant MSE over the whole tensor as well (sampled search, stable for this distribution)
    losses = []
    for s in ob.SCALE_CANDIDATES:
        c2 = FloatingQuantizationConfig(); c2.scale, c2.offset, c2.state = torch.tensor(s, device='cuda'), torch.tensor(0.0, device='cuda'), QuantizationStates.ACTIVATED
        losses.append(float(((PPQuantFunction(x, c2) - x) ** 2).mean()))
    assert ob.SCALE_CANDIDATES[int(np.argmin(losses))] == cfg.scale.item()


def test_config5_sharded_calibration_equals_single_run(ext):
    """Config 5's shape of work (YOLOv5s-like activation set at 640x640, samples sharded over 8 ranks): emulate the 8 ranks one after the
    other on one GPU, merge their arenas exactly as the two all-reduces do, and compare with the unsharded calibration bit for bit."""
    from ppq_b200.calibration import ArenaCalibrator, pack_minmax_for_max_reduce, shard_indices, unpack_minmax_after_max_reduce
    shapes = [(3, 640, 640), (32, 320, 320), (64, 160, 160), (128, 80, 80), (256, 40, 40), (512, 20, 20), (255, 80, 80), (255, 40, 40), (255, 20, 20)]
    g = torch.Generator(device='cuda').manual_seed(5)
    samples = [[(torch.randn((1,) + s, device='cuda', generator=g) * (1 + i % 3)).relu_() if k % 2 else torch.randn((1,) + s, device='cuda', generator=g)
                for k, s in enumerate(shapes)] for i in range(16)]
    T = len(shapes)
    full = ArenaCalibrator(T, 'cuda')
    for smp in samples: full.observe(smp)
    full.end_phase()
    for smp in samples: full.observe(smp)
    full.end_phase()
    R = 8
    ranks = [ArenaCalibrator(T, 'cuda') for _ in range(R)]
    for r, cal in enumerate(ranks):
        for i in shard_indices(len(samples), r, R): cal.observe(samples[i])
    packed = torch.stack([pack_minmax_for_max_reduce(c.minmax) for c in ranks]).amax(dim=0)       # == all_reduce(MAX)
    for cal in ranks:
        unpack_minmax_after_max_reduce(packed, cal.minmax); cal.end_phase()
    assert all(torch.equal(c.minmax, full.minmax) and torch.equal(c.hist_scale, full.hist_scale) for c in ranks)
    for r, cal in enumerate(ranks):
        for i in shard_indices(len(samples), r, R): cal.observe(samples[i])
    total = torch.stack([c.hist for c in ranks]).sum(dim=0, dtype=torch.int32)                   # == all_reduce(SUM)
    assert torch.equal(total, full.hist)
    for cal in ranks:
        cal.hist.copy_(total); cal.end_phase()
        assert torch.equal(cal.scale, full.scale) and torch.equal(cal.best_bin_range, full.best_bin_range)


def test_dynamic_quantization_on_device(ext, oracle):
    from ppq_b200 import LinearQuantizationConfig, QuantizationStates
    from ppq_b200.qfunction import PPQuantFunction
    g = torch.Generator(device='cuda').manual_seed(6)
    x = torch.randn(8, 24, 56, 56, device='cuda', generator=g) * 2
    cfg = LinearQuantizationConfig(symmetrical=False, dynamic=True, quant_min=0, quant_max=255)
    cfg.state = QuantizationStates.ACTIVATED
    y = PPQuantFunction(x, cfg)
    lo, hi = oracle.minmax_t(x.cpu().numpy())
    s, o = oracle.minmax_to_scale_offset(float(lo), float(hi), 0, 255, False)
    assert np.array_equal(bits(y), oracle.linear_quant_t(x.cpu().numpy(), np.float32(s), np.float32(o), 0, 255).view(np.uint32))
    cfg = LinearQuantizationConfig(symmetrical=True, dynamic=True, channel_axis=1)
    cfg.state = QuantizationStates.ACTIVATED
    y = PPQuantFunction(x, cfg)
    lo, hi = oracle.minmax_c(x.cpu().numpy(), 1)
    sc = np.float32([oracle.minmax_to_scale_offset(float(a), float(b), -128, 127, True)[0] for a, b in zip(lo, hi)])
    assert np.array_equal(bits(y), oracle.linear_quant_c(x.cpu().numpy(), sc, np.zeros_like(sc), 1, -128, 127).view(np.uint32))


def test_executor_calibration_paths_agree(ext):
    """The module executor: hook-driven RuntimeCalibrationPass (one observer per tensor, the reference flow) and the arena calibrator (immediate
    and deferred multi-tensor) give identical scales; weights are re-quantised per forward.  (Parity of the same flows with the REAL reference
    pipeline, incl. alignment and baked weights: tests/test_gpu_graph_parity.py.)"""
    import torchvision
    from ppq_b200.calibration import RuntimeCalibrationPass
    from ppq_b200.core import QuantizationStates
    from ppq_b200.executor import TorchExecutor, calibrate_arena
    torch.manual_seed(7)
    data = [torch.rand(4, 3, 64, 64, device='cuda') for _ in range(8)]

    def build():
        torch.manual_seed(11)
        ex = TorchExecutor(torchvision.models.resnet18(weights=None).cuda(), torch.zeros(2, 3, 64, 64, device='cuda'))
        ex.quantize_parameters()
        return ex
    ex1 = build()
    RuntimeCalibrationPass(method='kl').optimize(graph=ex1, dataloader=data, executor=ex1, calib_steps=8)
    s1 = torch.stack([c.scale for c in ex1.observed_configs_all()])
    ex2 = build()
    cal = calibrate_arena(ex2, data, method='kl', deferred=False)
    assert torch.equal(cal.scale, s1)
    ex4 = build()
    cal4 = calibrate_arena(ex4, data, method='kl')                         # deferred='auto': multi-tensor launch for everything not overwritten in place
    assert torch.equal(cal4.minmax, cal.minmax) and torch.equal(cal4.hist, cal.hist) and torch.equal(cal4.scale, s1)
    ex5 = build()                                                           # host batches, H2D of batch k+1 overlapped with forward k
    host = [d.cpu().pin_memory() for d in data]
    cal5 = calibrate_arena(ex5, host, method='kl', to_device=lambda t: t.to('cuda', non_blocking=True), prefetch=True)
    assert torch.equal(cal5.minmax, cal.minmax) and torch.equal(cal5.hist, cal.hist) and torch.equal(cal5.scale, s1)
    ex6 = build()                                                           # the same through the persistent two-slot device ring (bench.py's e2e path),
    cal6 = calibrate_arena(ex6, host, method='kl', to_device='ring')        # ... whose buffers are overwritten two batches later: nothing may read them late
    assert torch.equal(cal6.minmax, cal.minmax) and torch.equal(cal6.hist, cal.hist) and torch.equal(cal6.scale, s1)
    cal7 = calibrate_arena(build(), host, method='kl', to_device='ring', deferred=False)
    assert torch.equal(cal7.hist, cal.hist)
    assert all(c.state == QuantizationStates.ACTIVATED for c in ex2.observed_configs_all())
    # evaluation loop (graphwise error analysis): per-op SNR of the quantised network vs fp32, all below the reference's 0.1 bar
    from ppq_b200.executor import graphwise_error_analyse
    report = graphwise_error_analyse(ex2, data[:2])
    assert len(report) == len(ex2.quantable_operations()) and all(0 <= v < 0.1 for v in report.values()), max(report.values())
    assert all(c.state == QuantizationStates.ACTIVATED for c in ex2.observed_configs_all())       # states restored
    if os.environ.get('PYTORCH_NO_CUDA_MEMORY_CACHING') != '1':
        report_g = graphwise_error_analyse(ex2, data[:3], graphs=True)     # both forwards captured in CUDA graphs, replayed per batch
        report_e = graphwise_error_analyse(ex2, data[:3])
        assert report_g.keys() == report_e.keys() and all(abs(report_g[k] - report_e[k]) <= 1e-6 + 1e-4 * abs(report_e[k]) for k in report_e), \
            max(abs(report_g[k] - report_e[k]) for k in report_e)
        assert all(c.state == QuantizationStates.ACTIVATED for c in ex2.observed_configs_all())
    # CUDA-graph replay of the whole forward (network + weight fake-quant + collectors): identical statistics
    if os.environ.get('PYTORCH_NO_CUDA_MEMORY_CACHING') != '1':            # capture needs torch's caching allocator (tools/gpu_sanitize.sh turns it off)
        ex3 = build()
        cal3 = calibrate_arena(ex3, data, method='kl', graphs=True)
        assert torch.equal(cal3.minmax, cal.minmax) and torch.equal(cal3.hist, cal.hist) and torch.equal(cal3.scale, s1)
    # quantised forward runs and differs from fp32 only by quantisation noise
    x = data[0]
    yq = ex2.forward(x)
    for c in ex2.observed_configs_all(): c.state = QuantizationStates.FP32
    for _, op in ex2.quantable_operations():
        if op.weight_cfg is not None: op.weight_cfg.state = QuantizationStates.FP32
    yf = ex2.forward(x)
    snr = float(((yq - yf) ** 2).sum() / (yf ** 2).sum())
    assert 0 < snr < 0.1                                                    # tests/test_block.py:35-40 bar: SNR(quantised vs fp32) < 0.1


def test_channels_last_network_is_read_in_storage_order(ext, oracle):
    """A network run in NHWC (cuDNN's native layout): per-tensor collectors / fake-quant are order-independent and read the dense tensors as they lie;
    axis-0 weight rows stay contiguous in channels_last, so the multi-tensor weight fake-quant needs no copy either.  Same statistics as NCHW up to
    the convolutions' own rounding; element-wise results identical."""
    import torchvision
    from ppq_b200.calibration import MultiWeightQuantizer
    from ppq_b200.executor import TorchExecutor, calibrate_arena
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(8, 24, 14, 14, device='cuda', generator=g)
    xc = x.contiguous(memory_format=torch.channels_last)
    s, o = torch.tensor([0.03], device='cuda'), torch.tensor([0.0], device='cuda')
    y, yc = ext.QuantizeTensor_LT(x, s, o, -128, 127, 0), ext.QuantizeTensor_LT(xc, s, o, -128, 127, 0)
    assert yc.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, yc)          # same values at the same logical positions, no layout change
    assert torch.equal(ext.Quantile_T(x, 0.99), ext.Quantile_T(xc, 0.99))
    mm, mmc = torch.empty(2, device='cuda'), torch.empty(2, device='cuda')
    ext.MinMax_Init(mm[0:1], mm[1:2]); ext.MinMax_Init(mmc[0:1], mmc[1:2]); ext.MinMax_T(x, mm); ext.MinMax_T(xc, mmc)
    assert torch.equal(mm, mmc)
    w = torch.randn(32, 16, 3, 3, device='cuda', generator=g) * 0.1
    wc = w.contiguous(memory_format=torch.channels_last)
    sc = (w.abs().amax(dim=(1, 2, 3)) / 127).contiguous(); oc = torch.zeros_like(sc)
    out = MultiWeightQuantizer([wc, w], [sc, sc], [oc, oc], channel_axis=0)()
    assert out[0].is_contiguous(memory_format=torch.channels_last) and torch.equal(out[0], out[1])
    assert np.array_equal(bits(out[1]), oracle.linear_quant_c(w.cpu().numpy(), sc.cpu().numpy(), oc.cpu().numpy(), 0, -128, 127).view(np.uint32))
    torch.manual_seed(7)
    data = [torch.rand(4, 3, 64, 64, device='cuda') for _ in range(8)]

    def scales(channels_last):
        torch.manual_seed(11)
        ex = TorchExecutor(torchvision.models.resnet18(weights=None).cuda(), torch.zeros(2, 3, 64, 64, device='cuda'), channels_last=channels_last)
        ex.quantize_parameters()
        return calibrate_arena(ex, data, method='minmax').scale, len(ex.observed_configs_all())
    (a, na), (b, nb) = scales(False), scales(True)
    assert na == nb and torch.allclose(a, b, rtol=2e-3, atol=1e-7), (a - b).abs().max()


def test_config5_yolov5s_network_through_the_executor(ext):
    """BASELINE config 5's network itself (bench_models.YOLOv5s: Conv-SiLU fusion, shortcut Adds, Concats, Upsamples, SPPF max-pools) at a small
    resolution: the hook-driven pass and the arena calibrator (one multi-tensor launch per forward) give identical scales for all 83 observed tensors,
    alignment makes every Add / Concat / Upsample input share its master's scale, and the quantised forward runs (noise bounded)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench_models
    from ppq_b200.calibration import RuntimeCalibrationPass
    from ppq_b200.core import QuantizationStates as S
    from ppq_b200.executor import TorchExecutor, calibrate_arena, graphwise_error_analyse
    torch.manual_seed(5)
    data = [torch.rand(2, 3, 160, 160, device='cuda') for _ in range(8)]

    def build():
        torch.manual_seed(13)
        ex = TorchExecutor(bench_models.YOLOv5s().cuda(), torch.zeros(1, 3, 160, 160, device='cuda'))
        ex.quantize_parameters()
        return ex
    ex1 = build()
    assert len(ex1.observed_configs()) == 83
    RuntimeCalibrationPass(method='kl').optimize(graph=ex1, dataloader=data, executor=ex1, calib_steps=8)
    s1 = torch.stack([c.scale for c in ex1.observed_configs_all()])
    ex2 = build()
    cal = calibrate_arena(ex2, data, method='kl')
    assert torch.equal(cal.scale, s1)
    ex3 = build()
    cal3 = calibrate_arena(ex3, data, method='percentile')                 # the reference's default observer, one select table per forward
    ex4 = build()
    RuntimeCalibrationPass(method='percentile').optimize(graph=ex4, dataloader=data, executor=ex4, calib_steps=8)
    assert torch.equal(cal3.scale, torch.stack([c.scale for c in ex4.observed_configs_all()]))
    ex2.align_quantization()
    for name, op in ex2.quantable_operations():
        if op.kind in ('Add', 'Concat', 'Resize'):
            master = op.input_cfgs[0].dominated_by
            assert all(c.state == S.PASSIVE or c is master for c in op.input_cfgs) and all(c.scale is master.scale for c in op.input_cfgs), name
    report = graphwise_error_analyse(ex2, data[:2])                       # a random-init 25-layer-deep network: a sanity bound, not the 0.1 bar of trained ones
    assert len(report) == len(ex2.quantable_operations()) and all(0 <= v < 0.5 for v in report.values()), max(report.values())


@pytest.mark.gpu
def test_e2e_benchmark_arm_runs_and_reports_its_breakdown(ext):
    """bench.py's `e2e` arm at toy size (host batches -> H2D -> hooked forward -> collectors -> search -> D2H): the function the driver's headline comes
    from must run, count its bytes from the tensors it copies and attribute the slowest calibration to host or device (per-batch trace)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench_models
    from ppq_b200.executor import e2e_calibration_benchmark
    r = e2e_calibration_benchmark(batch=2, batches=3, steps=2, warmup=2, device=torch.device('cuda:0'), channels_last=True,
                                  model=bench_models.YOLOv5s(), image=(3, 96, 96), distinct_host_batches=2)
    assert r['value'] > 0 and r['observed_tensors'] == 83 and r['steps'] == 2
    assert r['h2d_bytes_per_step'] == 2 * 3 * 2 * 3 * 96 * 96 * 4 and r['d2h_bytes_per_step'] == 83 * 4
    slow = r['slowest_step']
    assert slow['ms'] == r['step_ms']['max'] and slow['host_enqueue_gap_ms']['max'] >= slow['host_enqueue_gap_ms']['median'] > 0
    assert slow['device_gap_ms']['at'].startswith('phase ')
    assert set(r['breakdown_ms_per_batch_pass']) == {'forward_fp32_cudnn', 'hooks_and_weight_fakequant', 'collectors_exchange_search', 'h2d_copy_overlapped', 'total'}
