"""Executor: the slice of TorchExecutor that the hot path lives in (/root/reference/ppq/executor/torch.py:457-577, 610-652)
for a torch.nn.Module network -- no graph IR, no ONNX (neither is part of the path).

What is mirrored, with the reference's semantics:
  * every quantable operation (a Conv2d / Linear / ReLU / pooling call site) owns TensorQuantizationConfigs for its inputs,
    parameters and outputs with the TensorRT-INT8 policy of ppq/quantization/quantizer/TensorRTQuantizer.py:12-105:
    per-tensor symmetric INT8 activations, per-channel symmetric INT8 weights on axis 0 ('minmax'), bias FP32;
  * forward(inputs, hooks): for each operation call quantize_function on every input and parameter (weights are re-quantised
    on EVERY forward until baked: torch.py:516-518), run pre-forward hooks with (fp32, quantised) values, run the op, quantise
    the outputs, run post-forward hooks (torch.py:526-553).  Only ACTIVATED / PASSIVE configs quantise (core/quant.py:357-359);
  * Conv -> ReLU fusion marks the conv output OVERLAPPED (QuantizeFusionPass) so it is neither observed nor
    quantised; downstream inputs are dominated by the producer's output config;
  * dummy_forward-style ParameterQuantizePass (optim/parameters.py:172-215): per-channel min/max + scale search for weights.
Conv / Gemm execution itself stays in torch (cuDNN / cuBLAS), as in the reference.

B200-native addition: `collect=True` makes the executor hand the observed fp32 tensors of a whole forward to an ArenaCalibrator
(one multi-tensor launch per phase) instead of firing one observer launch per tensor.
"""
from typing import Dict, List, Optional

import torch

from .core import LinearQuantizationConfig, QuantizationStates, TensorQuantizationConfig
from .qfunction import PPQuantFunction

_KINDS = {torch.nn.Conv2d: 'Conv', torch.nn.Linear: 'Gemm', torch.nn.ReLU: 'Relu', torch.nn.ReLU6: 'Clip',
          torch.nn.MaxPool2d: 'MaxPool', torch.nn.AdaptiveAvgPool2d: 'GlobalAveragePool', torch.nn.AvgPool2d: 'AveragePool'}


def fuse_conv_bn(model: torch.nn.Module) -> torch.nn.Module:
    """PPQ folds BatchNorm into the preceding convolution when it loads a graph (core/common.py:41 FORMATTER_FUSE_BN)."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    model.eval()

    def walk(mod):
        prev_name, prev = None, None
        for name, child in list(mod.named_children()):
            if isinstance(child, torch.nn.BatchNorm2d) and isinstance(prev, torch.nn.Conv2d):
                setattr(mod, prev_name, fuse_conv_bn_eval(prev, child))
                setattr(mod, name, torch.nn.Identity())
                prev_name, prev = None, None
                continue
            walk(child)
            prev_name, prev = name, child
    walk(model)
    return model


class QuantableOperation:
    """One call site of a quantable module.  `input_configs()` / `output_configs()` are what OperationObserver walks
    (observer/__init__.py:92-113)."""

    def __init__(self, name: str, module: torch.nn.Module, kind: str, is_graph_input: bool):
        self.name, self.module, self.kind = name, module, kind
        act = dict(symmetrical=True, quant_min=-128, quant_max=127, calibration='percentile')
        self.input_cfg: Optional[TensorQuantizationConfig] = LinearQuantizationConfig(**act) if is_graph_input else None
        self.weight_cfg: Optional[TensorQuantizationConfig] = None
        if kind in ('Conv', 'Gemm'):
            self.weight_cfg = LinearQuantizationConfig(symmetrical=True, quant_min=-128, quant_max=127, channel_axis=0, calibration='minmax')
        self.output_cfg: TensorQuantizationConfig = LinearQuantizationConfig(**act)

    def input_configs(self):
        if self.input_cfg is not None: yield (self.name + ':in', self.input_cfg, False)
        if self.weight_cfg is not None: yield (self.name + ':weight', self.weight_cfg, True)

    def output_configs(self):
        yield (self.name + ':out', self.output_cfg)


class TorchExecutor:
    def __init__(self, model: torch.nn.Module, example_input: torch.Tensor, fuse_bn: bool = True):
        self.model = fuse_conv_bn(model) if fuse_bn else model.eval()
        self.operations: Dict[str, QuantableOperation] = {}
        self._order: List[str] = []
        self._calls: Dict[int, int] = {}
        self._names = {id(m): n for n, m in self.model.named_modules()}
        self._hooks, self._collect, self._collected, self._sink = None, False, None, None
        self._tracing, self._last_out, self._wq = True, None, {}
        self._quant_fn = PPQuantFunction
        for m in self.model.modules():
            if type(m) in _KINDS:
                m.register_forward_pre_hook(self._pre)
                m.register_forward_hook(self._post)
        with torch.no_grad():
            self._begin()
            self.model(example_input)
        self._tracing = False

    # -- graph view used by the calibration pass
    def quantable_operations(self):
        return [(n, self.operations[n]) for n in self._order]

    def quantize_function(self, tensor: torch.Tensor, config: Optional[TensorQuantizationConfig]) -> torch.Tensor:
        """torch.py:610-613."""
        if config is None or not QuantizationStates.is_activated(config.state): return tensor
        return self._quant_fn(tensor, config)

    def _begin(self):
        self._calls.clear()
        self._collected = []
        self._wq = self._quantize_all_weights() if not self._tracing else {}

    def _quantize_all_weights(self):
        """All activated per-channel linear weight configs that share (axis, range, rounding) are fake-quantised by ONE multi-tensor launch
        at the start of the forward (the per-module results are what the pre-forward hooks hand to the operations).  Falls back to one
        launch per weight for anything else (FP8 weights, mixed policies)."""
        from .calibration import MultiWeightQuantizer
        from .core import QuantizationProperty as P
        ops = [self.operations[n] for n in self._order]
        todo = [op for op in ops if op.weight_cfg is not None and QuantizationStates.is_activated(op.weight_cfg.state)
                and op.weight_cfg.policy.has_property(P.LINEAR) and op.weight_cfg.policy.has_property(P.PER_CHANNEL)
                and not op.weight_cfg.policy.has_property(P.DYNAMIC) and self._quant_fn is PPQuantFunction]
        if len(todo) < 2: return {}
        c0 = todo[0].weight_cfg
        key0 = (c0.channel_axis, c0.quant_min, c0.quant_max, c0.rounding)
        todo = [op for op in todo if (op.weight_cfg.channel_axis, op.weight_cfg.quant_min, op.weight_cfg.quant_max, op.weight_cfg.rounding) == key0]
        sig = tuple((id(op.module), op.module.weight.data_ptr(), op.weight_cfg.scale.data_ptr(), op.weight_cfg.offset.data_ptr()) for op in todo)
        if getattr(self, '_mw_sig', None) != sig:
            dev = todo[0].module.weight.device
            self._mw = MultiWeightQuantizer([op.module.weight.data for op in todo], [op.weight_cfg.scale.to(dev) for op in todo],
                                            [op.weight_cfg.offset.to(dev) for op in todo], channel_axis=c0.channel_axis,
                                            quant_min=c0.quant_min, quant_max=c0.quant_max,
                                            rounding=c0.rounding.value if hasattr(c0.rounding, 'value') else int(c0.rounding))
            # the quantizer keeps its own references to scale/offset: recompute the signature from what it actually points at
            self._mw_sig = sig
            self._mw_ids = [id(op.module) for op in todo]
        outs = self._mw()
        return dict(zip(self._mw_ids, outs))

    def _op_of(self, module) -> QuantableOperation:
        idx = self._calls.get(id(module), 0)
        name = f'{self._names[id(module)]}#{idx}'
        if self._tracing and name not in self.operations:
            self.operations[name] = QuantableOperation(name, module, _KINDS[type(module)], is_graph_input=(len(self._order) == 0))
            self._order.append(name)
        return self.operations[name]

    def _pre(self, module, args):
        op = self._op_of(module)
        x = args[0]
        if self._tracing:
            # conv -> relu fusion: the activation takes over the producer's output config (QuantizeFusionPass)
            if (op.kind in ('Relu', 'Clip') and self._last_out is not None and self._last_out[1] is x and x._version == self._last_out[2]
                    and self._last_out[0].kind in ('Conv', 'Gemm')):      # same tensor object AND not modified in place since (residual +=)
                self._last_out[0].output_cfg.state = QuantizationStates.OVERLAPPED
            return None
        hook = self._hooks.get(op.name) if self._hooks else None
        inputs, qinputs, cfgs = [], [], []
        if op.input_cfg is not None:
            inputs.append(x); qinputs.append(self.quantize_function(x, op.input_cfg)); cfgs.append(op.input_cfg)
            if self._collect and op.input_cfg.state == QuantizationStates.INITIAL: self._emit(x)
        if op.weight_cfg is not None:
            w = module.weight
            wq = self._wq.get(id(module))                                # re-quantised every forward (torch.py:516-518): multi-tensor launch
            if wq is None: wq = self.quantize_function(w.data, op.weight_cfg)
            inputs.append(w.data); qinputs.append(wq); cfgs.append(op.weight_cfg)
            if wq is not w.data:
                module.__dict__['_ppq_fp32_weight'] = w.data
                w.data = wq
        if hook is not None: hook.pre_forward_hook(inputs=inputs, quant_inputs=qinputs, quant_configs=cfgs)
        if op.input_cfg is not None and qinputs[0] is not x:
            return (qinputs[0],) + tuple(args[1:])
        return None

    def _post(self, module, args, output):
        op = self._op_of(module)
        self._calls[id(module)] = self._calls.get(id(module), 0) + 1
        if self._tracing:
            self._last_out = (op, output, output._version)
            return None
        fp32 = module.__dict__.pop('_ppq_fp32_weight', None)
        if fp32 is not None: module.weight.data = fp32
        qout = self.quantize_function(output, op.output_cfg)
        hook = self._hooks.get(op.name) if self._hooks else None
        if hook is not None: hook.post_forward_hook(outputs=[output], quant_outputs=[qout], quant_configs=[op.output_cfg])
        if self._collect and op.output_cfg.state == QuantizationStates.INITIAL: self._emit(output)
        return qout if qout is not output else None

    def _emit(self, tensor: torch.Tensor):
        # sink(k, tensor) observes the k-th observed tensor of this forward immediately; unless it returns False the tensor is consumed
        consumed = self._sink is not None and self._sink(len(self._collected), tensor) is not False
        self._collected.append(None if consumed else tensor)

    @torch.no_grad()
    def forward(self, inputs: torch.Tensor, hooks: Optional[dict] = None, collect: bool = False, sink=None):
        """torch.py:365-410.  `collect=True` returns (output, observed fp32 tensors) for a deferred multi-tensor launch -- only
        valid for tensors the network does not overwrite in place; `sink(k, tensor)` observes the k-th tensor immediately (a sink that
        returns False leaves the tensor in the collected list instead)."""
        self._hooks, self._collect, self._sink = hooks, collect or sink is not None, sink
        self._begin()
        out = self.model(inputs)
        collected, self._collected, self._sink = self._collected, None, None
        return (out, collected) if collect else out

    def observed_configs(self) -> List[TensorQuantizationConfig]:
        """Configs in the order `collect=True` returns their tensors."""
        cfgs = []
        for n in self._order:
            op = self.operations[n]
            if op.input_cfg is not None and op.input_cfg.state == QuantizationStates.INITIAL: cfgs.append(op.input_cfg)
            if op.output_cfg.state == QuantizationStates.INITIAL: cfgs.append(op.output_cfg)
        return cfgs

    def observed_configs_all(self) -> List[TensorQuantizationConfig]:
        """All activation configs that take part in calibration, whatever their current state (same order as observed_configs())."""
        cfgs = []
        for n in self._order:
            op = self.operations[n]
            if op.input_cfg is not None: cfgs.append(op.input_cfg)
            if op.output_cfg.state != QuantizationStates.OVERLAPPED: cfgs.append(op.output_cfg)
        return cfgs

    @torch.no_grad()
    def quantize_parameters(self):
        """ParameterQuantizePass (optim/parameters.py:172-215): per-channel min/max observers on the weights, rendered to
        scale/offset by the on-device search."""
        from .observer import TensorObserverFactroy
        for n in self._order:
            op = self.operations[n]
            if op.weight_cfg is not None and op.weight_cfg.state == QuantizationStates.INITIAL:
                ob = TensorObserverFactroy.build_observer(n + ':weight', op.weight_cfg)
                ob.observe(op.module.weight.data)
                ob.render_quantization_config()

    @torch.no_grad()
    def bake_parameters(self):
        """ParameterBakingPass (optim/baking.py:34-47, IR/quantize.py:98-111): quantise each weight once and freeze it."""
        for n in self._order:
            op = self.operations[n]
            if op.weight_cfg is not None and op.weight_cfg.state == QuantizationStates.ACTIVATED:
                op.module.weight.data = self._quant_fn(op.module.weight.data, op.weight_cfg)
                op.weight_cfg.state = QuantizationStates.BAKED


# ------------------------------------------------------------------------------------------------------------------ calibration drivers
def prefetch_to_device(batches, to_device, device):
    """Yield `to_device(batch)` for every batch with the copy of batch k+1 enqueued on a side stream before batch k is handed out, so
    the host->device transfer of the next calibration batch (19 MB for 32 images) overlaps the forward of the current one."""
    copy_stream = torch.cuda.Stream(device=device)
    it = iter(batches)

    def fetch():
        x = next(it, None)
        if x is None: return None
        with torch.cuda.stream(copy_stream):
            y = to_device(x)
            done = torch.cuda.Event(); done.record(copy_stream)
        return y, done

    ahead = fetch()
    while ahead is not None:
        y, done = ahead
        main = torch.cuda.current_stream(device)
        main.wait_event(done)
        if isinstance(y, torch.Tensor): y.record_stream(main)            # allocated on the copy stream, consumed on the main stream
        ahead = fetch()
        yield y


@torch.no_grad()
def calibrate_arena(executor: TorchExecutor, batches, method: str = 'kl', group=None, to_device=None, deferred='auto',
                    graphs: bool = False, prefetch: bool = True):
    """Two-phase calibration of every observed activation through one ArenaCalibrator (statistics arena, one all-reduce per
    phase, on-device scale search).  `batches` is this rank's share of the calibration set (sample-sharded by the caller).
    deferred=False observes each tensor as the forward produces it (one launch per tensor);
    deferred=True keeps all tensors alive and issues ONE multi-tensor launch per forward (only valid when nothing is overwritten in place);
    deferred='auto' (default) finds out during the first forward which observed tensors the network later overwrites in place
    (torchvision adds residuals with `out += identity`), observes those immediately and everything else in one multi-tensor launch;
    graphs=True captures one forward per phase -- network kernels, per-forward weight fake-quant and the collectors -- into a CUDA
    graph and replays it for every batch (fixed batch shape).  Measured on B200 (ResNet-50, 8 x 32 images): capture + instantiate
    costs more than it saves at 8-16 batches per phase (1330 vs 3124 imgs/s end to end), so it is off by default; it pays off for long
    calibration sets or when the same graph is reused across calls."""
    from .calibration import ArenaCalibrator
    cfgs = executor.observed_configs()
    for c in cfgs: c.observer_algorithm = method
    if iter(batches) is batches: batches = list(batches)                  # a one-shot iterator would leave phase 2 without data
    dev = next(executor.model.parameters()).device
    cal = ArenaCalibrator(len(cfgs), dev, method=method, group=group)
    static_in = None
    mutated = None                                                        # slots whose tensors are overwritten later in the forward
    while True:
        graph = None
        overlapped = prefetch and to_device is not None and not graphs and dev.type == 'cuda'
        for x in (prefetch_to_device(batches, to_device, dev) if overlapped else batches):
            if graphs:
                if static_in is None:
                    static_in = torch.empty(x.shape, dtype=torch.float32, device=dev)
                static_in.copy_(x, non_blocking=True)
                if graph is None:
                    # eager run on a side stream first (cuDNN autotuning, lazy initialisation), statistics restored afterwards:
                    # min/max are idempotent under re-observation, histogram counts are not
                    keep_mm, keep_h = cal.minmax.clone(), cal.hist.clone()
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        executor.forward(static_in, sink=cal.observe_one)
                    torch.cuda.current_stream(dev).wait_stream(side)
                    cal.minmax.copy_(keep_mm); cal.hist.copy_(keep_h)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        executor.forward(static_in, sink=cal.observe_one)
                    cal.minmax.copy_(keep_mm); cal.hist.copy_(keep_h)      # capture does not execute, but stay safe
                graph.replay()
                continue
            if to_device is not None and not overlapped: x = to_device(x)
            if deferred is True:
                _, tensors = executor.forward(x, collect=True)
                cal.observe([t if t.is_contiguous() else t.contiguous() for t in tensors])
                del tensors
            elif deferred == 'auto':
                if mutated is None:                                       # probe forward: observe immediately, remember versions
                    seen = []

                    def probe(k, t):
                        cal.observe_one(k, t); seen.append((t, t._version))
                    executor.forward(x, sink=probe)
                    mutated = {k for k, (t, v) in enumerate(seen) if t._version != v or not t.is_contiguous()}
                    del seen
                else:
                    def sink(k, t):
                        if k in mutated:
                            cal.observe_one(k, t); return True
                        return False
                    _, tensors = executor.forward(x, collect=True, sink=sink)
                    rest = [(k, t) for k, t in enumerate(tensors) if t is not None]
                    cal.observe([t for _, t in rest], [k for k, _ in rest])
                    del tensors, rest
            else:
                executor.forward(x, sink=cal.observe_one)
        if cal.end_phase(): break
    for i, c in enumerate(cfgs):
        c.scale, c.offset, c.state = cal.scale[i], cal.offset[i], QuantizationStates.ACTIVATED
    return cal


@torch.no_grad()
def graphwise_error_analyse(executor: TorchExecutor, batches, to_device=None) -> Dict[str, float]:
    """The evaluation loop where fake-quant throughput shows up in user wall-clock (ppq/quantization/analyse/graphwise.py:64-183): run every batch
    through the network twice -- all configs dequantised (fp32) and all configs active -- and report, per quantable operation, the SNR
    mean((q - f)^2) / mean(f^2) per sample, averaged (torch_snr_error, ppq/quantization/measure/norm.py:52-93).  Every activation goes through
    QuantizeTensor_LT and every weight through the multi-tensor QuantizeTensor_LC on each quantised forward."""
    from .core import set_state
    cfgs = executor.observed_configs_all() + [op.weight_cfg for _, op in executor.quantable_operations() if op.weight_cfg is not None]
    saved = [c.state for c in cfgs]
    names = [n for n, op in executor.quantable_operations()]
    acc = {n: [0.0, 0] for n in names}

    class Tap:
        def __init__(self, name, store): self.name, self.store = name, store
        def pre_forward_hook(self, **kw): pass
        def post_forward_hook(self, outputs, quant_outputs, quant_configs): self.store[self.name] = quant_outputs[0]

    for x in batches:
        if to_device is not None: x = to_device(x)
        fp, qt = {}, {}
        for c in cfgs: c.state = type(c.state)['FP32'] if getattr(c.state, 'name', '') in ('ACTIVATED', 'PASSIVE') else c.state
        executor.forward(x, hooks={n: Tap(n, fp) for n in names})
        for c, st in zip(cfgs, saved): c.state = st
        executor.forward(x, hooks={n: Tap(n, qt) for n in names})
        for n in names:
            f, q = fp[n].flatten(1), qt[n].flatten(1)
            snr = (torch.pow(q - f, 2).sum(dim=-1) / (torch.pow(f, 2).sum(dim=-1) + 1e-7)).mean()
            acc[n][0] += float(snr); acc[n][1] += 1
    for c, st in zip(cfgs, saved): c.state = st
    return {n: v[0] / max(v[1], 1) for n, v in acc.items()}


def e2e_calibration_benchmark(batch: int, steps: int, warmup: int, device, world: int = 1, seed: int = 0, graphs: bool = False):
    """bench.py's `e2e`: ResNet-50 (random init, BN folded) calibrated end to end through the public API -- images in pinned host
    memory, H2D copy of every batch inside the timed region (both phases), torch forward with per-forward weight fake-quant,
    multi-tensor collectors, the two all-reduces, on-device KL search and a D2H read of the resulting scales."""
    import torch.distributed as dist
    import torchvision
    torch.manual_seed(seed)
    torch.backends.cudnn.benchmark = True
    model = torchvision.models.resnet50(weights=None).eval()
    ex = TorchExecutor(model.to(device), torch.zeros(2, 3, 224, 224, device=device))
    ex.quantize_parameters()
    g = torch.Generator().manual_seed(seed + 1)
    host = [torch.rand(batch, 3, 224, 224, generator=g).pin_memory() for _ in range(steps)]
    stream = torch.cuda.current_stream()

    act_cfgs = ex.observed_configs()

    def reset():
        for c in act_cfgs: c.state = QuantizationStates.INITIAL

    def run(batches):
        cal = calibrate_arena(ex, batches, method='kl', to_device=lambda x: x.to(device, non_blocking=True), graphs=graphs)
        return cal.scale.cpu()                                            # D2H of the result (synchronises)

    for _ in range(max(warmup, 1)):
        run(host[:2]); reset()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    scales = run(host)
    t1.record(stream)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=device); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
    assert bool(torch.isfinite(scales).all()) and bool((scales > 0).all())
    return {'value': round(world * steps * batch / (ms * 1e-3), 1), 'unit': 'imgs/s',
            'h2d_bytes_per_step': 2 * batch * 3 * 224 * 224 * 4, 'd2h_bytes_per_step': int(scales.numel() * 4 / steps) + 1,
            'ms_per_step': round(ms / steps, 3), 'steps': steps, 'observed_tensors': int(scales.numel()), 'cuda_graphs': graphs,
            'what': 'pinned-host images -> H2D -> torch ResNet-50 forward (fp32, cuDNN) with per-forward INT8 per-channel weight fake-quant '
                    '-> multi-tensor min/max (phase 1) / histogram (phase 2) -> all-reduce -> on-device KL search -> scales D2H'}
