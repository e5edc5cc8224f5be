"""The reference's USING_CUDA_KERNEL=False quantization pipeline, end to end on the host cores -- TEST / BASELINE INFRASTRUCTURE.

Used only by bench.py (`--impl reference`) and tests/.  It restates, with torch CPU ops, what the reference's TRT_INT8 pass list does on the
CPU path for a torch.nn.Module network (the reference itself needs an ONNX graph, which cannot be loaded here: no `onnx`):
  * BatchNorm folded into the convolutions (ppq/core/common.py:41),
  * which tensors get a live config: QuantizeFusionPass (ppq/quantization/optim/refine.py:186-306) + QuantizeSimplifyPass (:64-88),
  * ParameterQuantizePass: per-channel min-max, Python loop per channel (observer/range.py:93-98, 120-135),
  * RuntimeCalibrationPass (optim/calibration.py:108-213) with the four range observers on the CPU branch:
      minmax      value.min() / value.max() appended per batch (range.py:85-100)
      kl          + torch.histc(abs(x), 4096, 0, hist_scale * 4096) accumulated (range.py:183), KL search (range.py:190-282)
      mse         the same with 2048 bins, grid search with the Python loss (range.py:431-520)
      percentile  two torch.kthvalue per batch, int() truncated indices (range.py:338-346), fp32 mean over batches (:369)
    per-channel INT8 weight fake-quant on EVERY forward, torch formulation (qfunction/linear.py:73-81),
  * QuantAlignmentPass for element-wise ops ('Align to Large', force overlap: refine.py:443-546, api/setting.py:251-256),
  * ParameterBakingPass (optim/baking.py:34-47) and the quantised forward (executor/torch.py:457-577).

PINNED: tests/test_cpu_graph_parity.py checks that this file reproduces, bit for bit, the fixture that the UNMODIFIED reference
pipeline produced on a programmatically built BaseGraph (tests/golden/graph_pipeline.npz: states, dominators, scales, offsets, the
quantised output and the baked weights, for all four observers).
"""
import time

import torch

from . import (kl_search, minmax_to_scale_offset, mse_loss_python_twin, mse_search, torch_cpu_hist_sym, torch_cpu_linear_quant_c,
               torch_cpu_linear_quant_t, torch_cpu_minmax)

KL_BINS, MSE_BINS, PERCENTILE = 4096, 2048, 0.9999
COMPUTING, ACTIVATIONS = {'Conv', 'Gemm', 'ConvTranspose', 'MatMul'}, {'Relu', 'Clip', 'Swish', 'SoftPlus', 'Sigmoid', 'Gelu'}
PASSIVE = {'MaxPool', 'GlobalMaxPool', 'Reshape', 'Flatten', 'Identity', 'DropoutSlice', 'Pad', 'Split', 'Transpose', 'Interp', 'Squeeze', 'Unsqueeze'}
ELEMENTWISE = {'Add', 'Sub', 'Sum'}


def _kind(m):
    table = {torch.nn.Conv2d: 'Conv', torch.nn.Linear: 'Gemm', torch.nn.ReLU: 'Relu', torch.nn.ReLU6: 'Clip', torch.nn.MaxPool2d: 'MaxPool',
             torch.nn.AdaptiveAvgPool2d: 'GlobalAveragePool', torch.nn.AvgPool2d: 'AveragePool', torch.nn.Flatten: 'Flatten',
             torch.nn.Sigmoid: 'Sigmoid', torch.nn.SiLU: 'Swish', torch.nn.Upsample: 'Resize'}
    for t in type(m).__mro__:
        if t in table: return table[t]
        if t.__name__ in ('Add', 'Concat') and t.__module__ != 'torch.nn.modules.module': return t.__name__      # element-wise ops written as modules
    return None


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


class Cfg:
    """A TensorQuantizationConfig reduced to what the CPU path reads: state, group root, scale / offset of the root."""

    def __init__(self, label, per_channel=False):
        self.label, self.per_channel = label, per_channel
        self.state, self.parent, self._scale, self._offset = 'INITIAL', self, None, None

    @property
    def root(self):
        if self.parent is self: return self
        self.parent = self.parent.root
        return self.parent

    def dominate(self, other):                      # self.dominated_by = other (core/quant.py:676-691)
        root, dom = self.root, other.root
        if dom is not root:
            root.parent = dom; self.parent = dom
            root.state = self.state = 'OVERLAPPED'

    def slave_of(self, master):                     # self.master_by = master (core/quant.py:702-712)
        self.parent = master
        self.state = 'PASSIVE'

    scale = property(lambda s: s.root._scale)
    offset = property(lambda s: s.root._offset)
    active = property(lambda s: s.state in ('ACTIVATED', 'PASSIVE'))


class Op:
    def __init__(self, name, module, kind, n_in):
        self.name, self.module, self.kind = name, module, kind
        self.ins = [Cfg(f'{name}|in{i}') for i in range(n_in)]
        self.out = Cfg(f'{name}|out')
        self.w = Cfg(f'{name}|w', per_channel=True) if kind in ('Conv', 'Gemm') else None
        self.sources, self.invisible, self.consumers = [None] * n_in, [False] * n_in, []


class CpuPipeline:
    def __init__(self, model, example, fuse_bn=True):
        self.model = _fuse_bn(model) if fuse_bn else model.eval()
        self.ops, self.by_call, self.calls, self.tracing = [], {}, {}, True
        self.produced, self.observers, self.phase, self.baked = {}, {}, 0, False
        names = {id(m): n for n, m in self.model.named_modules()}
        self.names = names
        for m in self.model.modules():
            if _kind(m) is not None:
                m.register_forward_pre_hook(self._pre); m.register_forward_hook(self._post)
        with torch.no_grad():
            self.calls.clear(); self.model(example)
        self.tracing, self.produced = False, {}
        self._fusion(); self._simplify()

    # ---- topology -------------------------------------------------------------------------------------------------------------
    def _op(self, m, n_in=1):
        key = (id(m), self.calls.get(id(m), 0))
        if key not in self.by_call:
            op = Op(f'{self.names[id(m)]}#{key[1]}', m, _kind(m), n_in)
            self.by_call[key] = op; self.ops.append(op)
        return self.by_call[key]

    def _fuse_act(self, producer, act):
        if len(producer.consumers) == 1 and sum(s is not None for s in act.sources) == 1:
            producer.out.dominate(act.out); act.ins[0].dominate(act.out)

    def _fusion(self):
        for op in self.ops:
            if op.kind in COMPUTING:
                for act, i in op.consumers:
                    if act.kind in ACTIVATIONS and i == 0: self._fuse_act(op, act)
        for op in self.ops:
            if op.kind in PASSIVE and op.ins and op.sources[0] is not None: op.out.dominate(op.ins[0])
        for op in self.ops:
            for act, i in op.consumers:
                if act.kind in ('Relu', 'Clip') and i == 0: self._fuse_act(op, act)
        for op in self.ops:                         # Relu fed by an in-place / functional op the hooks cannot see (`out += identity`)
            if op.kind in ('Relu', 'Clip') and op.ins and op.sources[0] is None and op.invisible[0]: op.ins[0].dominate(op.out)

    def _simplify(self):
        for src in self.ops:
            for dst, i in src.consumers:
                if dst.ins[i].state == 'INITIAL': dst.ins[i].dominate(src.out)

    # ---- hooks ----------------------------------------------------------------------------------------------------------------
    def _q_act(self, x, cfg):
        return torch_cpu_linear_quant_t(x, cfg.scale, cfg.offset, -128, 127) if cfg.active else x

    def _observe(self, cfg, x):
        ob = self.observers.get(id(cfg))
        if ob is not None: ob.observe(x, self.phase)

    def _pre(self, m, args):
        tensors = [a for a in args if isinstance(a, torch.Tensor)]
        op = self._op(m, len(tensors))
        if self.tracing:
            for i, x in enumerate(tensors):
                hit, holder = self.produced.get(id(x)), x
                if hit is None and x._base is not None: hit, holder = self.produced.get(id(x._base)), x._base
                if hit is None: continue
                if holder._version != hit[2]: op.invisible[i] = True; continue
                op.sources[i] = hit[0]; hit[0].consumers.append((op, i))
            return None
        q = []
        for x, cfg in zip(tensors, op.ins):
            q.append(self._q_act(x, cfg)); self._observe(cfg, x)
        if op.w is not None and op.w.active and not self.baked:
            m.__dict__['_fp32_w'] = m.weight.data
            m.weight.data = torch_cpu_linear_quant_c(m.weight.data, op.w.scale, op.w.offset, 0, -128, 127)
        if any(a is not b for a, b in zip(q, tensors)):
            it = iter(q)
            return tuple(next(it) if isinstance(a, torch.Tensor) else a for a in args)
        return None

    def _post(self, m, args, out):
        op = self._op(m)
        self.calls[id(m)] = self.calls.get(id(m), 0) + 1
        if self.tracing:
            self.produced[id(out)] = (op, out, out._version)
            return None
        w = m.__dict__.pop('_fp32_w', None)
        if w is not None: m.weight.data = w
        self._observe(op.out, out)
        qo = self._q_act(out, op.out)
        return qo if qo is not out else None

    @torch.no_grad()
    def forward(self, x):
        self.calls.clear()
        return self.model(x)

    # ---- passes ---------------------------------------------------------------------------------------------------------------
    def observed(self):
        """[(label, cfg)] in execution order: the configs RuntimeCalibrationPass builds observers for (state INITIAL)."""
        out = []
        for op in self.ops:
            out += [(c.label, c) for c in op.ins if c.state == 'INITIAL']
            if op.out.state == 'INITIAL': out.append((op.out.label, op.out))
        return out

    @torch.no_grad()
    def quantize_parameters(self):
        for op in self.ops:
            if op.w is None: continue
            w = op.module.weight.data
            flat = w.reshape(w.shape[0], -1)
            sc = [minmax_to_scale_offset(lo, hi, -128, 127, True)[0]                           # per-channel python loop, as upstream
                  for lo, hi in zip(flat.min(dim=1)[0].tolist(), flat.max(dim=1)[0].tolist())]
            op.w._scale, op.w._offset, op.w.state = torch.tensor(sc, dtype=torch.float32), torch.zeros(len(sc)), 'ACTIVATED'

    def observed_all(self):
        """The activation configs that take part in calibration whatever their current state (group roots): what observed() returned before."""
        out = []
        for op in self.ops:
            out += [(c.label, c) for c in op.ins if c.root is c]
            if op.out.root is op.out: out.append((op.out.label, op.out))
        return out

    @torch.no_grad()
    def calibrate(self, batches, method='kl', return_search_seconds=False):
        self.observers = {id(c): _Observer(method) for _, c in self.observed()}
        cfgs = {id(c): c for _, c in self.observed()}
        self.phase = 1
        for x in batches: self.forward(x)
        t0 = time.perf_counter()
        for k, ob in self.observers.items(): ob.render(cfgs[k], 1)
        search = time.perf_counter() - t0
        if method in ('kl', 'mse'):
            self.phase = 2
            for x in batches: self.forward(x)
            t0 = time.perf_counter()
            for k, ob in self.observers.items(): ob.render(cfgs[k], 2)
            search += time.perf_counter() - t0
        self.observers, self.phase = {}, 0
        if return_search_seconds: return search                              # the render / scale-search part (once per calibration, not per batch)
        return {c.label: c for c in cfgs.values()}

    @torch.no_grad()
    def align(self):
        for op in self.ops:
            if not op.ins: continue
            if op.kind in ELEMENTWISE:                                                         # 'Align to Large' (refine.py:443-482)
                lo = hi = 0
                for c in op.ins:
                    hi = max(hi, (c.scale * (127 - c.offset)).item()); lo = min(lo, (c.scale * (-128 - c.offset)).item())
                s, o = minmax_to_scale_offset(lo, hi, -128, 127, True)
                master = op.ins[0]
                master.parent, master.state = master, 'PASSIVE'
                master._scale, master._offset = torch.tensor(s, dtype=torch.float32), torch.tensor(float(o), dtype=torch.float32)
                for c in op.ins[1:]: c.slave_of(master)
            elif op.kind in ('Concat', 'Resize'):                                              # 'Align to Output' (refine.py:484-496)
                master = op.out
                for c in op.ins: c.slave_of(master)
            else:
                continue
            for src in op.sources:
                if src is not None: src.out.slave_of(master)                                   # force_alignment_overlap = True

    @torch.no_grad()
    def bake(self):
        for op in self.ops:
            if op.w is not None and op.w.active:
                op.module.weight.data = torch_cpu_linear_quant_c(op.module.weight.data, op.w.scale, op.w.offset, 0, -128, 127)
                op.w.state = 'BAKED'
        self.baked = True


class _Observer:
    """The four range observers of observer/range.py on their CPU branch."""

    def __init__(self, method):
        self.method, self.mins, self.maxs, self.pct, self.hist, self.hist_scale = method, [], [], [], None, None

    def observe(self, x, phase):
        if self.method == 'percentile':
            flat, n = x.flatten(), x.numel()
            lo_i, hi_i = max(0, int(n * (1 - PERCENTILE))) + 1, min(int(n * PERCENTILE), n - 1) + 1
            lo, hi = torch.kthvalue(flat, k=lo_i, dim=0)[0].view(1, -1), torch.kthvalue(flat, k=hi_i, dim=0)[0].view(1, -1)
            self.pct.append(torch.cat([hi, lo], dim=-1))
        elif phase == 1:
            lo, hi = torch_cpu_minmax(x)
            self.mins.append(lo.reshape(1)); self.maxs.append(hi.reshape(1))
        else:
            bins = KL_BINS if self.method == 'kl' else MSE_BINS
            h = torch_cpu_hist_sym(x, self.hist_scale, bins)
            self.hist = h if self.hist is None else self.hist + h

    def render(self, cfg, phase):
        def put(s, o):
            cfg._scale, cfg._offset, cfg.state = torch.tensor(s, dtype=torch.float32), torch.tensor(float(o), dtype=torch.float32), 'ACTIVATED'
        if self.method == 'percentile':
            m = torch.cat(self.pct, dim=0).float().mean(dim=0)
            return put(*minmax_to_scale_offset(m[1].item(), m[0].item(), -128, 127, True))
        if phase == 1:
            self.lo, self.hi = torch.min(torch.cat(self.mins)).item(), torch.max(torch.cat(self.maxs)).item()
            if self.method == 'minmax': return put(*minmax_to_scale_offset(self.lo, self.hi, -128, 127, True))
            self.hist_scale = float(max(abs(self.hi), abs(self.lo))) / (KL_BINS if self.method == 'kl' else MSE_BINS)
        elif self.method == 'kl':
            put(*kl_search(self.hist, self.hist_scale, 8))
        else:
            put(*mse_search(self.hist, self.hist_scale, self.lo, -128, 127, True, loss_fn=mse_loss_python_twin))
