"""A neutral description of small test networks + deterministic parameters, and the two builders that turn it into
  * the REFERENCE's BaseGraph (programmatic construction: /root/reference/ppq/IR/base/graph.py:696-830), and
  * a plain torch.nn.Module for ppq_b200.executor.TorchExecutor,
so that the same weights and the same calibration data go through the unmodified reference pipeline and through ours.
Test infrastructure: imported by tests/ and tests/golden/make_graph_golden.py only.

`run_reference_pipeline` drives the reference's own quantizer pass list (BaseQuantizer.build_quant_pipeline,
ppq/quantization/quantizer/base.py:249-350) one pass at a time and snapshots every TensorQuantizationConfig after
RuntimeCalibrationPass (the hot path's result) and after the last pass (alignment / passive parameters / baking).
"""
import numpy as np
import torch

# TinyRes: conv-relu-maxpool, a depth-wise conv (epc 9), a residual Add (alignment), global pooling, flatten, gemm.
TINY_RES = [
    dict(op='Conv', name='conv1', inputs=['x'], out='conv1.o', cout=16, cin=3, k=3, pad=1, stride=1, group=1),
    dict(op='Relu', name='relu1', inputs=['conv1.o'], out='relu1.o'),
    dict(op='MaxPool', name='pool1', inputs=['relu1.o'], out='pool1.o', k=2, stride=2),
    dict(op='Conv', name='conv2', inputs=['pool1.o'], out='conv2.o', cout=16, cin=1, k=3, pad=1, stride=1, group=16),
    dict(op='Relu', name='relu2', inputs=['conv2.o'], out='relu2.o'),
    dict(op='Conv', name='conv3', inputs=['relu2.o'], out='conv3.o', cout=16, cin=16, k=1, pad=0, stride=1, group=1),
    dict(op='Add', name='add1', inputs=['conv3.o', 'pool1.o'], out='add1.o'),
    dict(op='Relu', name='relu3', inputs=['add1.o'], out='relu3.o'),
    dict(op='Conv', name='conv4', inputs=['relu3.o'], out='conv4.o', cout=24, cin=16, k=3, pad=1, stride=2, group=1),
    dict(op='GlobalAveragePool', name='gap', inputs=['conv4.o'], out='gap.o', k=8),
    dict(op='Flatten', name='flat', inputs=['gap.o'], out='flat.o'),
    dict(op='Gemm', name='fc', inputs=['flat.o'], out='y', cout=10, cin=24),
]
# TinyCat: two branches joined by a Concat ('Align to Output' by default, api/setting.py:254), a conv -> Sigmoid activation fusion, an AveragePool
# (quantable, not passive, alignment 'None'), a max-pool whose producer has two consumers.
TINY_CAT = [
    dict(op='Conv', name='stem', inputs=['x'], out='stem.o', cout=8, cin=3, k=3, pad=1, stride=2, group=1),
    dict(op='Relu', name='act0', inputs=['stem.o'], out='act0.o'),
    dict(op='Conv', name='branch_a', inputs=['act0.o'], out='branch_a.o', cout=8, cin=8, k=3, pad=1, stride=1, group=1),
    dict(op='Sigmoid', name='sig_a', inputs=['branch_a.o'], out='sig_a.o'),
    dict(op='MaxPool', name='pool_b', inputs=['act0.o'], out='pool_b.o', k=3, stride=1, pad=1),
    dict(op='Concat', name='cat', inputs=['sig_a.o', 'pool_b.o'], out='cat.o'),
    dict(op='Conv', name='mix', inputs=['cat.o'], out='mix.o', cout=12, cin=16, k=1, pad=0, stride=1, group=1),
    dict(op='Relu', name='act1', inputs=['mix.o'], out='act1.o'),
    dict(op='AveragePool', name='avg', inputs=['act1.o'], out='avg.o', k=2, stride=2),
    dict(op='Conv', name='head', inputs=['avg.o'], out='head.o', cout=6, cin=12, k=1, pad=0, stride=1, group=1),
]
SPECS = {'tinyres': TINY_RES, 'tinycat': TINY_CAT}
INPUT_SHAPE = {'tinyres': (3, 32, 32), 'tinycat': (3, 32, 32)}


def make_params(spec, seed):
    """name -> float32 numpy array (numpy legacy RandomState: bit-stable)."""
    r = np.random.RandomState(seed)
    p = {}
    for o in spec:
        if o['op'] == 'Conv':
            fan = o['cin'] * o['k'] * o['k']
            p[o['name'] + '.w'] = (r.standard_normal((o['cout'], o['cin'], o['k'], o['k'])) * (1.6 / np.sqrt(fan))).astype(np.float32)
            p[o['name'] + '.b'] = (r.standard_normal(o['cout']) * 0.1).astype(np.float32)
        elif o['op'] == 'Gemm':
            p[o['name'] + '.w'] = (r.standard_normal((o['cout'], o['cin'])) * (1.0 / np.sqrt(o['cin']))).astype(np.float32)
            p[o['name'] + '.b'] = (r.standard_normal(o['cout']) * 0.1).astype(np.float32)
    return p


def make_data(name, seed, steps, batch):
    r = np.random.RandomState(seed)
    return [r.rand(batch, *INPUT_SHAPE[name]).astype(np.float32) for _ in range(steps)]       # torch.rand-style images in [0, 1)


# ------------------------------------------------------------------------------------------------ torch.nn.Module (our executor)
class RefGemm(torch.nn.Linear):
    """The reference's Gemm forward, operation for operation (executor/op/torch/default.py:2055-2063): a matmul, then the bias add."""

    def forward(self, x):
        return 1.0 * torch.matmul(x, self.weight.transpose(0, 1)) + 1.0 * self.bias


class SpecNet(torch.nn.Module):
    """Executes the spec with one sub-module per operation (names = the spec's names), so that module hooks see every op."""

    def __init__(self, spec, params):
        super().__init__()
        from ppq_b200.executor import Add as QAdd
        self.spec = spec
        for o in spec:
            if o['op'] == 'Conv':
                m = torch.nn.Conv2d(o['cin'] * o['group'], o['cout'], o['k'], stride=o['stride'], padding=o['pad'], groups=o['group'])
            elif o['op'] == 'Gemm':
                m = RefGemm(o['cin'], o['cout'])
            elif o['op'] == 'Relu':
                m = torch.nn.ReLU()
            elif o['op'] == 'MaxPool':
                m = torch.nn.MaxPool2d(o['k'], o['stride'], padding=o.get('pad', 0))
            elif o['op'] == 'AveragePool':
                m = torch.nn.AvgPool2d(o['k'], o['stride'])
            elif o['op'] == 'Sigmoid':
                m = torch.nn.Sigmoid()
            elif o['op'] == 'Concat':
                from ppq_b200.executor import Concat as QConcat
                m = QConcat(dim=1)
            elif o['op'] == 'GlobalAveragePool':
                m = torch.nn.AvgPool2d(o['k'])             # the reference runs F.avg_pool2d(x, kernel_size=x.size()[2:]) (executor/op/torch/default.py:770)
            elif o['op'] == 'Flatten':
                m = torch.nn.Flatten(1)
            elif o['op'] == 'Add':
                m = QAdd()
            else:
                raise NotImplementedError(o['op'])
            if o['op'] in ('Conv', 'Gemm'):
                with torch.no_grad():
                    m.weight.copy_(torch.from_numpy(params[o['name'] + '.w']))
                    m.bias.copy_(torch.from_numpy(params[o['name'] + '.b']))
            self.add_module(o['name'], m)

    def forward(self, x):
        v = {'x': x}
        for o in self.spec:
            v[o['out']] = getattr(self, o['name'])(*[v[i] for i in o['inputs']])
        return v[self.spec[-1]['out']]


# ------------------------------------------------------------------------------------------------ the reference's BaseGraph
def build_ppq_graph(ppq, spec, params):
    from ppq import BaseGraph, NetworkFramework
    g = BaseGraph(name='spec', built_from=NetworkFramework.ONNX)
    var = {}

    def v(name, value=None):
        if name not in var:
            var[name] = g.create_variable(name=name, value=None if value is None else torch.from_numpy(value.copy()), is_parameter=value is not None)
        return var[name]

    for o in spec:
        ins = [v(i) for i in o['inputs']]
        attrs = {}
        if o['op'] == 'Conv':
            ins += [v(o['name'] + '.w', params[o['name'] + '.w']), v(o['name'] + '.b', params[o['name'] + '.b'])]
            attrs = {'kernel_shape': [o['k']] * 2, 'pads': [o['pad']] * 4, 'strides': [o['stride']] * 2, 'dilations': [1, 1], 'group': o['group']}
        elif o['op'] == 'Gemm':
            ins += [v(o['name'] + '.w', params[o['name'] + '.w']), v(o['name'] + '.b', params[o['name'] + '.b'])]
            attrs = {'alpha': 1.0, 'beta': 1.0, 'transB': 1}
        elif o['op'] == 'MaxPool':
            attrs = {'kernel_shape': [o['k']] * 2, 'strides': [o['stride']] * 2, 'pads': [o.get('pad', 0)] * 4}
        elif o['op'] == 'AveragePool':
            attrs = {'kernel_shape': [o['k']] * 2, 'strides': [o['stride']] * 2, 'pads': [0, 0, 0, 0]}
        elif o['op'] == 'Concat':
            attrs = {'axis': 1}
        elif o['op'] == 'Flatten':
            attrs = {'axis': 1}
        g.create_operation(o['op'], name=o['name'], inputs=ins, outputs=[v(o['out'])], attributes=attrs)
    g.mark_variable_as_graph_input(var['x'])
    g.mark_variable_as_graph_output(var[spec[-1]['out']])
    return g


def snapshot(graph):
    """[(op, variable, state, observer_algorithm, dominator 'op|variable' or None, scale, offset)] for every TQC of every quantable op."""
    from ppq import QuantableOperation
    owner = {}
    for name, op in graph.operations.items():
        if isinstance(op, QuantableOperation):
            for cfg, var in op.config_with_variable:
                owner[cfg._hash] = f'{name}|{var.name}'
    rows = []
    for name, op in graph.operations.items():
        if not isinstance(op, QuantableOperation): continue
        for cfg, var in op.config_with_variable:
            dom = cfg.dominated_by
            s, o = cfg.scale, cfg.offset
            rows.append(dict(op=name, var=var.name, state=cfg.state.name, algo=cfg.observer_algorithm,
                             dominator=None if dom is cfg else owner.get(dom._hash),
                             scale=None if s is None else s.detach().cpu().float().flatten().numpy().copy(),
                             offset=None if o is None else o.detach().cpu().float().flatten().numpy().copy()))
    return rows


def run_reference_pipeline(ppq, spec, params, data, method, device='cpu', cuda_kernel=False, analyse=False):
    """The unmodified reference: TRT_INT8 quantizer -> its own pass list -> snapshots + the quantised graph's output on data[0].
    With cuda_kernel=True the passes run inside `with ENABLE_CUDA_KERNEL():` (whatever extension ppq.core.ffi currently serves)."""
    import contextlib

    import ppq.lib as PFL
    from ppq import TargetPlatform, TorchExecutor
    from ppq.api.interface import ENABLE_CUDA_KERNEL
    from ppq.api.setting import QuantizationSettingFactory
    graph = build_ppq_graph(ppq, spec, params)
    batches = [torch.from_numpy(x).to(device) for x in data]
    qz = PFL.Quantizer(platform=TargetPlatform.TRT_INT8, graph=graph)
    table = PFL.Dispatcher(graph=graph).dispatch(quant_types=qz.quant_operation_types)
    for op in list(graph.operations.values()):
        qz.quantize_operation(op_name=op.name, platform=table[op.name])
    ex = TorchExecutor(graph=graph, device=device)
    ex.tracing_operation_meta(inputs=batches[0])
    ex.load_graph(graph=graph)
    setting = QuantizationSettingFactory.default_setting()
    setting.quantize_activation_setting.calib_algorithm = method
    passes = list(qz.build_quant_pipeline(setting))
    names = [type(p).__name__ for p in passes]
    assert names == ['QuantizeFusionPass', 'QuantizeSimplifyPass', 'ParameterQuantizePass', 'RuntimeCalibrationPass',
                     'QuantAlignmentPass', 'PassiveParameterQuantizePass', 'ParameterBakingPass'], names
    res = {'passes': names}
    with (ENABLE_CUDA_KERNEL() if cuda_kernel else contextlib.nullcontext()):
        for p in passes:
            p.optimize(graph=graph, dataloader=batches, executor=ex, verbose=False, calib_steps=len(batches), collate_fn=None)
            if type(p).__name__ == 'RuntimeCalibrationPass':
                res['calibrated'] = snapshot(graph)
        res['final'] = snapshot(graph)
        res['output'] = ex.forward(batches[0])[0].detach().cpu().numpy().copy()
        res['baked'] = {k: graph.variables[k].value.detach().cpu().numpy().copy() for k in params if k.endswith('.w')}
        if analyse:                                                      # the reference's own evaluation loop (quantization/analyse/graphwise.py:64-183)
            from ppq.quantization.analyse import graphwise_error_analyse
            res['graphwise'] = graphwise_error_analyse(graph=graph, running_device=device, dataloader=batches, collate_fn=None, method='snr',
                                                       steps=len(batches), verbose=False)
    return res
