"""Drop-in check against the REAL reference package (only where /root/reference exists, i.e. the build container): after
ppq_b200.install.install(), ppq.core.ffi.CUDA.* reaches ppq_b200/_C.so with the reference's own argument orders, and
ENABLE_CUDA_KERNEL() no longer tries to JIT-compile anything.  No GPU here, so the calls must fail with OUR loud CPU-tensor error."""
import os
import sys

import pytest
import torch

REF = os.environ.get('PPQ_REFERENCE_ROOT', '/root/reference')


@pytest.fixture(scope='module')
def ppq():
    if not os.path.isdir(os.path.join(REF, 'ppq')):
        pytest.skip('reference package not present on this machine')
    os.environ['PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION'] = 'python'
    from unittest.mock import MagicMock
    for m in ['onnx', 'onnx.helper', 'onnx.numpy_helper', 'onnx.mapping', 'onnx.checker', 'onnx.shape_inference']:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    import ppq as _ppq
    return _ppq


def test_install_routes_every_ffi_call_to_our_extension(ppq):
    import ppq.core.ffi as ffi
    import ppq_b200.install
    from ppq.api.interface import ENABLE_CUDA_KERNEL
    from ppq.core import PPQ_CONFIG
    ext = ppq_b200.install.install()
    assert ffi.CUDA_COMPLIER.CUDA_EXTENSION is ext
    with ENABLE_CUDA_KERNEL():                       # calls complie(): must be a no-op now (the reference would JIT into its package dir)
        assert PPQ_CONFIG.USING_CUDA_KERNEL is True
        assert ffi.CUDA_COMPLIER.CUDA_EXTENSION is ext
    assert PPQ_CONFIG.USING_CUDA_KERNEL is False
    x, s, o = torch.zeros(2, 3, 4, 4), torch.ones(3), torch.zeros(3)
    calls = [
        lambda: ffi.CUDA.LinearQuantize_T(x, s[:1], o[:1], -128, 127, 0),
        lambda: ffi.CUDA.LinearQuantize_C(x, s, o, 1, -128, 127, 0),
        lambda: ffi.CUDA.LinearQuantize_T_B(x, s[:1], o[:1], x, -128, 127, 0),
        lambda: ffi.CUDA.LinearQuantize_C_B(x, s, o, x, -128, 127, 1, 0),
        lambda: ffi.CUDA.Histogram_T(x, torch.zeros(8, dtype=torch.int32), 0.1),
        lambda: ffi.CUDA.Histogram_Asymmetric_T(-1.0, 1.0, x, torch.zeros(8, dtype=torch.int32)),
        lambda: ffi.CUDA.Histogram_C(x, 1, torch.zeros(3, 8, dtype=torch.int32), 0.1),
        lambda: ffi.CUDA.Quantile(x, 0.9999),
        lambda: ffi.CUDA.TensorClip_T(x, x, s[:1]),
        lambda: ffi.CUDA.TensorClip_C(x, x, s, 1),
        lambda: ffi.CUDA.RoundingLoss_LT(x, s[:1], o[:1], -128, 127, 0),
        lambda: ffi.CUDA.RoundingLoss_LT_B(x, s[:1], s[:1], o[:1], -128, 127, 0),
        lambda: ffi.CUDA.RoundingLoss_LC(x, s, o, 1, -128, 127, 0),
        lambda: ffi.CUDA.RoundingLoss_LC_B(x, s[:1], s, o, 1, -128, 127, 0),
        lambda: ffi.CUDA.FloatingQuantize_T(x, s[:1], o[:1], 4, 3, -448.0, 448.0, 0),
        lambda: ffi.CUDA.FloatingQuantize_C(x, s, o, 1, 4, 3, -448.0, 448.0, 0),
        lambda: ffi.CUDA.FloatingQuantize_T_B(x, s[:1], o[:1], x, 4, 3, -448.0, 448.0, 0),
        lambda: ffi.CUDA.FloatingQuantize_C_B(x, s, o, x, 4, 3, -448.0, 448.0, 1, 0),
    ]
    for i, call in enumerate(calls):                  # argument counts / types bind (no TypeError), then our device check fires
        with pytest.raises(RuntimeError, match='not on a CUDA device'):
            call()
    assert isinstance(ffi.CUDA.compute_mse_loss([1, 2, 3, 4], 0, 1, 2), float)
    original = ppq_b200.install._saved['complie']
    ppq_b200.install.uninstall()
    assert type(ffi.CUDA_COMPLIER).complie is original and ffi.CUDA_COMPLIER.__CUDA_EXTENTION__ is not ext


def test_observer_table_replaced(ppq):
    import ppq.quantization.observer as ref_obs
    import ppq_b200.install
    import ppq_b200.observer as ours
    before = dict(ref_obs.OBSERVER_TABLE)
    ppq_b200.install.install(replace_observers=True)
    try:
        assert ref_obs.OBSERVER_TABLE['minmax'] is ours.TorchMinMaxObserver and ref_obs.OBSERVER_TABLE['kl'] is ours.TorchHistObserver
    finally:
        ppq_b200.install.uninstall()                      # leave the imported reference as found (other tests use its CPU path)
    assert ref_obs.OBSERVER_TABLE == before
