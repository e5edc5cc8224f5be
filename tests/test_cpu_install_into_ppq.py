"""Drop-in check against the REAL reference package (only where /root/reference exists, i.e. the build container): after
ppq_b200.install.install(), ppq.core.ffi.CUDA.* reaches ppq_b200/_C.so with the reference's own argument orders, and
ENABLE_CUDA_KERNEL() no longer tries to JIT-compile anything.  No GPU here, so the calls must fail with OUR loud CPU-tensor error."""
import pytest
import torch

import ppq_b200

import refppq


@pytest.fixture(scope='module')
def ppq():
    mod = refppq.load()
    if mod is None: pytest.skip('reference package not present on this machine (neither /root/reference nor baseline/_ref)')
    return mod


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
    import ppq.quantization.optim.calibration as ref_cal
    import ppq.quantization.optim.ssd as ref_ssd
    import ppq_b200.install
    import ppq_b200.observer as ours
    before = dict(ref_obs.OBSERVER_TABLE)
    ref_hist, ref_mse = ref_cal.TorchHistObserver, ref_cal.TorchMSEObserver
    ppq_b200.install.install(replace_observers=True)
    try:
        for k in ('minmax', 'kl', 'mse', 'percentile'):
            assert ref_obs.OBSERVER_TABLE[k] is ours.OBSERVER_TABLE[k]
        # the reference keeps an observer for calibration phase 2 only if `type(observer) in {TorchHistObserver, TorchMSEObserver}` as
        # imported by optim/calibration.py:10-13 (:195) and optim/ssd.py:14 (:445): the replaced classes must satisfy that test
        cfg = ppq_b200.LinearQuantizationConfig(calibration='kl')
        assert type(ref_obs.OBSERVER_TABLE['kl'](watch_on=None, quant_cfg=cfg)) in {ref_cal.TorchHistObserver, ref_cal.TorchMSEObserver}
        assert type(ref_obs.OBSERVER_TABLE['mse'](watch_on=None, quant_cfg=cfg)) in {ref_cal.TorchHistObserver, ref_cal.TorchMSEObserver}
        assert type(ref_obs.OBSERVER_TABLE['kl'](watch_on=None, quant_cfg=cfg)) in {ref_ssd.TorchHistObserver}
        assert type(ref_obs.OBSERVER_TABLE['minmax'](watch_on=None, quant_cfg=cfg)) not in {ref_cal.TorchHistObserver, ref_cal.TorchMSEObserver}
    finally:
        ppq_b200.install.uninstall()                      # leave the imported reference as found (other tests use its CPU path)
    assert ref_obs.OBSERVER_TABLE == before
    assert ref_cal.TorchHistObserver is ref_hist and ref_cal.TorchMSEObserver is ref_mse and ref_ssd.TorchHistObserver is ref_hist
