"""Install ppq_b200 into an unmodified OpenPPL/ppq checkout (drop-in boundary).

    import ppq, ppq_b200.install
    ppq_b200.install.install()            # CUDA_COMPLIER now serves the sm_100a extension; complie() is a no-op
    with ppq.api.ENABLE_CUDA_KERNEL(): ...

Everything ppq.core.ffi.CUDA calls (`CUDA_COMPLIER.CUDA_EXTENSION.<name>(...)`, /root/reference/ppq/core/ffi.py:78-344)
lands in ppq_b200/_C.so, which exports the same 20 names with the same positional signatures.
"""


def install(replace_observers: bool = True):
    import ppq.core.ffi as ref_ffi
    from .ffi import extension
    ext = extension()
    helper = ref_ffi.CUDA_COMPLIER
    helper.__CUDA_EXTENTION__ = ext                      # attribute name ends with "__": no name mangling (ffi.py:19)
    type(helper).complie = lambda self: None             # ENABLE_CUDA_KERNEL.__init__ always calls complie() (api/interface.py:925-927)
    if replace_observers:
        try:
            import ppq.quantization.observer as ref_obs
            from . import observer as obs
            ref_obs.OBSERVER_TABLE['minmax'] = obs.TorchMinMaxObserver      # fused single-pass min/max
            ref_obs.OBSERVER_TABLE['kl'] = obs.TorchHistObserver           # device-resident hist_scale + on-device KL search
        except Exception:                                                   # graph-level pieces are optional
            pass
    return ext
