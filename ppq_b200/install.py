"""Install ppq_b200 into an unmodified OpenPPL/ppq checkout (drop-in boundary).

    import ppq, ppq_b200.install
    ppq_b200.install.install()            # CUDA_COMPLIER now serves the sm_100a extension; complie() is a no-op
    with ppq.api.ENABLE_CUDA_KERNEL(): ...

Everything ppq.core.ffi.CUDA calls (`CUDA_COMPLIER.CUDA_EXTENSION.<name>(...)`, /root/reference/ppq/core/ffi.py:78-344)
lands in ppq_b200/_C.so, which exports the same 20 names with the same positional signatures.
"""


_saved = {}


def install(replace_observers: bool = True):
    """Route ppq.core.ffi.CUDA to the sm_100a extension; with replace_observers also swap the 'minmax' / 'kl' observers for the
    device-resident ones (they need CUDA tensors: a CPU-path calibration of the real PPQ should call uninstall() first or pass False)."""
    import ppq.core.ffi as ref_ffi
    from .ffi import extension
    ext = extension()
    helper = ref_ffi.CUDA_COMPLIER
    if 'complie' not in _saved:
        _saved['extension'] = getattr(helper, '__CUDA_EXTENTION__', None)
        _saved['complie'] = type(helper).complie
    helper.__CUDA_EXTENTION__ = ext                      # attribute name ends with "__": no name mangling (ffi.py:19)
    type(helper).complie = lambda self: None             # ENABLE_CUDA_KERNEL.__init__ always calls complie() (api/interface.py:925-927)
    if replace_observers:
        try:
            import ppq.quantization.observer as ref_obs
            from . import observer as obs
            _saved.setdefault('observers', {k: ref_obs.OBSERVER_TABLE[k] for k in ('minmax', 'kl')})
            ref_obs.OBSERVER_TABLE['minmax'] = obs.TorchMinMaxObserver      # fused single-pass min/max
            ref_obs.OBSERVER_TABLE['kl'] = obs.TorchHistObserver           # device-resident hist_scale + on-device KL search
        except Exception:                                                   # graph-level pieces are optional
            pass
    return ext


def uninstall():
    """Undo install(): the reference's own JIT helper, extension slot and observers are back in place."""
    if not _saved: return
    import ppq.core.ffi as ref_ffi
    helper = ref_ffi.CUDA_COMPLIER
    helper.__CUDA_EXTENTION__ = _saved.pop('extension')
    type(helper).complie = _saved.pop('complie')
    obs = _saved.pop('observers', None)
    if obs:
        import ppq.quantization.observer as ref_obs
        ref_obs.OBSERVER_TABLE.update(obs)
