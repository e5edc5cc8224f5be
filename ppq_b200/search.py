"""Host side of the scale search: the MSE grid search of TorchMSEObserver.hist_to_scale_offset
(/root/reference/ppq/quantization/observer/range.py:456-520).  Like the reference with USING_CUDA_KERNEL=True, the loop
runs on the host and every candidate's loss comes from the native compute_mse_loss (ffi.py:263-270 ->
ppq/csrc/cpu/hist_mse.cc:3-28; here ppq_b200_compute_mse_loss in ppq_b200/csrc/host.cc).  The min-max and KL searches run
on the device (csrc/search.cu) and have no host twin in this package.
"""
from decimal import ROUND_HALF_EVEN, Decimal
from math import ceil, log2

from .core import OBSERVER_MSE_COMPUTE_INTERVAL
from .ffi import CUDA


def _round_half_even(v: float) -> int:
    return int(Decimal(v).quantize(exp=Decimal(1), rounding=ROUND_HALF_EVEN))      # ppq/utils/round.py:78


def minmax_to_scale_offset_host(min_val: float, max_val: float, quant_min: int, quant_max: int, symmetrical: bool,
                                power_of_2: bool, scale_threshold: float):
    """range.py:22-75 for the two scalars the MSE search ends with (Python doubles, as upstream)."""
    if min_val > 0: min_val = 0
    if max_val < 0: max_val = 0
    if symmetrical:
        scale = max(2 * float(max(abs(max_val), abs(min_val))) / (quant_max - quant_min), scale_threshold)
        offset = 0
    else:
        scale = max(float(max_val - min_val) / (quant_max - quant_min), scale_threshold)
        offset = _round_half_even(-min_val / scale)
    if power_of_2 and scale != 0:
        scale = float(pow(2, ceil(log2(scale))))                                     # ROUND_UP on the exponent (range.py:73-74)
    return scale, offset


def mse_search_host(histogram: list, hist_scale: float, range_min: float, quant_min: int, quant_max: int, symmetrical: bool,
                    power_of_2: bool, scale_threshold: float):
    hist_bins = len(histogram)
    levels = (quant_max - quant_min) + 1
    best = None          # (loss, start, end); first minimum wins, like sorted(..)[0] on a stable sort

    def consider(start, step, end):
        nonlocal best
        loss = CUDA.compute_mse_loss(histogram=histogram, start=start, step=step, end=end)
        if best is None or loss < best[0]: best = (loss, start, end)

    step = hist_bins // levels + 1
    consider(0, step, levels * step)
    if not symmetrical:
        for start in range(0, hist_bins, OBSERVER_MSE_COMPUTE_INTERVAL):
            if (start * hist_scale) + range_min > 0: break
            for step in range(1, hist_bins // levels + 1):
                end = start + levels * step
                if end > (hist_bins + levels): break
                consider(start, step, end)
        _, s0, e0 = best
        lo, hi = (s0 * hist_scale) + range_min, (e0 * hist_scale) + range_min
    else:
        for step in range(1, hist_bins // levels + 1):
            end = levels * step
            if end > (hist_bins + levels): break
            consider(0, step, end)
        _, _, e0 = best
        lo, hi = -(e0 * hist_scale), (e0 * hist_scale)
    return minmax_to_scale_offset_host(lo, hi, quant_min, quant_max, symmetrical, power_of_2, scale_threshold)
