// Our shim (not reference code): exposes the reference's C++ compute_mse_loss
// (ppq/csrc/cpu/hist_mse.cc:3-28, compiled in place from /root/reference) through a C ABI so that
// the oracle restatement can be pinned against it with ctypes.  Test infrastructure only.
#include <cstdint>
#include "hist_mse.h"
extern "C" float ref_compute_mse_loss(const int64_t* hist, int64_t n, int start, int step, int end) {
    std::vector<int64_t> h(hist, hist + n);
    return compute_mse_loss(h, start, step, end);
}
