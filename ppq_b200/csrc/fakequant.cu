// fakequant.cu -- sm_100a fake-quant kernels (quantize -> clamp -> dequantize) behind the C ABI of include/ppq_b200.h.
//
//   ppq_b200_linear_quant_t / _c   replace QuantizeTensor_LT / _LC  (/root/reference/ppq/csrc/cuda/linear.cu:38-233)
//   ppq_b200_float_quant_t  / _c   replace QuantizeTensor_FT / _FC  (ppq/csrc/cuda/floating.cu:36-131, common.cuh:154-226)
//   ppq_b200_linear_quant_*_toint  device twin of PPQLinearQuant_toInt (ppq/quantization/qfunction/linear.py:218-238)
//
// Design: pure streaming kernels, 8 B/element of algorithmic HBM traffic (4 read + 4 write),
// no reuse -> no shared-memory staging in the default variant.  Each thread keeps 4 independent 128-bit loads in flight
// (64 B), 256 threads per CTA, persistent grid of 148 SMs x 8 CTAs, grid-stride over float4 vectors, so ~128 KB are in
// flight per SM -- enough to cover HBM3e latency x bandwidth.  The arithmetic per element is kept under ~15 issue slots
// (hoisted exact reciprocal + 2 Markstein corrections instead of div.rn, see common.cuh) so the kernel stays
// bandwidth-bound rather than issue-bound.  Variant 1 (ppq_b200_set_variant("linear_quant_t", 1)) is the
// TMA-staged pipeline (cp.async.bulk global->shared ring with mbarriers, bulk store back) kept for comparison.
#include "ops.cuh"
#include "../../include/ppq_b200.h"
#include "variants.h"

namespace ppqb {

// ---- output writers ------------------------------------------------------------------------------------------------
template <class Op, class OutT> struct Emit;
template <class Op> struct Emit<Op, float> {
    static __device__ __forceinline__ float one(const Op &op, float x) { return op.apply(x); }
    static __device__ __forceinline__ void vec(const Op &op, const float4 &v, float *y, int64_t vi) {
        reinterpret_cast<float4 *>(y)[vi] = op.apply4(v);
    }
};
template <class Op> struct Emit<Op, int32_t> {
    static __device__ __forceinline__ int32_t one(const Op &op, float x) { return op.quant(x); }
    static __device__ __forceinline__ void vec(const Op &op, const float4 &v, int32_t *y, int64_t vi) {
        reinterpret_cast<int4 *>(y)[vi] = op.quant4(v);
    }
};
template <class Op> struct Emit<Op, int8_t> {       // also used for uint8 (same low byte)
    static __device__ __forceinline__ int8_t one(const Op &op, float x) { return (int8_t)op.quant(x); }
    static __device__ __forceinline__ void vec(const Op &op, const float4 &v, int8_t *y, int64_t vi) {
        const int4 q = op.quant4(v);
        reinterpret_cast<uint32_t *>(y)[vi] = ((uint32_t)q.x & 0xFFu) | (((uint32_t)q.y & 0xFFu) << 8) |
                                              (((uint32_t)q.z & 0xFFu) << 16) | (((uint32_t)q.w & 0xFFu) << 24);
    }
};

constexpr int kThreads = 256;
constexpr int kUnroll = 4;          // independent 128-bit loads in flight per thread

// ---- per-tensor: one (scale, offset) for the whole tensor ---------------------------------------------------------------
// VEC: x and y are 16-byte aligned -> float4 main loop + scalar tail; otherwise everything scalar.
template <class Op, class OutT, bool VEC>
__global__ void __launch_bounds__(kThreads)
ew_tensor_kernel(const float *__restrict__ x, OutT *__restrict__ y, int64_t n,
                 const float *__restrict__ scale, const float *__restrict__ offset, typename Op::Params p) {
    const Op op(p, __ldg(scale), __ldg(offset));
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    if constexpr (VEC) {
        const int64_t n4 = n >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        int64_t i = tid;
        for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
            float4 v[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; j++) v[j] = ld_stream4(x4 + i + j * stride);
#pragma unroll
            for (int j = 0; j < kUnroll; j++) Emit<Op, OutT>::vec(op, v[j], y, i + j * stride);
        }
        for (; i < n4; i += stride) Emit<Op, OutT>::vec(op, ld_stream4(x4 + i), y, i);
        const int64_t t = (n4 << 2) + tid;                          // <= 3 leftover elements
        if (t < n) y[t] = Emit<Op, OutT>::one(op, x[t]);
    } else {
        for (int64_t i = tid; i < n; i += stride) y[i] = Emit<Op, OutT>::one(op, x[i]);
    }
}

// ---- per-channel, vectorised: epc % 4 == 0 and 16-byte aligned bases, so a float4 never straddles a channel row ----------
// Flat grid-stride over float4 vectors exactly like the per-tensor kernel (same memory-level parallelism); the channel of
// vector vi is (vi / (epc/4)) % C, from two multiply-high divisions, and the operator (exact reciprocal, integer offset)
// is rebuilt per vector: ~9 extra issue slots per element, still under the HBM-bound budget.
template <class Op, class OutT>
__global__ void __launch_bounds__(kThreads)
ew_channel_vec_kernel(const float *__restrict__ x, OutT *__restrict__ y, int64_t n4, int C,
                      FastDiv div_epc4, FastDiv div_C,
                      const float *__restrict__ scale, const float *__restrict__ offset, typename Op::Params p) {
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    auto emit = [&](const float4 &v, int64_t vi) {
        const int64_t row = (int64_t)div_epc4.quot((uint64_t)vi);
        const int c = (int)(row - (int64_t)div_C.quot((uint64_t)row) * C);
        const Op op(p, __ldg(scale + c), __ldg(offset + c));
        Emit<Op, OutT>::vec(op, v, y, vi);
    };
    for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
        float4 v[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) v[j] = ld_stream4(x4 + i + j * stride);
#pragma unroll
        for (int j = 0; j < kUnroll; j++) emit(v[j], i + j * stride);
    }
    for (; i < n4; i += stride) emit(ld_stream4(x4 + i), i);
}

// ---- per-channel, generic: any epc (1, 9, 27, ...), any alignment ---------------------------------------------------------
// Each thread owns 4 consecutive elements; (row, col) of the first comes from one fast division, the rest by walking.
template <class Op, class OutT>
__global__ void __launch_bounds__(kThreads)
ew_channel_generic_kernel(const float *__restrict__ x, OutT *__restrict__ y, int64_t n, int64_t epc, int C,
                          FastDiv div_epc, FastDiv div_C,
                          const float *__restrict__ scale, const float *__restrict__ offset, typename Op::Params p) {
    const int64_t groups = (n + 3) >> 2;
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kThreads) {
        const int64_t e0 = g << 2;
        int64_t row = (int64_t)div_epc.quot((uint64_t)e0);
        int64_t col = e0 - row * epc;
        int c = (int)(row - (int64_t)div_C.quot((uint64_t)row) * C);
        const int cnt = (int)((n - e0) < 4 ? (n - e0) : 4);
        if (col + cnt <= epc) {                                      // all in one row: one operator for the group
            const Op op(p, __ldg(scale + c), __ldg(offset + c));
            for (int j = 0; j < cnt; j++) y[e0 + j] = Emit<Op, OutT>::one(op, ld_stream1(x + e0 + j));
        } else {
            for (int j = 0; j < cnt; j++) {
                const Op op(p, __ldg(scale + c), __ldg(offset + c));
                y[e0 + j] = Emit<Op, OutT>::one(op, ld_stream1(x + e0 + j));
                if (++col == epc) { col = 0; if (++c == C) c = 0; }
            }
        }
    }
}

// ---- host-side launch helpers --------------------------------------------------------------------------------------------
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
template <class OutT> static inline bool out_aligned(const void *p) {
    return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(OutT) - 1)) == 0;
}

template <class Op, class OutT>
static int launch_tensor(const float *x, OutT *y, int64_t n, const float *scale, const float *offset,
                         typename Op::Params p, cudaStream_t st) {
    if (n <= 0 || !x || !y || !scale || !offset) return (int)cudaErrorInvalidValue;
    const bool vec = aligned16(x) && out_aligned<OutT>(y);
    const int grid = grid_for(vec ? (n + 3) / 4 : n, kThreads, vec ? kUnroll : 4, 8);
    if (vec) ew_tensor_kernel<Op, OutT, true><<<grid, kThreads, 0, st>>>(x, y, n, scale, offset, p);
    else     ew_tensor_kernel<Op, OutT, false><<<grid, kThreads, 0, st>>>(x, y, n, scale, offset, p);
    return (int)cudaGetLastError();
}

template <class Op, class OutT>
static int launch_channel(const float *x, OutT *y, int64_t n, int64_t epc, int C, const float *scale, const float *offset,
                          typename Op::Params p, cudaStream_t st) {
    if (n <= 0 || epc <= 0 || C <= 0 || !x || !y || !scale || !offset) return (int)cudaErrorInvalidValue;
    if (epc > 0x7fffffffLL || n % epc != 0) return (int)cudaErrorInvalidValue;
    if (epc % 4 == 0 && aligned16(x) && out_aligned<OutT>(y)) {
        const int64_t n4 = n / 4;
        const int grid = grid_for(n4, kThreads, kUnroll, 8);
        ew_channel_vec_kernel<Op, OutT><<<grid, kThreads, 0, st>>>(x, y, n4, C, FastDiv((uint32_t)(epc / 4)), FastDiv((uint32_t)C),
                                                                      scale, offset, p);
        return (int)cudaGetLastError();
    }
    const int grid = grid_for((n + 3) / 4, kThreads, 1, 8);
    ew_channel_generic_kernel<Op, OutT><<<grid, kThreads, 0, st>>>(x, y, n, epc, C, FastDiv((uint32_t)epc), FastDiv((uint32_t)C),
                                                                      scale, offset, p);
    return (int)cudaGetLastError();
}

static inline bool valid_fp_format(int E, int M) {
    // The reference forms min_subnormal with an int shift `1 << (2^(E-1) + M - 2)` (common.cuh:209): defined only
    // while that shift is in 0..30, which also keeps every constant a normal fp32.  E4M3, E5M2, E5M10 (fp16) are inside.
    if (E < 1 || E > 5 || M < 0 || M > 22) return false;
    const int k = (1 << (E - 1)) + M - 2;
    return k >= 0 && k <= 30;
}

}  // namespace ppqb

using namespace ppqb;

// TMA-staged variant lives in fakequant_tma.cu
int launch_linear_quant_t_tma(const float *x, float *y, int64_t n, const float *scale, const float *offset,
                              int qmin, int qmax, cudaStream_t st);

extern "C" {

int ppq_b200_linear_quant_t(const float *x, float *y, int64_t n, const float *scale, const float *offset,
                            int qmin, int qmax, int rounding, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (qmin > qmax) return (int)cudaErrorInvalidValue;
    if (rounding == RND_HALF_EVEN) {
        if (variant_of(kVarLinearT) == 1 && n >= 4096 && aligned16(x) && aligned16(y))
            return launch_linear_quant_t_tma(x, y, n, scale, offset, qmin, qmax, st);
        return launch_tensor<LinearOp<0>, float>(x, y, n, scale, offset, {qmin, qmax, 0}, st);
    }
    return launch_tensor<LinearOp<-1>, float>(x, y, n, scale, offset, {qmin, qmax, rounding}, st);
}

int ppq_b200_linear_quant_c(const float *x, float *y, int64_t n, int64_t epc, int C, const float *scale, const float *offset,
                            int qmin, int qmax, int rounding, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (qmin > qmax) return (int)cudaErrorInvalidValue;
    if (rounding == RND_HALF_EVEN)
        return launch_channel<LinearOp<0>, float>(x, y, n, epc, C, scale, offset, {qmin, qmax, 0}, st);
    return launch_channel<LinearOp<-1>, float>(x, y, n, epc, C, scale, offset, {qmin, qmax, rounding}, st);
}

static int toint_bits_ok(int out_bits) { return out_bits == 8 || out_bits == 32; }

int ppq_b200_linear_quant_t_toint(const float *x, void *q, int out_bits, int64_t n, const float *scale, const float *offset,
                                  int qmin, int qmax, int rounding, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (qmin > qmax) return (int)cudaErrorInvalidValue;
    if (!toint_bits_ok(out_bits)) return (int)cudaErrorInvalidValue;
    if (out_bits == 8) return launch_tensor<LinearOp<-1>, int8_t>(x, (int8_t *)q, n, scale, offset, {qmin, qmax, rounding}, st);
    return launch_tensor<LinearOp<-1>, int32_t>(x, (int32_t *)q, n, scale, offset, {qmin, qmax, rounding}, st);
}

int ppq_b200_linear_quant_c_toint(const float *x, void *q, int out_bits, int64_t n, int64_t epc, int C,
                                  const float *scale, const float *offset, int qmin, int qmax, int rounding, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (qmin > qmax) return (int)cudaErrorInvalidValue;
    if (!toint_bits_ok(out_bits)) return (int)cudaErrorInvalidValue;
    if (out_bits == 8) return launch_channel<LinearOp<-1>, int8_t>(x, (int8_t *)q, n, epc, C, scale, offset, {qmin, qmax, rounding}, st);
    return launch_channel<LinearOp<-1>, int32_t>(x, (int32_t *)q, n, epc, C, scale, offset, {qmin, qmax, rounding}, st);
}

int ppq_b200_float_quant_t(const float *x, float *y, int64_t n, const float *scale, const float *offset,
                           int exponent, int mantissa, float clip_min, float clip_max, int rounding, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!valid_fp_format(exponent, mantissa)) return (int)cudaErrorInvalidValue;
    if (rounding == RND_HALF_EVEN)
        return launch_tensor<FloatOp<0>, float>(x, y, n, scale, offset, {exponent, mantissa, 0, clip_min, clip_max}, st);
    return launch_tensor<FloatOp<-1>, float>(x, y, n, scale, offset, {exponent, mantissa, rounding, clip_min, clip_max}, st);
}

int ppq_b200_float_quant_c(const float *x, float *y, int64_t n, int64_t epc, int C, const float *scale, const float *offset,
                           int exponent, int mantissa, float clip_min, float clip_max, int rounding, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!valid_fp_format(exponent, mantissa)) return (int)cudaErrorInvalidValue;
    if (rounding == RND_HALF_EVEN)
        return launch_channel<FloatOp<0>, float>(x, y, n, epc, C, scale, offset, {exponent, mantissa, 0, clip_min, clip_max}, st);
    return launch_channel<FloatOp<-1>, float>(x, y, n, epc, C, scale, offset, {exponent, mantissa, rounding, clip_min, clip_max}, st);
}

}  // extern "C"
