// ops.cuh -- per-element quantisation operators shared by the fake-quant kernels (fakequant.cu, fakequant_tma.cu,
// train.cu).  LinearOp == QuantizeScalar + DequantizeScalar (/root/reference/ppq/csrc/cuda/common.cuh:116-147),
// FloatOp == QuantizeScalarFloating (common.cuh:154-226) followed by the float dequantise of floating.cu:50-53.
//
// Every operator has a scalar form and a 4-wide form.  The 4-wide form computes the four exact quotients as one
// straight-line block (four independent 5-deep FMA chains the scheduler can interleave) and tests them for the rare
// "needs the IEEE slow path" condition once per vector instead of once per element.
#pragma once
#include "common.cuh"

namespace ppqb {

// MODE >= 0: compile-time rounding mode; MODE == -1: run-time mode (uniform switch).
template <int MODE>
struct LinearOp {
    struct Params { int lo, hi, mode; };
    ExactDiv d; int o, lo, hi, mode;
    __device__ __forceinline__ LinearOp(const Params &p, float s, float off) : lo(p.lo), hi(p.hi), mode(p.mode) {
        d.init(s);
        o = offset_to_int(off);
    }
    __device__ __forceinline__ int finish(float t) const {
        int q;
        if constexpr (MODE >= 0) q = round2int<MODE>(t); else q = round2int_dyn(t, mode);
        return min(max(q + o, lo), hi);                              // int32 wrap on the add, as on the reference device path
    }
    __device__ __forceinline__ int quant(float x) const { return finish(d.div(x)); }
    __device__ __forceinline__ float dequant(int q) const { return __fmul_rn(__int2float_rn(q - o), d.s); }
    __device__ __forceinline__ float apply(float x) const { return dequant(quant(x)); }
    __device__ __forceinline__ int4 quant4(const float4 &v) const {
        const float4 t = d.div4(v);
        return make_int4(finish(t.x), finish(t.y), finish(t.z), finish(t.w));
    }
    __device__ __forceinline__ float4 apply4(const float4 &v) const {
        const int4 q = quant4(v);
        return make_float4(dequant(q.x), dequant(q.y), dequant(q.z), dequant(q.w));
    }
};

template <int MODE>
struct FloatOp {
    struct Params { int E, M, mode; float cmin, cmax; };
    ExactDiv d; float off, hi, lo, cmin, cmax, min_sub, inv_min_sub;
    uint32_t sub_thresh_bits, half_minus1, keep_mask; int M, mode;
    __device__ __forceinline__ FloatOp(const Params &p, float s, float o) : off(o), cmin(p.cmin), cmax(p.cmax), M(p.M), mode(p.mode) {
        d.init(s);
        const int emin = -(1 << (p.E - 1)) + 1, emax = 1 << (p.E - 1);
        const uint32_t top = ~(0x007FFFFFu >> p.M) & 0x007FFFFFu;
        const float tmax = __uint_as_float((uint32_t)((emax + 127) << 23) + top);   // E4M3: 480
        hi = fminf(p.cmax, tmax);
        lo = fmaxf(p.cmin, -tmax);
        const int k = (1 << (p.E - 1)) + p.M - 2;                                   // min subnormal = 2^-k
        min_sub = __uint_as_float((uint32_t)(127 - k) << 23);
        inv_min_sub = __uint_as_float((uint32_t)(127 + k) << 23);                   // u / 2^-k == u * 2^k exactly
        sub_thresh_bits = (uint32_t)(emin + 1 + 127) << 23;                         // |u| < 2^(emin+1) -> subnormal grid
        half_minus1 = (1u << (22 - p.M)) - 1u;
        keep_mask = ~((1u << (23 - p.M)) - 1u);
    }
    static constexpr float kDivLimit = 1.15e18f;                                    // ~2^60: beyond this use div.rn
    // u = x / s already computed exactly
    __device__ __forceinline__ float grid(float u) const {
        if (u > hi) return hi;
        if (u < lo) return lo;
        const uint32_t b = __float_as_uint(u);
        const uint32_t sign = b & 0x80000000u, mag = b & 0x7FFFFFFFu;
        if (mag < sub_thresh_bits) {
            int r;
            const float v = __fmul_rn(u, inv_min_sub);
            if constexpr (MODE >= 0) r = round2int<MODE>(v); else r = round2int_dyn(v, mode);
            return __fmul_rn(__int2float_rn(r), min_sub);
        }
        uint32_t rb;
        if constexpr (MODE == RND_HALF_EVEN) {
            // rint(frac) with frac in [0,1): 1 iff the discarded bits exceed one half (an exact tie gives 0)
            rb = ((mag + half_minus1) & keep_mask) + sign;
        } else {
            const uint32_t mant = b & 0x007FFFFFu;
            const float frac = __fsub_rn(__uint_as_float(((mant << M) & 0x007FFFFFu) + 0x3F800000u), 1.0f);
            int r;
            if constexpr (MODE >= 0) r = round2int<MODE>(frac); else r = round2int_dyn(frac, mode);
            rb = sign + (((mant >> (23 - M)) + (uint32_t)r) << (23 - M)) + (b & 0x7F800000u);
        }
        const float q = __uint_as_float(rb);
        return q > cmax ? cmax : (q < cmin ? cmin : q);
    }
    __device__ __forceinline__ float dequant(float q) const { return __fmul_rn(__fsub_rn(q, off), d.s); }
    __device__ __forceinline__ float apply(float x) const { return dequant(grid(d.div(x, kDivLimit))); }
    __device__ __forceinline__ float4 apply4(const float4 &v) const {
        const float4 u = d.div4(v, kDivLimit);
        return make_float4(dequant(grid(u.x)), dequant(grid(u.y)), dequant(grid(u.z)), dequant(grid(u.w)));
    }
};

}  // namespace ppqb
