// ops.cuh -- per-element quantisation operators shared by the fake-quant kernels (fakequant.cu, fakequant_tma.cu,
// train.cu).  LinearOp == QuantizeScalar + DequantizeScalar (/root/reference/ppq/csrc/cuda/common.cuh:116-147),
// FloatOp == QuantizeScalarFloating (common.cuh:154-226) followed by the float dequantise of floating.cu:50-53.
//
// Structure: Params (host, by value) -> Plan (device, once per thread: everything that depends only on the format /
// clip range) -> Op (device, once per tensor or per channel row: the scale's exact reciprocal, the offset).
// Every operator has a scalar form and a 4-wide form.  The 4-wide form computes the four exact quotients as one
// straight-line block (four independent 5-deep FMA chains the scheduler can interleave) and tests them for the rare
// "needs the IEEE slow path" condition once per vector instead of once per element.
#pragma once
#include "common.cuh"

namespace ppqb {

__device__ __forceinline__ float fmin_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmax_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// MODE >= 0: compile-time rounding mode; MODE == -1: run-time mode (uniform switch).
template <int MODE>
struct LinearOp {
    struct Params { int lo, hi, mode; };
    static constexpr bool kLight = (MODE == RND_HALF_EVEN);          // few registers per element: the 8-deep load variant pays off
    struct Plan {
        int lo, hi, mode;
        __device__ __forceinline__ explicit Plan(const Params &p) : lo(p.lo), hi(p.hi), mode(p.mode) {}
    };
    ExactDiv d; int o, lo, hi, mode;
    __device__ __forceinline__ LinearOp(const Plan &p, float s, float off) : lo(p.lo), hi(p.hi), mode(p.mode) {
        d.init(s);
        o = offset_to_int(off);
    }
    __device__ __forceinline__ void rebind(float s, float off) { d.init(s); o = offset_to_int(off); }     // same plan, next channel
    // the per-channel part of the operator as a 16-byte shared-memory table entry (built once per row of a tile, see ew_channel_table_kernel)
    static __device__ __forceinline__ float4 entry(float s, float off) {
        ExactDiv e(s);
        return make_float4(e.s, e.r, __int_as_float(offset_to_int(off)), 0.f);
    }
    __device__ __forceinline__ LinearOp(const Plan &p, const float4 &e) : lo(p.lo), hi(p.hi), mode(p.mode) { d.s = e.x; d.r = e.y; o = __float_as_int(e.z); }
    __device__ __forceinline__ void rebind(const float4 &e) { d.s = e.x; d.r = e.y; o = __float_as_int(e.z); }
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
    // lane j of the vector belongs to operator o[j] (channel-last layouts)
    static __device__ __forceinline__ int4 quant4x(const LinearOp (&o)[4], const float4 &v) {
        const float4 t = exact_div4x(o[0].d, o[1].d, o[2].d, o[3].d, v);
        return make_int4(o[0].finish(t.x), o[1].finish(t.y), o[2].finish(t.z), o[3].finish(t.w));
    }
    static __device__ __forceinline__ float4 apply4x(const LinearOp (&o)[4], const float4 &v) {
        const int4 q = quant4x(o, v);
        return make_float4(o[0].dequant(q.x), o[1].dequant(q.y), o[2].dequant(q.z), o[3].dequant(q.w));
    }
};

// FAST (compile-time MODE, chosen on the host by float_fast_path_ok): the branch-free path, valid when both saturation bounds are
// fixed points of the rounding (lie on the FP(E,M) grid), so that round(clamp(u)) equals the reference's early returns and its
// final CLIP is a no-op.  In the normal range the reference rounds the DISCARDED mantissa bits, taken as a fraction in [0, 1), with
// _round2int (common.cuh:218-222): the fraction is never negative, so every policy reduces to "carry iff d > half" (HALF_EVEN -- rint(0.5)
// == 0 --, HALF_DOWN, HALF_TOWARDS_ZERO), "carry iff d >= half" (HALF_UP, HALF_FAR_FROM_ZERO, TO_NEAR_INT), "carry iff d > 0" (UP) or
// "never" (DOWN): one integer add of a per-mode constant and a mask.  The sub-normal grid rounds the SIGNED quotient with the policy itself.
template <int MODE, bool FAST = false>
struct FloatOp {
    struct Params { int E, M, mode; float cmin, cmax; };
    static constexpr bool kLight = FAST;
    struct Plan {
        float hi, lo, cmin, cmax, min_sub, inv_min_sub, sub_magic, sub_thresh;
        uint32_t sub_thresh_bits, half_minus1, keep_mask, carry_add; int M, mode;
        __device__ __forceinline__ explicit Plan(const Params &p) : cmin(p.cmin), cmax(p.cmax), M(p.M), mode(p.mode) {
            const int emin = -(1 << (p.E - 1)) + 1, emax = 1 << (p.E - 1);
            const uint32_t top = ~(0x007FFFFFu >> p.M) & 0x007FFFFFu;
            const float tmax = __uint_as_float((uint32_t)((emax + 127) << 23) + top);   // E4M3: 480
            hi = fminf(p.cmax, tmax);
            lo = fmaxf(p.cmin, -tmax);
            const int k = (1 << (p.E - 1)) + p.M - 2;                                   // min subnormal = 2^-k
            min_sub = __uint_as_float((uint32_t)(127 - k) << 23);
            inv_min_sub = __uint_as_float((uint32_t)(127 + k) << 23);                   // u / 2^-k == u * 2^k exactly
            sub_magic = __uint_as_float(((uint32_t)(127 + 23 - k) << 23) | 0x00400000u);  // 1.5 * 2^(23-k): ulp == 2^-k
            sub_thresh_bits = (uint32_t)(emin + 1 + 127) << 23;                         // |u| < 2^(emin+1) -> subnormal grid
            sub_thresh = __uint_as_float(sub_thresh_bits);
            half_minus1 = (1u << (22 - p.M)) - 1u;
            keep_mask = ~((1u << (23 - p.M)) - 1u);
            constexpr int m = MODE;
            carry_add = (m == RND_HALF_UP || m == RND_HALF_FAR_FROM_ZERO || m == RND_TO_NEAR_INT) ? (1u << (22 - p.M))
                      : (m == RND_UP ? (1u << (23 - p.M)) - 1u : (m == RND_DOWN ? 0u : half_minus1));
        }
    };
    const Plan &pl; ExactDiv d; float off;
    __device__ __forceinline__ FloatOp(const Plan &p, float s, float o) : pl(p), off(o) { d.init(s); }
    __device__ __forceinline__ void rebind(float s, float o) { d.init(s); off = o; }
    static __device__ __forceinline__ float4 entry(float s, float o) { ExactDiv e(s); return make_float4(e.s, e.r, o, 0.f); }
    __device__ __forceinline__ FloatOp(const Plan &p, const float4 &e) : pl(p), off(e.z) { d.s = e.x; d.r = e.y; }
    __device__ __forceinline__ void rebind(const float4 &e) { d.s = e.x; d.r = e.y; off = e.z; }
    static constexpr float kDivLimit = 1.15e18f;                                        // ~2^60: beyond this use div.rn
    // u = x / s already computed exactly; returns the value on the FP(E,M) grid
    __device__ __forceinline__ float grid(float u) const {
        if constexpr (FAST) {
            {
                const float uc = fmin_nan(fmax_nan(u, pl.lo), pl.hi);                   // NaN stays NaN (canonical 0x7FFFFFFF), like the comparisons upstream
                // normal range: round the magnitude's discarded mantissa bits half-DOWN (an exact tie keeps the lower value,
                // because rint(0.5) == 0 upstream); the carry may bump the exponent.  Adding to the signed pattern is the same as
                // adding to the magnitude: a clamped finite value cannot carry into bit 31, and the canonical NaN has sign 0
                // (its carry into bit 31 reproduces the reference's -0.0 for NaN).
                const uint32_t nb = (__float_as_uint(uc) + pl.carry_add) & pl.keep_mask;
                // subnormal range: ties-to-even on the 2^-k grid, computed with the add-magic-constant trick (sign of zero lost,
                // exactly like the int round trip upstream); the other policies round the signed quotient with the policy itself
                float sub;
                if constexpr (MODE == RND_HALF_EVEN) sub = __fsub_rn(__fadd_rn(uc, pl.sub_magic), pl.sub_magic);
                else sub = __fmul_rn(__int2float_rn(round2int<MODE>(__fmul_rn(uc, pl.inv_min_sub))), pl.min_sub);
                return fabsf(uc) < pl.sub_thresh ? sub : __uint_as_float(nb);            // false for NaN -> nb, as with the integer compare
            }
        }
        if (u > pl.hi) return pl.hi;
        if (u < pl.lo) return pl.lo;
        const uint32_t b = __float_as_uint(u);
        const uint32_t sign = b & 0x80000000u, mag = b & 0x7FFFFFFFu;
        if (mag < pl.sub_thresh_bits) {
            int r;
            const float v = __fmul_rn(u, pl.inv_min_sub);
            if constexpr (MODE >= 0) r = round2int<MODE>(v); else r = round2int_dyn(v, pl.mode);
            return __fmul_rn(__int2float_rn(r), pl.min_sub);
        }
        uint32_t rb;
        if constexpr (MODE == RND_HALF_EVEN) {
            rb = ((mag + pl.half_minus1) & pl.keep_mask) + sign;
        } else {
            const uint32_t mant = b & 0x007FFFFFu;
            const float frac = __fsub_rn(__uint_as_float(((mant << pl.M) & 0x007FFFFFu) + 0x3F800000u), 1.0f);
            int r;
            if constexpr (MODE >= 0) r = round2int<MODE>(frac); else r = round2int_dyn(frac, pl.mode);
            rb = sign + (((mant >> (23 - pl.M)) + (uint32_t)r) << (23 - pl.M)) + (b & 0x7F800000u);
        }
        const float q = __uint_as_float(rb);
        return q > pl.cmax ? pl.cmax : (q < pl.cmin ? pl.cmin : q);
    }
    __device__ __forceinline__ float dequant(float q) const { return __fmul_rn(__fsub_rn(q, off), d.s); }
    __device__ __forceinline__ float apply(float x) const { return dequant(grid(d.div(x, kDivLimit))); }
    __device__ __forceinline__ float4 apply4(const float4 &v) const {
        const float4 u = d.div4(v, kDivLimit);
        return make_float4(dequant(grid(u.x)), dequant(grid(u.y)), dequant(grid(u.z)), dequant(grid(u.w)));
    }
    static __device__ __forceinline__ float4 apply4x(const FloatOp (&o)[4], const float4 &v) {
        const float4 u = exact_div4x(o[0].d, o[1].d, o[2].d, o[3].d, v, kDivLimit);
        return make_float4(o[0].dequant(o[0].grid(u.x)), o[1].dequant(o[1].grid(u.y)), o[2].dequant(o[2].grid(u.z)), o[3].dequant(o[3].grid(u.w)));
    }
};

}  // namespace ppqb
