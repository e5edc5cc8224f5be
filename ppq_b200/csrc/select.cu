// select.cu -- exact order statistics without sorting (sm_100a), behind the C ABI of include/ppq_b200.h.
//
//   ppq_b200_quantile_t   replaces Quantile_T (/root/reference/ppq/csrc/cuda/sort.cu:6-20, 42-59): the reference clones the
//                         tensor, thrust::sort's the clone and reads sorted[clip(rn(n*q))] and sorted[clip(rn(n*(1-q)))].
//   ppq_b200_isotone_t    replaces Isotone_T (sort.cu:23-40, 61-73): sorted[n-1], sorted[n-2], sorted[0], sorted[1].
//
// Here: MSD radix *select* on the order-preserving 32-bit key of each float (the same total order thrust's radix sort uses:
// -NaN < -inf < ... < -0 == +0 < ... < +inf < +NaN; the two zeros share one key, see order_key), 11 + 11 + 10 bits, both requested
// ranks resolved together.  Three
// streaming passes over the input (12 B/element, shared-memory privatised digit histograms) instead of a clone plus a full
// device sort; no allocation -- the caller provides a small workspace.  The result is the identical element, bit for bit, except that a
// selected zero is always reported as +0.0 (equal as a float to whichever zero the reference's sort left at that index).
#include "common.cuh"
#include "../../include/ppq_b200.h"

namespace ppqb {

constexpr int kSelThreads = 1024;
constexpr int kDigits = 2048;                  // 11-bit digits (the last pass uses 10 bits)

struct SelectState {                           // lives in the caller's workspace
    unsigned long long hist[2][kDigits];       // per rank: digit histogram of the current pass
    unsigned int prefix[2];                    // key bits resolved so far (high bits)
    unsigned int done;                         // CTAs that have flushed their digits in the current pass
    unsigned int pad_;
    long long rank[2];                         // remaining rank inside the current prefix bucket
    long long ranks_in[4];
};

__device__ __forceinline__ uint32_t order_key(float v) {
    uint32_t b = __float_as_uint(v);
    if (b == 0x80000000u) b = 0u;          // -0.0 and +0.0 are one key, as in CUB's radix sort (which thrust::sort dispatches to);
                                           // the selected zero is reported as +0.0 (the reference reports whichever zero its stable
                                           // sort left at that index -- equal as floats, the sign is not recoverable without a sort)
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void select_init_kernel(SelectState *st, long long r0, long long r1) {
    for (int i = threadIdx.x; i < 2 * kDigits; i += blockDim.x) (&st->hist[0][0])[i] = 0ull;
    if (threadIdx.x == 0) { st->prefix[0] = st->prefix[1] = 0u; st->done = 0u; st->rank[0] = r0; st->rank[1] = r1; }
}

// Block-wide prefix sum over the digit histogram of each rank (kDigits / TPB consecutive digits per thread, warp shuffles): pick the digit
// bucket that contains the rank, extend the prefix, clear the histograms for the next pass.  Run by the LAST CTA of a pass to finish its flush.
template <int PASS, int TPB>
__device__ __noinline__ void select_scan(SelectState *st, float *out, int out_stride) {
    constexpr int shift = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
    constexpr int D = kDigits / TPB;                                    // digits per thread: 2 (1024 threads) or 8 (256 threads)
    __shared__ unsigned long long warp_tot[TPB / 32];
    __shared__ unsigned int new_prefix[2];
    __shared__ long long new_rank[2];
    const bool same = st->prefix[0] == st->prefix[1];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int r = 0; r < 2; r++) {
        const unsigned long long *h = st->hist[(r == 1 && same) ? 0 : r];
        const unsigned long long k = (unsigned long long)st->rank[r];
        unsigned long long c[D], mine = 0;
#pragma unroll
        for (int j = 0; j < D; j++) { c[j] = __ldcg(h + D * t + j); mine += c[j]; }        // written by other CTAs' atomics: read at L2
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        unsigned long long before = incl - mine;
        for (int i = 0; i < w; i++) before += warp_tot[i];
        // the bucket d with  before(d) <= k < before(d) + count(d); ranks are always < total, so exactly one (thread, digit) matches
#pragma unroll
        for (int j = 0; j < D; j++) {
            if (k >= before && k < before + c[j]) { new_prefix[r] = st->prefix[r] | ((unsigned int)(D * t + j) << shift); new_rank[r] = (long long)(k - before); }
            before += c[j];
        }
        __syncthreads();
    }
    for (int i = t; i < 2 * kDigits; i += TPB) (&st->hist[0][0])[i] = 0ull;
    if (t < 2) {
        st->prefix[t] = new_prefix[t];
        st->rank[t] = new_rank[t];
        if (PASS == 2) out[t * out_stride] = key_to_float(new_prefix[t]);
    }
    if (t == 0) st->done = 0u;
}

// PASS 0: digit = key[31:21];  PASS 1: key[20:10] among keys whose top 11 bits match;  PASS 2: key[9:0] among top-22 matches.
// Pass 0 is a histogram of every element (unconditional shared red, as in collectors.cu): two 1024-thread CTAs per SM, grid-stride.
// In passes 1 and 2 only the elements of the one or two buckets chosen so far count: a vector whose four elements all miss both prefixes
// -- the common case, the requested ranks sit in the tails -- costs one AND + three compares per element and no shared-memory traffic,
// so these passes use the launch shape of the min/max collector: 256-thread CTAs, 8 per SM, every warp walking contiguous 2 KB segments.
template <int PASS, int TPB>
__global__ void __launch_bounds__(TPB, PASS == 0 ? 2 : 6)            // 32 / 40 registers; the scan of the last CTA may spill, it runs once
select_hist_kernel(const float *__restrict__ x, int64_t n, SelectState *__restrict__ st, float *out, int out_stride) {
    __shared__ int sh[2][kDigits];
    __shared__ bool is_last;
    for (int i = threadIdx.x; i < 2 * kDigits; i += TPB) (&sh[0][0])[i] = 0;
    __syncthreads();
    constexpr int shift = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
    constexpr uint32_t dmask = PASS == 2 ? 0x3FFu : 0x7FFu;
    constexpr uint32_t pmask = PASS == 0 ? 0u : (PASS == 1 ? 0xFFE00000u : 0xFFFFFC00u);
    const uint32_t p0 = st->prefix[0], p1 = st->prefix[1];
    const bool same = (p0 == p1);
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(&sh[0][0]), a1 = (uint32_t)__cvta_generic_to_shared(&sh[1][0]);
    auto count = [&](uint32_t k) {
        const uint32_t hi = k & pmask, d = (k >> shift) & dmask;
        if (PASS == 0) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0 + d * 4u) : "memory"); return; }
        if (hi == p0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0 + d * 4u) : "memory");
        if (!same && hi == p1) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1 + d * 4u) : "memory");
    };
    // Fast rejection on the RAW bits (passes 1-2): key & pmask == p  <=>  bits & pmask == raw(p), because the key transform only depends on
    // the sign, which the prefix fixes.  -0.0 is the one exception (its key is +0's): when a prefix is the bucket of +0 the raw pattern of
    // -0.0 is accepted too; a false positive only costs the exact test in count().  One AND + three compares per element.
    auto raw_of = [&](uint32_t p) { return (p & 0x80000000u) ? (p & 0x7FFFFFFFu) : (~p & pmask); };
    const uint32_t c0 = raw_of(p0), c1 = raw_of(p1), c2 = (p0 == 0x80000000u || p1 == 0x80000000u) ? 0x80000000u : c0;
    auto hit = [&](float f) { const uint32_t t = __float_as_uint(f) & pmask; return (t == c0) | (t == c1) | (t == c2); };
    auto visit4 = [&](const float4 &v) {
        if (PASS != 0 && !(hit(v.x) | hit(v.y) | hit(v.z) | hit(v.w))) return;
        count(order_key(v.x)); count(order_key(v.y)); count(order_key(v.z)); count(order_key(v.w));
    };
    const int64_t first = (int64_t)blockIdx.x * TPB + threadIdx.x, stride = (int64_t)gridDim.x * TPB;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
        const int64_t n4 = n >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        constexpr int U = 4;
        if (PASS == 0) {
            for (int64_t i = first; i < n4; i += U * stride) {
                float4 v[U];
#pragma unroll
                for (int j = 0; j < U; j++) if (i + j * stride < n4) v[j] = ld_stream4(x4 + i + j * stride);
#pragma unroll
                for (int j = 0; j < U; j++) if (i + j * stride < n4) visit4(v[j]);
            }
        } else {
            // warp-contiguous segments of 32 lanes x U vectors (2 KB)
            const int64_t lane = threadIdx.x & 31, warps = stride >> 5;
            for (int64_t sg = first >> 5; sg * (32 * U) < n4; sg += warps) {
                const int64_t base = sg * (32 * U) + lane;
                float4 v[U];
#pragma unroll
                for (int j = 0; j < U; j++) if (base + j * 32 < n4) v[j] = ld_stream4(x4 + base + j * 32);
#pragma unroll
                for (int j = 0; j < U; j++) if (base + j * 32 < n4) visit4(v[j]);
            }
        }
        const int64_t t = (n4 << 2) + first;
        if (t < n) count(order_key(x[t]));
    } else {
        for (int64_t i = first; i < n; i += stride) count(order_key(ld_stream1(x + i)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kDigits; i += TPB) {
        if (sh[0][i]) atomicAdd(&st->hist[0][i], (unsigned long long)sh[0][i]);
        if (!same && sh[1][i]) atomicAdd(&st->hist[1][i], (unsigned long long)sh[1][i]);
    }
    // the last CTA to get here resolves the pass (no separate scan launch): every CTA's atomics are ordered before its ticket
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        select_scan<PASS, TPB>(st, out, out_stride);
    }
}

constexpr int kFilterThreads = 256;

static int select_two(const float *x, int64_t n, long long r0, long long r1, float *out, int out_stride, SelectState *st, cudaStream_t s) {
    // pass 0 needs 32 registers per thread: two 1024-thread CTAs fit on an SM
    int64_t g0 = (n + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
    if (g0 > 2 * kSMs) g0 = 2 * kSMs;
    if (g0 < 1) g0 = 1;
    int64_t g = (n + (int64_t)kFilterThreads * 16 - 1) / ((int64_t)kFilterThreads * 16);
    if (g > 6 * kSMs) g = 6 * kSMs;                                       // 40 registers, 17.5 KB of smem: six CTAs per SM, one wave
    if (g < 1) g = 1;
    select_init_kernel<<<1, 1024, 0, s>>>(st, r0, r1);
    select_hist_kernel<0, kSelThreads><<<(int)g0, kSelThreads, 0, s>>>(x, n, st, out, out_stride);
    select_hist_kernel<1, kFilterThreads><<<(int)g, kFilterThreads, 0, s>>>(x, n, st, out, out_stride);
    select_hist_kernel<2, kFilterThreads><<<(int)g, kFilterThreads, 0, s>>>(x, n, st, out, out_stride);
    return (int)cudaGetLastError();
}

}  // namespace ppqb

using namespace ppqb;

extern "C" {

int64_t ppq_b200_quantile_workspace_bytes(void) { return (int64_t)sizeof(SelectState); }

int ppq_b200_quantile_t(const float *x, int64_t n, float q, float *out2, void *workspace, void *stream) {
    if (n <= 0 || !x || !out2 || !workspace) return (int)cudaErrorInvalidValue;
    // index arithmetic of _Quantile_T (sort.cu:13-19): int64 * float -> float, __float2int_rn (half-even, saturating), CLIP
    const float fa = (float)n * q;
    const float fb = (float)n * (1.0f - q);
    auto rn = [](float v) -> long long {
        if (v != v) return 0;
        if (v >= 2147483648.0f) return 2147483647LL;
        if (v <= -2147483648.0f) return -2147483648LL;
        return (long long)__builtin_nearbyintf(v);
    };
    long long a = rn(fa), b = rn(fb);
    // the reference clips with CLIP<int>(pos, 0, n - 1) where n - 1 is converted to int
    const long long last = (long long)(int)(n - 1);
    a = a > last ? last : (a < 0 ? 0 : a);
    b = b > last ? last : (b < 0 ? 0 : b);
    return select_two(x, n, a, b, out2, 1, (SelectState *)workspace, (cudaStream_t)stream);
}

int ppq_b200_isotone_t(const float *x, int64_t n, float *out4, void *workspace, void *stream) {
    if (n <= 0 || !x || !out4 || !workspace) return (int)cudaErrorInvalidValue;
    SelectState *st = (SelectState *)workspace;
    if (n == 1) {
        int rc = select_two(x, n, 0, 0, out4, 1, st, (cudaStream_t)stream);
        if (rc) return rc;
        return select_two(x, n, 0, 0, out4 + 2, 1, st, (cudaStream_t)stream);
    }
    // out = { sorted[n-1], sorted[n-2], sorted[0], sorted[1] }
    int rc = select_two(x, n, n - 1, n - 2, out4, 1, st, (cudaStream_t)stream);
    if (rc) return rc;
    return select_two(x, n, 0, 1, out4 + 2, 1, st, (cudaStream_t)stream);
}

}  // extern "C"
