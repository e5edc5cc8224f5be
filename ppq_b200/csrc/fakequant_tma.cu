// fakequant_tma.cu -- TMA-staged variant of the per-tensor INT fake-quant (variant 1 of "linear_quant_t").
//
// Same numerics as ew_tensor_kernel<LinearOp<0>> (fakequant.cu); different data movement:
//   producer warp : cp.async.bulk (1-D TMA, SASS UBLKCP) global -> shared ring of kStages tiles, completion on
//                   mbarriers (complete_tx); after the consumers are done with a tile, bulk store shared -> global.
//   8 consumer warps: wait on the tile's "full" barrier, quantise in place in shared memory (2 float4 per thread),
//                   fence.proxy.async, arrive on the tile's "done" barrier.
// No register staging of the stream and only one thread issues memory instructions, so the LSU/issue pressure of the
// load/store path disappears; kStages x 8 KB x resident CTAs bytes are in flight per SM.
// Kept as a selectable variant so that bench.py / ncu can A-B it against the LDG.128 kernel (profiles/).
#include "ops.cuh"

namespace ppqb {

constexpr int kConsumers = 256;
constexpr int kTmaThreads = kConsumers + 32;
constexpr int kTileVec = 1024;                  // float4 per tile  (16 KB)
constexpr int kTileBytes = kTileVec * 16;
constexpr int kStages = 4;                      // 64 KB of shared memory per CTA -> 3 CTAs / SM, 192 KB of tiles in flight per SM

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kTmaThreads)
linear_quant_t_tma_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n,
                          const float *__restrict__ scale, const float *__restrict__ offset, int lo, int hi) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);                       // [kStages][kTileVec]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kStages * kTileBytes);   // TMA landed
    uint64_t *done = full + kStages;                                                // consumers finished

    const int64_t n4 = n >> 2;
    const int64_t num_tiles = (n4 + kTileVec - 1) / kTileVec;
    const int64_t my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(full + s, 1); mbar_init(done + s, kConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_bytes = [&](int64_t k) -> uint32_t {
        const int64_t t = blockIdx.x + k * (int64_t)gridDim.x;
        const int64_t v = n4 - t * kTileVec;
        return (uint32_t)((v < kTileVec ? v : kTileVec) * 16);
    };

    if (threadIdx.x >= kConsumers) {
        // ---------------- producer warp (one elected lane) ----------------
        if (threadIdx.x == kConsumers) {
            const int64_t pre = my_tiles < kStages ? my_tiles : kStages;
            for (int64_t k = 0; k < pre; k++) {
                const int64_t t = blockIdx.x + k * (int64_t)gridDim.x;
                mbar_expect_tx(full + k, tile_bytes(k));
                tma_load_1d(tiles + k * kTileVec, reinterpret_cast<const float4 *>(x) + t * kTileVec, tile_bytes(k), full + k);
            }
            for (int64_t k = 0; k < my_tiles; k++) {
                const int s = (int)(k % kStages);
                const uint32_t ph = (uint32_t)((k / kStages) & 1);
                const int64_t t = blockIdx.x + k * (int64_t)gridDim.x;
                mbar_wait(done + s, ph);                                          // tile k quantised in shared memory
                tma_store_1d(reinterpret_cast<float4 *>(y) + t * kTileVec, tiles + s * kTileVec, tile_bytes(k));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // Refill the slot of the PREVIOUS tile: its store was committed one iteration ago, so waiting for "all but the newest
                // store have finished reading shared memory" does not stall on the store just issued.
                if (k >= 1 && (k - 1) + kStages < my_tiles) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int64_t k2 = (k - 1) + kStages;
                    const int s2 = (int)((k - 1) % kStages);
                    const int64_t t2 = blockIdx.x + k2 * (int64_t)gridDim.x;
                    mbar_expect_tx(full + s2, tile_bytes(k2));
                    tma_load_1d(tiles + s2 * kTileVec, reinterpret_cast<const float4 *>(x) + t2 * kTileVec, tile_bytes(k2), full + s2);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");           // all stores complete before exit
        }
    } else {
        // ---------------- consumers ----------------
        const LinearOp<0>::Plan plan(LinearOp<0>::Params{lo, hi, 0});
        const LinearOp<0> op(plan, __ldg(scale), __ldg(offset));
        for (int64_t k = 0; k < my_tiles; k++) {
            const int s = (int)(k % kStages);
            const uint32_t ph = (uint32_t)((k / kStages) & 1);
            const int nv = (int)(tile_bytes(k) >> 4);
            mbar_wait(full + s, ph);
            float4 *tile = tiles + s * kTileVec;
#pragma unroll
            for (int j = 0; j < kTileVec / kConsumers; j++) {
                const int vi = threadIdx.x + j * kConsumers;
                if (vi < nv) {
                    tile[vi] = op.apply4(tile[vi]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to TMA
            mbar_arrive(done + s);
        }
        // <= 3 leftover scalars
        if (blockIdx.x == 0) {
            const int64_t t = (n4 << 2) + threadIdx.x;
            if (t < n) y[t] = op.apply(x[t]);
        }
    }
}

}  // namespace ppqb

using namespace ppqb;

int launch_linear_quant_t_tma(const float *x, float *y, int64_t n, const float *scale, const float *offset,
                              int qmin, int qmax, cudaStream_t st) {
    const int smem = kStages * kTileBytes + 2 * kStages * (int)sizeof(uint64_t);
    // the > 48 KB opt-in is per device, not per process: set it on every launch of this (non-default) variant
    if (cudaFuncSetAttribute(linear_quant_t_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return (int)cudaGetLastError();
    const int64_t n4 = n >> 2;
    const int64_t tiles = (n4 + kTileVec - 1) / kTileVec;
    int64_t grid = (int64_t)sm_count() * 3;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    linear_quant_t_tma_kernel<<<(int)grid, kTmaThreads, smem, st>>>(x, y, n, scale, offset, qmin, qmax);
    return (int)cudaGetLastError();
}
