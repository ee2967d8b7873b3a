// K1 / K2 with TMA bulk copies (cp.async.bulk, SASS UBLKCP) + an mbarrier producer/consumer ring — the A/B partner of
// stream_kernel<PackBf16> / <UnpackBf16> (bucket_kernels.cu).  Same arithmetic, same bytes; what changes is HOW the source
// gets on chip: one elected producer thread per CTA issues 16 KB bulk copies global -> shared memory that complete on an
// mbarrier (no registers, no LSU instructions for the loads), four stages deep, while four consumer warps convert from
// shared memory and store bf16 with 128-bit STG.
//
//   smem ring : kStages x 16 KB fp32 tiles (64 KB per CTA -> 3 CTAs per SM -> 192 KB in flight per SM)
//   full[s]   : producer arms it with expect_tx(16 KB); the bulk copy completes it            (TMA -> consumers)
//   empty[s]  : one arrive per consumer warp when the stage has been read                      (consumers -> producer)
//
// Whether this beats plain vectorised loads for a no-reuse streaming pass is an empirical question; bench.py's
// `roofline_more` reports both and DESIGN.md §3 states the outcome (round 1, 1 GiB cold: K1 0.96 vs 0.93, K2 0.87 vs 0.85 of
// the measured HBM peak; at 2-47 MB buckets the register path is faster by the ring's ~2 us fill/drain).
// dmlb_bucket_pack_f32_bf16 / dmlb_bucket_unpack_bf16_f32 dispatch to these from 32 Mi elements upward.
#include "dmlb_common.cuh"

namespace dmlb {

constexpr int kTmaStages = 4;
constexpr int kTileElems = 4096;                   // 16 KB of fp32 in, 8 KB of bf16 out
constexpr int kTileBytes = kTileElems * 4;
constexpr int kConsumerWarps = 4;
constexpr int kTmaThreads = (kConsumerWarps + 1) * 32;  // + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(kTmaThreads)
pack_bf16_tma_kernel(const float *__restrict__ src, uint16_t *__restrict__ dst, size_t n_tiles, float scale) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    __shared__ uint64_t full[kTmaStages], empty[kTmaStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===== producer: one elected lane feeds the ring =====
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % kTmaStages;
                const uint32_t round = it / kTmaStages;
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);  // consumers drained this stage's previous use
                mbar_expect_tx(&full[s], kTileBytes);
                tma_load_1d(tiles + (size_t)s * kTileElems, src + t * kTileElems, kTileBytes, &full[s]);
            }
        }
    } else {
        // ===== consumers: 4 warps, 128 threads x 8 elements per pass, 4 passes per tile =====
        const int ct = threadIdx.x;  // 0..127
        uint32_t it = 0;
        for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int s = it % kTmaStages;
            mbar_wait(&full[s], (it / kTmaStages) & 1);
            const float4 *tile = reinterpret_cast<const float4 *>(tiles + (size_t)s * kTileElems);
            uint4 *out = reinterpret_cast<uint4 *>(dst + t * kTileElems);
#pragma unroll
            for (int p = 0; p < kTileElems / (8 * kConsumerWarps * 32); ++p) {
                const int i = p * (kConsumerWarps * 32) + ct;  // 8-element group inside the tile
                const float4 a = tile[2 * i], b = tile[2 * i + 1];
                uint4 o;
                o.x = pack_bf16x2(a.x * scale, a.y * scale), o.y = pack_bf16x2(a.z * scale, a.w * scale);
                o.z = pack_bf16x2(b.x * scale, b.y * scale), o.w = pack_bf16x2(b.z * scale, b.w * scale);
                out[i] = o;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done reading the stage
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2 with TMA on both sides: bulk load of the bf16 tile (8 KB), consumers widen it into a shared-memory fp32 tile
// (16 KB, double buffered), one elected thread sends it out with a bulk STORE (cp.async.bulk.global.shared::cta).
// Two thirds of K2's traffic are writes; a bulk store hands the memory system whole 16 KB bursts instead of 128 STG.128
// per warp-pass.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTileBytesBf16 = kTileElems * 2;

__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory"); }

__global__ void __launch_bounds__(kTmaThreads)
unpack_bf16_tma_kernel(const uint16_t *__restrict__ src, float *__restrict__ dst, size_t n_tiles, float scale) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint16_t *in_tiles = reinterpret_cast<uint16_t *>(smem_raw);                                  // kStages x 8 KB
    float *out_tiles = reinterpret_cast<float *>(smem_raw + (size_t)kTmaStages * kTileBytesBf16);  // 2 x 16 KB
    __shared__ uint64_t full[kTmaStages], empty[kTmaStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % kTmaStages;
                const uint32_t round = it / kTmaStages;
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                mbar_expect_tx(&full[s], kTileBytesBf16);
                tma_load_1d(in_tiles + (size_t)s * kTileElems, src + t * kTileElems, kTileBytesBf16, &full[s]);
            }
        }
    } else {
        const int ct = threadIdx.x;  // 0..127
        uint32_t it = 0;
        for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int s = it % kTmaStages;
            float4 *out = reinterpret_cast<float4 *>(out_tiles + (size_t)(it & 1) * kTileElems);
            // the bulk store that last read this out buffer (iteration it-2) must be done reading shared memory
            if (ct == 0 && it >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            consumers_sync();
            mbar_wait(&full[s], (it / kTmaStages) & 1);
            const uint2 *in = reinterpret_cast<const uint2 *>(in_tiles + (size_t)s * kTileElems);
#pragma unroll
            for (int p = 0; p < kTileElems / (4 * kConsumerWarps * 32); ++p) {
                const int i = p * (kConsumerWarps * 32) + ct;  // 4-element group: LDS.64 in, STS.128 out, conflict-free
                const uint2 v = in[i];
                out[i] = make_float4(bf16_lo(v.x) * scale, bf16_hi(v.x) * scale, bf16_lo(v.y) * scale, bf16_hi(v.y) * scale);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            consumers_sync();
            if (ct == 0) tma_store_1d(dst + t * kTileElems, out, kTileBytes);
        }
        if (ct == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// 3 CTAs (64 KB of ring each) per SM; the CTA count is chosen so that every CTA gets the same number of tiles (+-1):
// "min(tiles, cap)" would leave e.g. 967 tiles on 444 CTAs as 2-or-3 tiles each, i.e. a third of the machine idle at the end.
static int tma_grid(size_t n_tiles) {
    const size_t cap = (size_t)sm_count() * 3;
    if (n_tiles <= cap) return (int)n_tiles;
    const size_t per_cta = (n_tiles + cap - 1) / cap;
    return (int)((n_tiles + per_cta - 1) / per_cta);
}

}  // namespace dmlb

using namespace dmlb;

extern "C" {

int dmlb_bucket_pack_f32_bf16_tma(const float *src, uint16_t *dst, size_t n, float scale, void *stream) {
    if ((!src || !dst) && n) return DMLB_EINVAL;
    if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return DMLB_EALIGN;  // bulk copies need 16-byte alignment
    if (n == 0) return DMLB_OK;
    const size_t n_tiles = n / kTileElems;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tiles) {
        static bool configured[64] = {false};  // the attribute is per device (context)
        const int smem = kTmaStages * kTileBytes;
        int dev = 0;
        DMLB_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            DMLB_CUDA(cudaFuncSetAttribute(pack_bf16_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
        const int grid = tma_grid(n_tiles);
        pack_bf16_tma_kernel<<<grid, kTmaThreads, smem, st>>>(src, dst, n_tiles, scale);
        int rc = launched();
        if (rc != DMLB_OK) return rc;
    }
    const size_t done = n_tiles * kTileElems;
    if (done < n)  // ragged end (< one tile): the register-path kernel
        return dmlb_bucket_pack_f32_bf16_regs(src + done, dst + done, n - done, scale, stream);
    return DMLB_OK;
}

int dmlb_bucket_unpack_bf16_f32_tma(const uint16_t *src, float *dst, size_t n, float scale, void *stream) {
    if ((!src || !dst) && n) return DMLB_EINVAL;
    if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return DMLB_EALIGN;
    if (n == 0) return DMLB_OK;
    const size_t n_tiles = n / kTileElems;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tiles) {
        static bool configured[64] = {false};
        const int smem = kTmaStages * kTileBytesBf16 + 2 * kTileBytes;
        int dev = 0;
        DMLB_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            DMLB_CUDA(cudaFuncSetAttribute(unpack_bf16_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
        const int grid = tma_grid(n_tiles);
        unpack_bf16_tma_kernel<<<grid, kTmaThreads, smem, st>>>(src, dst, n_tiles, scale);
        int rc = launched();
        if (rc != DMLB_OK) return rc;
    }
    const size_t done = n_tiles * kTileElems;
    if (done < n) return dmlb_bucket_unpack_bf16_f32_regs(src + done, dst + done, n - done, scale, nullptr, stream);
    return DMLB_OK;
}

}  // extern "C"
