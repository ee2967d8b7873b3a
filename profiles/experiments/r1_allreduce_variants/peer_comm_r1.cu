// Fused gradient-bucket all-reduce over NVLink 5 / NVSwitch peer memory (sm_100a).
//
// Replaces, for one DDP bucket, the chain the reference runs through torch (pipeline.py:74 -> Reducer -> c10d):
//     bucket * (1/W)  [-> bf16]   ->   allreduce(SUM)   ->   [bf16 ->] fp32 copy back into .grad
// with ONE kernel: scale+cast into this rank's staging half (K1), flag barrier through peer-mapped memory, rank-ordered
// fp32 sum over every rank's staging read across NVLink (the collective), write-back into the fp32 bucket (K2), plus
// an optional fused sum of squares for gradient clipping.  No NCCL, no host round trip, CUDA-graph capturable (the
// sequence number lives in device memory).
//
//   one-shot  (message <= oneshot_max):  every rank reads all W staging buffers  — (W-1)*M bytes over NVLink per GPU,
//                                        one barrier; latency-optimal for the 41 KB MNIST bucket.
//   two-shot  (larger):                  reduce-scatter then all-gather through peer memory — 2*(W-1)/W*M bytes per
//                                        GPU, two barriers; bandwidth-optimal for ResNet-18's 1.96/27.5/15.1 MiB buckets.
//
// Numerics: fp32 accumulate in rank order 0..W-1 on every rank => results are bit-identical across ranks and equal to
// oracle/grad_oracle.py allreduce_f32 / allreduce_bf16.  Two-shot with the bf16 wire rounds the sum to bf16 for the
// all-gather phase (same as an NCCL bf16 all-reduce); the fp32 wire is exact in both algorithms.
#include <cstdlib>
#include <new>

#include "peer_comm.cuh"

namespace dmlb {

template <int kWire>
struct Wire;

template <>
struct Wire<DMLB_WIRE_F32> {  // 4 elements per 16-byte wire vector
    static constexpr int kElems = 4;
    __device__ static __forceinline__ uint4 pack(const float *v) {
        uint4 o;
        o.x = __float_as_uint(v[0]), o.y = __float_as_uint(v[1]), o.z = __float_as_uint(v[2]), o.w = __float_as_uint(v[3]);
        return o;
    }
    __device__ static __forceinline__ void accumulate(float *acc, uint4 w) {
        acc[0] += __uint_as_float(w.x), acc[1] += __uint_as_float(w.y);
        acc[2] += __uint_as_float(w.z), acc[3] += __uint_as_float(w.w);
    }
};

template <>
struct Wire<DMLB_WIRE_BF16> {  // 8 elements per 16-byte wire vector
    static constexpr int kElems = 8;
    __device__ static __forceinline__ uint4 pack(const float *v) {
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]), o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]), o.w = pack_bf16x2(v[6], v[7]);
        return o;
    }
    __device__ static __forceinline__ void accumulate(float *acc, uint4 w) {
        acc[0] += bf16_lo(w.x), acc[1] += bf16_hi(w.x), acc[2] += bf16_lo(w.y), acc[3] += bf16_hi(w.y);
        acc[4] += bf16_lo(w.z), acc[5] += bf16_hi(w.z), acc[6] += bf16_lo(w.w), acc[7] += bf16_hi(w.w);
    }
};

// load kElems fp32 bucket elements of wire vector g (guarded at the ragged end), scaled
template <int E>
__device__ __forceinline__ void load_bucket(const float *bucket, size_t g, size_t n, float scale, float *v) {
    const size_t e0 = g * E;
    if (e0 + E <= n) {
#pragma unroll
        for (int j = 0; j < E; j += 4) {
            float4 t = *reinterpret_cast<const float4 *>(bucket + e0 + j);
            v[j] = t.x * scale, v[j + 1] = t.y * scale, v[j + 2] = t.z * scale, v[j + 3] = t.w * scale;
        }
    } else {
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = (e0 + j < n) ? bucket[e0 + j] * scale : 0.0f;
    }
}

template <int E>
__device__ __forceinline__ double store_bucket(float *bucket, size_t g, size_t n, const float *v, bool sumsq) {
    const size_t e0 = g * E;
    double p = 0.0;
    if (e0 + E <= n) {
#pragma unroll
        for (int j = 0; j < E; j += 4)
            *reinterpret_cast<float4 *>(bucket + e0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (sumsq) {
#pragma unroll
            for (int j = 0; j < E; ++j) p += (double)v[j] * v[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < E; ++j)
            if (e0 + j < n) {
                bucket[e0 + j] = v[j];
                if (sumsq) p += (double)v[j] * v[j];
            }
    }
    return p;
}

// ---------------------------------------------------------------------------------------------------------------------
// Memory-level parallelism.  A peer load over NVLink takes ~2-3 us; to keep 770 GB/s busy ~2 MB must be in flight per
// GPU.  With <= 296 x 256 threads that means several independent 16-byte loads per thread: every loop below gathers
// kU vectors x W ranks into registers before the first add (kU = 4 for W <= 2, 2 for W <= 4, 1 for W <= 8 keeps the
// register budget at ~32 data registers).
// ---------------------------------------------------------------------------------------------------------------------
template <int kWire, int kU, int kStride = kCommThreads>
__device__ __forceinline__ double pack_range(const CommDev &c, const float *bucket, uint4 *mine, size_t lo, size_t hi,
                                             size_t n, float scale, int tid = threadIdx.x) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    for (size_t g0 = lo + tid; g0 < hi; g0 += (size_t)kStride * kU) {
        float v[kU][E];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t g = g0 + (size_t)u * kStride;
            if (g < hi) load_bucket<E>(bucket, g, n, scale, v[u]);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t g = g0 + (size_t)u * kStride;
            if (g < hi) mine[g] = W::pack(v[u]);
        }
    }
    return 0.0;
}

// out(g) = sum over ranks of stage[r][g] for g in [lo, hi) (index space of the staging buffers, offset `goff`)
template <int kWire, int kU, int kStride = kCommThreads, class Sink>
__device__ __forceinline__ void reduce_range(const CommDev &c, int half, size_t lo, size_t hi, size_t goff, Sink sink,
                                             int tid = threadIdx.x) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    constexpr int kMaxW = DMLB_MAX_WORLD / kU;  // the host picks kU so that world <= kMaxW: kU x kMaxW = 8 vectors in flight
    for (size_t i0 = lo + tid; i0 < hi; i0 += (size_t)kStride * kU) {
        uint4 w[kU][kMaxW];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t i = i0 + (size_t)u * kStride;
            if (i < hi) {
#pragma unroll
                for (int r = 0; r < kMaxW; ++r)
                    if (r < c.world) w[u][r] = ld_coherent_u4(reinterpret_cast<const uint4 *>(c.stage(r, half)) + goff + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t i = i0 + (size_t)u * kStride;
            if (i < hi) {
                float acc[E];
#pragma unroll
                for (int j = 0; j < E; ++j) acc[j] = 0.0f;
#pragma unroll
                for (int r = 0; r < kMaxW; ++r)
                    if (r < c.world) W::accumulate(acc, w[u][r]);
                sink(i, acc);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// one-shot
// ---------------------------------------------------------------------------------------------------------------------
template <int kWire, int kU>
__global__ void __launch_bounds__(kCommThreads, 2)
allreduce_oneshot_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t nvec, float scale,
                         double *sumsq_out) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    const uint32_t s = comm_begin(c);
    const int half = s & 1;
    const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(nvec, lo + per);

    pack_range<kWire, kU>(c, bucket, reinterpret_cast<uint4 *>(c.stage(c.rank, half)), lo, hi, n, scale);
    comm_barrier(c, 0, s);

    double part = 0.0;
    const bool want_sumsq = sumsq_out != nullptr;
    reduce_range<kWire, kU>(c, half, lo, hi, 0, [&](size_t g, const float *acc) {
        part += store_bucket<E>(bucket, g, n, acc, want_sumsq);
    });
    if (sumsq_out) {
        double tot = block_sum(part);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sumsq_out, tot);
    }
    comm_end(c, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// one-shot, tile-pipelined: the CTA's range is cut into chunks; warps 0-3 only pack (HBM: read fp32, write wire dtype),
// warps 4-7 only reduce (NVLink: read every rank's staging chunk, sum, write fp32).  Chunk k+1 is being packed while the
// peers' chunk k crosses NVLink, so the HBM pass and the NVLink pass overlap INSIDE the kernel instead of running as two
// phases.  Per-chunk flags live in flag region 2 with values (s << 8) | (k + 1): monotonic across collectives, so the
// ">= target" test of the non-pipelined kernels carries over; the staging double buffer gives the same WAR guarantee.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRole = 128;  // threads per role
__device__ __forceinline__ void role_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kRole) : "memory"); }

template <int kWire, int kU>
__global__ void __launch_bounds__(2 * kRole, 2)
allreduce_oneshot_pipelined_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t nvec,
                                   size_t chunk, float scale, double *sumsq_out) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    const uint32_t s = comm_begin(c);
    const int half = s & 1;
    const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(nvec, lo + per);
    const int n_chunks = hi > lo ? (int)((hi - lo + chunk - 1) / chunk) : 0;
    const uint32_t base = s << 8;
    const bool reducer = threadIdx.x >= kRole;
    const int tid = threadIdx.x & (kRole - 1);

    if (!reducer) {
        uint4 *mine = reinterpret_cast<uint4 *>(c.stage(c.rank, half));
        for (int k = 0; k < n_chunks; ++k) {
            const size_t clo = lo + (size_t)k * chunk, chi = min(hi, clo + chunk);
            pack_range<kWire, kU, kRole>(c, bucket, mine, clo, chi, n, scale, tid);
            __threadfence_system();
            role_sync(1);
            if (tid < c.world) st_release_sys(c.flags(tid, 2, blockIdx.x) + c.rank, base + (uint32_t)k + 1u);
        }
    } else {
        double part = 0.0;
        const bool want_sumsq = sumsq_out != nullptr;
        for (int k = 0; k < n_chunks; ++k) {
            const size_t clo = lo + (size_t)k * chunk, chi = min(hi, clo + chunk);
            if (tid < c.world) {
                const uint32_t *mine = c.flags(c.rank, 2, blockIdx.x) + tid;
                const uint32_t target = base + (uint32_t)k + 1u;
                const unsigned long long t0 = globaltimer_ns();
                while ((int32_t)(ld_acquire_sys(mine) - target) < 0) {
                    if (globaltimer_ns() - t0 > c.timeout_ns) {
                        atomicExch(c.err(), 1u);
                        break;
                    }
                }
            }
            role_sync(2);
            reduce_range<kWire, kU, kRole>(c, half, clo, chi, 0, [&](size_t g, const float *acc) {
                part += store_bucket<E>(bucket, g, n, acc, want_sumsq);
            }, tid);
        }
        if (want_sumsq) {  // reduce over the 4 reducer warps only
            __shared__ double s_red[kRole / 32];
            part = warp_sum(part);
            if ((tid & 31) == 0) s_red[tid >> 5] = part;
            role_sync(2);
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < kRole / 32; ++w) tot += s_red[w];
                if (tot != 0.0) atomicAdd(sumsq_out, tot);
            }
        }
    }
    comm_end(c, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// two-shot: slice q (S wire vectors) is reduced by rank q.  CTA b owns vector range [b*per, (b+1)*per) of EVERY slice,
// so it only ever depends on what the peers' CTA b wrote (per-CTA barriers suffice).
// ---------------------------------------------------------------------------------------------------------------------
// kPush: the all-gather half is PUSHED — the rank that reduced a slice stores it into every rank's result half (posted
// NVLink writes that overlap the reduce-scatter's pulls, which use the other link direction), and after the second
// barrier every rank widens from its LOCAL copy at HBM speed.  !kPush: peers pull the slices after the second barrier.
template <int kWire, int kU, bool kPush>
__global__ void __launch_bounds__(kCommThreads, 2)
allreduce_twoshot_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t nvec, size_t S, float scale,
                         double *sumsq_out) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    const uint32_t s = comm_begin(c);
    const int half = s & 1;
    const size_t per = (S + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(S, lo + per);

    // phase 1 (K1): scale + cast my whole bucket into my staging half, slice by slice
    uint4 *mine = reinterpret_cast<uint4 *>(c.stage(c.rank, half));
    for (int q = 0; q < c.world; ++q) {
        const size_t off = (size_t)q * S;
        if (off >= nvec) break;
        pack_range<kWire, kU>(c, bucket, mine, off + lo, min(off + hi, nvec), n, scale);
    }
    comm_barrier(c, 0, s);

    // phase 2 (reduce-scatter): I reduce slice `rank` from every peer's staging — into my result half (pull variant:
    // slice-local index) or into every rank's result half (push variant: global vector index)
    constexpr int kMaxW = DMLB_MAX_WORLD / kU;
    {
        uint4 *res = reinterpret_cast<uint4 *>(c.result(c.rank, half));
        const size_t off = (size_t)c.rank * S;
        const size_t lim = off < nvec ? min(hi, nvec - off) : 0;
        if (lo < lim) {
            if (kPush)
                reduce_range<kWire, kU>(c, half, lo, lim, off, [&](size_t i, const float *acc) {
                    const uint4 v = W::pack(acc);
#pragma unroll
                    for (int r = 0; r < kMaxW; ++r)
                        if (r < c.world) reinterpret_cast<uint4 *>(c.result(r, half))[off + i] = v;
                });
            else
                reduce_range<kWire, kU>(c, half, lo, lim, off, [&](size_t i, const float *acc) { res[i] = W::pack(acc); });
        }
    }
    comm_barrier(c, 1, s);

    // phase 3 (all-gather + K2): W loads in flight per thread — from every rank's reduced slice over NVLink (pull) or
    // from this rank's own, already complete, result half (push) — widened into the bucket
    double part = 0.0;
    const bool want_sumsq = sumsq_out != nullptr;
    for (size_t i = lo + threadIdx.x; i < hi; i += kCommThreads) {
        uint4 w[kMaxW];
#pragma unroll
        for (int q = 0; q < kMaxW; ++q)
            if (q < c.world && (size_t)q * S + i < nvec)
                w[q] = kPush ? ld_coherent_u4(reinterpret_cast<const uint4 *>(c.result(c.rank, half)) + (size_t)q * S + i)
                             : ld_coherent_u4(reinterpret_cast<const uint4 *>(c.result(q, half)) + i);
#pragma unroll
        for (int q = 0; q < kMaxW; ++q) {
            const size_t g = (size_t)q * S + i;
            if (q < c.world && g < nvec) {
                float acc[E];
#pragma unroll
                for (int j = 0; j < E; ++j) acc[j] = 0.0f;
                W::accumulate(acc, w[q]);
                part += store_bucket<E>(bucket, g, n, acc, want_sumsq);
            }
        }
    }
    if (sumsq_out) {
        double tot = block_sum(part);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sumsq_out, tot);
    }
    comm_end(c, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// two-shot, PUSH-pipelined (algo 5).  Every NVLink transfer is a posted store, every load is local:
//
//   A(t)  pack chunk t of every slice q and store it into rank q's staging half at [my rank][i]      (scatter push)
//   B(t)  sum the W contributions of my slice's chunk t from my LOCAL staging half (rank order), cast,
//         store the reduced vectors into every rank's result half at the global index                 (gather push)
//   C(t)  widen chunk t of every slice from my LOCAL result half into the fp32 bucket (+ sum of squares)
//
// A thread never waits for an NVLink round trip: peer loads (the latency x parallelism limit of the pull kernels) are
// gone, and the links stay busy while the HBM passes run.  The six worker warps of a CTA run A(t), B(t-1), C(t-2)
// back to back; two control warps (one for stage A, one for stage B) do all the signalling so that the system-scope
// fence before a flag store (which waits for the pushes to be acknowledged) never stalls a worker:
//   workers  -> control : shared-memory arrival counter per stage (release: __syncwarp + fence.cta + atomicAdd)
//   control  -> peers   : fence.sys + st.release.sys of (s << 8 | t + 1) into flag region 2 (A) / 3 (B), per CTA
//   peers    -> control : ld.acquire.sys polling of this CTA's own flag words
//   control  -> workers : shared-memory "chunks ready" counter per stage
// WAR safety is the double buffer again: a peer can only be in collective s+1 (writing the other half of my arena)
// once all my CTAs have finished stages A and B of s, and it cannot finish s+1 before I have taken part in it.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kCtrlWarps = 2;
constexpr int kWorkerWarps = kCommThreads / 32 - kCtrlWarps;
constexpr int kWorkers = kWorkerWarps * 32;

struct PushShared {
    uint32_t done[2];   // worker warps that finished stage X of (done / kWorkerWarps) chunks
    uint32_t ready[2];  // chunks of stage X whose data from every rank has landed in this rank's arena
};

__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }

template <int kWire, int kU>
__global__ void __launch_bounds__(kCommThreads, 2)
allreduce_push_pipelined_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t nvec, size_t S,
                                size_t chunk, float scale, double *sumsq_out) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    constexpr int kMaxW = DMLB_MAX_WORLD / kU;
    constexpr int kUA = 16 / E;  // wire vectors per worker thread per iteration of stages A and C: four 128-bit bucket loads in flight
    __shared__ PushShared sh;
    __shared__ double s_red[kWorkerWarps];
    if (threadIdx.x == 0) sh.done[0] = sh.done[1] = sh.ready[0] = sh.ready[1] = 0u;
    const uint32_t s = comm_begin(c);  // (contains the __syncthreads that publishes the zeroed counters)
    const int half = s & 1;
    const size_t per = (S + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(S, lo + per);
    const int K = hi > lo ? (int)((hi - lo + chunk - 1) / chunk) : 0;
    const uint32_t base = s << 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double part = 0.0;

    if (warp < kCtrlWarps) {
        // ---- control warp of stage X: lane r talks to rank r ----
        const int X = warp;
        const int region = 2 + X;
        int sig = 0, rdy = 0;
        const unsigned long long t0 = globaltimer_ns();
        while (sig < K || rdy < K) {
            if (sig < K) {
                uint32_t d = 0;
                if (lane == 0) {
                    d = ld_volatile_shared(&sh.done[X]);
                    __threadfence_block();
                }
                d = __shfl_sync(0xffffffffu, d, 0);
                __syncwarp();
                if (d >= (uint32_t)kWorkerWarps * (uint32_t)(sig + 1)) {  // every worker warp has issued chunk `sig`
                    if (lane < c.world) {
                        __threadfence_system();  // cumulative: the workers' pushes are ordered before the flag
                        st_release_sys(c.flags(lane, region, blockIdx.x) + c.rank, base + (uint32_t)sig + 1u);
                    }
                    ++sig;
                }
            }
            if (rdy < K) {
                bool ok = true;
                if (lane < c.world)
                    ok = (int32_t)(ld_acquire_sys(c.flags(c.rank, region, blockIdx.x) + lane) - (base + (uint32_t)rdy + 1u)) >= 0;
                if (__all_sync(0xffffffffu, ok)) {
                    __syncwarp();
                    ++rdy;
                    if (lane == 0) {
                        __threadfence_block();
                        *reinterpret_cast<volatile uint32_t *>(&sh.ready[X]) = (uint32_t)rdy;
                    }
                }
            }
            if (__any_sync(0xffffffffu, globaltimer_ns() - t0 > c.timeout_ns)) {  // a peer died: record it, release the workers, stop
                if (lane == 0) {
                    atomicExch(c.err(), 1u);
                    *reinterpret_cast<volatile uint32_t *>(&sh.ready[X]) = (uint32_t)K;
                }
                break;
            }
        }
    } else {
        const int w = threadIdx.x - kCtrlWarps * 32;
        const bool want_sumsq = sumsq_out != nullptr;
        const uint4 *my_stage = reinterpret_cast<const uint4 *>(c.stage(c.rank, half));
        const uint4 *my_result = reinterpret_cast<const uint4 *>(c.result(c.rank, half));
        auto arrive = [&](int X) {
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&sh.done[X], 1u);
            }
        };
        auto await = [&](int X, int chunks) {
            if (lane == 0)
                while (ld_volatile_shared(&sh.ready[X]) < (uint32_t)chunks) {}
            __syncwarp();
            __threadfence_block();
        };
        for (int t = 0; t < K + 2; ++t) {
            if (t < K) {  // ---- A(t): scale + cast, scatter to the slice owners ----
                const size_t clo = lo + (size_t)t * chunk;
                const uint32_t cw = (uint32_t)(min(hi, clo + chunk) - clo);
                const uint32_t items = cw * (uint32_t)c.world;
                for (uint32_t j0 = w; j0 < items; j0 += kWorkers * kUA) {
                    float v[kUA][E];
                    size_t idx[kUA];
                    int owner[kUA];
#pragma unroll
                    for (int u = 0; u < kUA; ++u) {
                        const uint32_t j = j0 + u * kWorkers;
                        owner[u] = -1;
                        if (j < items) {
                            const uint32_t q = j / cw;
                            const size_t i = clo + (j - q * cw);
                            const size_t g = (size_t)q * S + i;
                            if (g < nvec) {
                                owner[u] = (int)q;
                                idx[u] = (size_t)c.rank * S + i;
                                load_bucket<E>(bucket, g, n, scale, v[u]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kUA; ++u)
                        if (owner[u] >= 0) reinterpret_cast<uint4 *>(c.stage(owner[u], half))[idx[u]] = W::pack(v[u]);
                }
                arrive(0);
            }
            if (t >= 1 && t - 1 < K) {  // ---- B(t-1): reduce my slice's chunk locally, push it to everyone ----
                await(0, t);
                const size_t clo = lo + (size_t)(t - 1) * chunk;
                const size_t chi = min(hi, clo + chunk);
                const size_t off = (size_t)c.rank * S;
                const size_t lim = off < nvec ? min(chi, nvec - off) : 0;
                for (size_t i0 = clo + w; i0 < lim; i0 += (size_t)kWorkers * kU) {
                    uint4 x[kU][kMaxW];
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const size_t i = i0 + (size_t)u * kWorkers;
                        if (i < lim) {
#pragma unroll
                            for (int r = 0; r < kMaxW; ++r)
                                if (r < c.world) x[u][r] = ld_coherent_u4(my_stage + (size_t)r * S + i);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const size_t i = i0 + (size_t)u * kWorkers;
                        if (i < lim) {
                            float acc[E];
#pragma unroll
                            for (int j = 0; j < E; ++j) acc[j] = 0.0f;
#pragma unroll
                            for (int r = 0; r < kMaxW; ++r)
                                if (r < c.world) W::accumulate(acc, x[u][r]);
                            const uint4 red = W::pack(acc);
#pragma unroll
                            for (int r = 0; r < kMaxW; ++r)
                                if (r < c.world) reinterpret_cast<uint4 *>(c.result(r, half))[off + i] = red;
                        }
                    }
                }
                arrive(1);
            }
            if (t >= 2) {  // ---- C(t-2): widen every slice's chunk from my local result half ----
                await(1, t - 1);
                const size_t clo = lo + (size_t)(t - 2) * chunk;
                const uint32_t cw = (uint32_t)(min(hi, clo + chunk) - clo);
                const uint32_t items = cw * (uint32_t)c.world;
                for (uint32_t j0 = w; j0 < items; j0 += kWorkers * kUA) {
                    uint4 x[kUA];
                    size_t gi[kUA];
#pragma unroll
                    for (int u = 0; u < kUA; ++u) {
                        const uint32_t j = j0 + u * kWorkers;
                        gi[u] = nvec;
                        if (j < items) {
                            const uint32_t q = j / cw;
                            const size_t g = (size_t)q * S + clo + (j - q * cw);
                            if (g < nvec) {
                                gi[u] = g;
                                x[u] = ld_coherent_u4(my_result + g);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kUA; ++u)
                        if (gi[u] < nvec) {
                            float acc[E];
#pragma unroll
                            for (int j = 0; j < E; ++j) acc[j] = 0.0f;
                            W::accumulate(acc, x[u]);
                            part += store_bucket<E>(bucket, gi[u], n, acc, want_sumsq);
                        }
                }
            }
        }
        if (want_sumsq) {
            part = warp_sum(part);
            if (lane == 0) s_red[warp - kCtrlWarps] = part;
        }
    }
    __syncthreads();
    if (sumsq_out && threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < kWorkerWarps; ++i) tot += s_red[i];
        if (tot != 0.0) atomicAdd(sumsq_out, tot);
    }
    comm_end(c, s);
}

__global__ void __launch_bounds__(kCommThreads) barrier_kernel(const __grid_constant__ CommDev c) {
    const uint32_t s = comm_begin(c);
    comm_barrier(c, 0, s);
    comm_end(c, s);
}

constexpr size_t kOneshotMaxBytes = 512 * 1024;
constexpr size_t kPipelineMinBytes = 1 << 20;  // below ~1 MB a single pack/barrier/reduce round is already latency-bound
// Measured on 2x B200 (profiles/r1_comm_sweep_n2_v3_pipelined.json): the warp-specialised pipeline LOSES to the phase-serial
// kernel (98.8 vs 65.5 us at 23 MB bf16): with half the threads per role there are half as many peer loads in flight, and
// the NVLink phase is latency x parallelism bound.  Kept as opt-in algo 3 (bit-exact, tested); not the default.
constexpr bool kPipelineDefault = false;
// two-shot all-gather half: pushed by the reducing rank (algo 4) or pulled by the consumers (algo 2)
constexpr bool kPushDefault = false;

// vectors (all slices together) one CTA moves per pipeline step of algo 5; DMLB_PUSH_STEP_VECTORS overrides it for tuning
static size_t push_step_vectors() {
    static size_t value = [] {
        const char *e = getenv("DMLB_PUSH_STEP_VECTORS");
        long v = e ? atol(e) : 0;
        return (size_t)(v >= 8 ? v : 1536);
    }();
    return value;
}

}  // namespace dmlb

using namespace dmlb;

extern "C" {

size_t dmlb_comm_arena_bytes(size_t max_message_bytes) {
    size_t m = (max_message_bytes + 255) & ~(size_t)255;
    return kHeaderBytes + 4 * m;
}

int dmlb_comm_create(void **comm, int world, int rank, void *const *arenas, size_t max_message_bytes) {
    if (!comm || !arenas || world < 1 || world > DMLB_MAX_WORLD || rank < 0 || rank >= world) return DMLB_EINVAL;
    Comm *c = new (std::nothrow) Comm();
    if (!c) return DMLB_EINVAL;
    c->dev.world = world;
    c->dev.rank = rank;
    c->dev.msg_cap = (max_message_bytes + 255) & ~(size_t)255;
    c->dev.timeout_ns = 10ull * 1000 * 1000 * 1000;
    for (int r = 0; r < DMLB_MAX_WORLD; ++r) c->dev.arena[r] = r < world ? (unsigned char *)arenas[r] : nullptr;
    for (int r = 0; r < world; ++r)
        if (!c->dev.arena[r] || ((uintptr_t)c->dev.arena[r] & 255)) {
            delete c;
            return DMLB_EALIGN;
        }
    *comm = c;
    return DMLB_OK;
}

int dmlb_comm_destroy(void *comm) {
    delete reinterpret_cast<Comm *>(comm);
    return DMLB_OK;
}

int dmlb_comm_allreduce(void *comm, float *bucket, size_t n, int wire, float scale, double *sumsq, int algo,
                        void *stream) {
    if (!comm || (!bucket && n)) return DMLB_EINVAL;
    if (wire != DMLB_WIRE_F32 && wire != DMLB_WIRE_BF16) return DMLB_EINVAL;
    if ((uintptr_t)bucket & 15) return DMLB_EALIGN;
    if (n == 0) return DMLB_OK;
    Comm *c = reinterpret_cast<Comm *>(comm);
    const int E = wire == DMLB_WIRE_BF16 ? 8 : 4;
    const size_t nvec = (n + E - 1) / E;
    const size_t bytes = nvec * 16;
    if (bytes > c->dev.msg_cap) return DMLB_ECAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    const bool pipelined = algo == 3 || (algo == 0 && kPipelineDefault && c->dev.world <= 2 && bytes >= kPipelineMinBytes);
    if (pipelined) {
        const int W = c->dev.world;
        const size_t chunk = 1024;  // wire vectors per chunk per CTA (16 KB)
        size_t want = (nvec + 4 * chunk - 1) / (4 * chunk);  // ~4 chunks per CTA before spreading wider
        size_t cap = (size_t)min(kMaxCtas, sm_count() * 2);
        if (want > cap) want = cap;
        const int grid = (int)(want < 1 ? 1 : want);
        const size_t per = (nvec + grid - 1) / grid;
        if ((per + chunk - 1) / chunk > 250) return DMLB_ECAPACITY;  // chunk index must fit the flag's low 8 bits
#define DMLB_LAUNCH_PIPE(WIRE, U) \
    allreduce_oneshot_pipelined_kernel<WIRE, U><<<grid, 2 * kRole, 0, st>>>(c->dev, bucket, n, nvec, chunk, scale, sumsq)
        if (wire == DMLB_WIRE_BF16) {
            if (W <= 2) DMLB_LAUNCH_PIPE(DMLB_WIRE_BF16, 4);
            else if (W <= 4) DMLB_LAUNCH_PIPE(DMLB_WIRE_BF16, 2);
            else DMLB_LAUNCH_PIPE(DMLB_WIRE_BF16, 1);
        } else {
            if (W <= 2) DMLB_LAUNCH_PIPE(DMLB_WIRE_F32, 4);
            else if (W <= 4) DMLB_LAUNCH_PIPE(DMLB_WIRE_F32, 2);
            else DMLB_LAUNCH_PIPE(DMLB_WIRE_F32, 1);
        }
#undef DMLB_LAUNCH_PIPE
        return launched();
    }
    if (algo == 5) {
        const int W = c->dev.world;
        const size_t S = (nvec + W - 1) / W;
        if ((size_t)W * S * 16 > c->dev.msg_cap) return DMLB_ECAPACITY;  // the owner's staging half holds W x S vectors
        size_t chunk = push_step_vectors() / W;  // vectors of ONE slice per pipeline step
        if (chunk < 1) chunk = 1;
        size_t cap = (size_t)min(kMaxCtas, sm_count() * 2);
        size_t want = (S + chunk - 1) / chunk;
        if (want > cap) want = cap;
        const int grid = (int)(want < 1 ? 1 : want);
        const size_t per = (S + grid - 1) / grid;
        if ((per + chunk - 1) / chunk > 250) chunk = (per + 249) / 250;  // chunk index must fit the flag's low 8 bits
#define DMLB_LAUNCH_PUSH(WIRE, U) \
    allreduce_push_pipelined_kernel<WIRE, U><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, S, chunk, scale, sumsq)
        if (wire == DMLB_WIRE_BF16) {
            if (W <= 2) DMLB_LAUNCH_PUSH(DMLB_WIRE_BF16, 4);
            else if (W <= 4) DMLB_LAUNCH_PUSH(DMLB_WIRE_BF16, 2);
            else DMLB_LAUNCH_PUSH(DMLB_WIRE_BF16, 1);
        } else {
            if (W <= 2) DMLB_LAUNCH_PUSH(DMLB_WIRE_F32, 4);
            else if (W <= 4) DMLB_LAUNCH_PUSH(DMLB_WIRE_F32, 2);
            else DMLB_LAUNCH_PUSH(DMLB_WIRE_F32, 1);
        }
#undef DMLB_LAUNCH_PUSH
        return launched();
    }
    const bool oneshot = algo == 1 || (algo == 0 && (bytes <= kOneshotMaxBytes || c->dev.world <= 2));
    const bool push = algo == 4 || (algo == 0 && kPushDefault);
    const int W = c->dev.world;
    const int kU = W <= 2 ? 4 : (W <= 4 ? 2 : 1);
    const size_t items = oneshot ? nvec : (nvec + W - 1) / W;  // vectors a CTA grid is spread over
    size_t want = (items + (size_t)kCommThreads * kU - 1) / ((size_t)kCommThreads * kU);
    size_t cap = (size_t)min(kMaxCtas, sm_count() * 2);  // all CTAs co-resident: the per-CTA barriers need that
    if (want > cap) want = cap;
    const int grid = (int)(want < 1 ? 1 : want);
#define DMLB_LAUNCH_AR(WIRE, U)                                                                                       \
    do {                                                                                                              \
        if (oneshot)                                                                                                  \
            allreduce_oneshot_kernel<WIRE, U><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, scale, sumsq);  \
        else if (push)                                                                                                \
            allreduce_twoshot_kernel<WIRE, U, true><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, items,   \
                                                                                   scale, sumsq);                     \
        else                                                                                                          \
            allreduce_twoshot_kernel<WIRE, U, false><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, items,  \
                                                                                    scale, sumsq);                    \
    } while (0)
    if (wire == DMLB_WIRE_BF16) {
        if (kU == 4) DMLB_LAUNCH_AR(DMLB_WIRE_BF16, 4);
        else if (kU == 2) DMLB_LAUNCH_AR(DMLB_WIRE_BF16, 2);
        else DMLB_LAUNCH_AR(DMLB_WIRE_BF16, 1);
    } else {
        if (kU == 4) DMLB_LAUNCH_AR(DMLB_WIRE_F32, 4);
        else if (kU == 2) DMLB_LAUNCH_AR(DMLB_WIRE_F32, 2);
        else DMLB_LAUNCH_AR(DMLB_WIRE_F32, 1);
    }
#undef DMLB_LAUNCH_AR
    return launched();
}

int dmlb_comm_error(void *comm, int *error) {
    if (!comm || !error) return DMLB_EINVAL;
    Comm *c = reinterpret_cast<Comm *>(comm);
    uint32_t word = 0;
    // the error word lives in this rank's own arena (control block, word 2); a blocking 4-byte read: call it per epoch
    DMLB_CUDA(cudaMemcpy(&word, c->dev.arena[c->dev.rank] + 2 * sizeof(uint32_t), sizeof(word), cudaMemcpyDeviceToHost));
    *error = (int)word;
    return DMLB_OK;
}

int dmlb_comm_barrier(void *comm, void *stream) {
    if (!comm) return DMLB_EINVAL;
    Comm *c = reinterpret_cast<Comm *>(comm);
    barrier_kernel<<<1, kCommThreads, 0, (cudaStream_t)stream>>>(c->dev);
    return launched();
}

}  // extern "C"
