// Fused gradient-bucket all-reduce + per-step metric exchange over NVLink 5 / NVSwitch peer memory (sm_100a).
//
// Replaces, for one DDP bucket, the chain the reference runs through torch (pipeline.py:74 -> Reducer -> c10d):
//     bucket * (1/W)  [-> bf16]   ->   allreduce(SUM)   ->   [bf16 ->] fp32 copy back into .grad
// with ONE kernel: scale+cast into this rank's staging half (K1), flag barrier through peer-mapped memory, fp32 sum over
// every rank's staging read across NVLink (the collective), write-back into the fp32 bucket (K2), plus an optional
// fused sum of squares for gradient clipping.  No NCCL, no host round trip, CUDA-graph capturable (the sequence number
// lives in device memory).
//
//   one-shot  (message <= oneshot_max):  every rank reads all W staging buffers  — (W-1)*M bytes over NVLink per GPU,
//                                        one barrier; latency-optimal for the 41 KB MNIST bucket.
//   two-shot  (larger):                  reduce-scatter then all-gather through peer loads — 2*(W-1)/W*M bytes per GPU,
//                                        two barriers.
//   NVLS      (larger, multicast bound): multimem.ld_reduce pulls this rank's 1/W slice already SUMMED BY THE SWITCH,
//                                        multimem.st broadcasts it — (1 + 1/W)*M bytes per GPU and direction, two barriers.
//
// The fused STEP EXCHANGE: when a dmlb_step_metrics descriptor is attached, one extra CTA of the same kernel folds the
// step's tracked values into the metric slab, finalises the selected cells, exchanges 16-byte records under the SAME flag
// barrier as the gradients and writes the cross-rank results into a ring in mapped host memory — the reference's
// per-step `track_reduce` traffic (stage.py:305-314) and its cross-rank reduction (metrics.py:121-141) cost no launch
// and no barrier of their own.
//
// Numerics: one-shot / two-shot accumulate in fp32 in rank order 0..W-1 on every rank => results are bit-identical
// across ranks and equal to oracle/grad_oracle.py allreduce_f32 / allreduce_bf16.  Two-shot and NVLS with the bf16 wire
// round the sum to bf16 for the all-gather phase (same as an NCCL bf16 all-reduce).  NVLS sums in the switch (fp32
// accumulation, order fixed by the hardware, identical on all ranks because every rank receives the same broadcast).
#include <new>

#include "metric_dev.cuh"
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
    __device__ static __forceinline__ uint4 mc_reduce(const void *mc) {  // in-switch sum over all ranks' copies
        uint4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(mc)
                     : "memory");
        return v;
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
    __device__ static __forceinline__ uint4 mc_reduce(const void *mc) {  // fp32 accumulation in the switch, bf16 result
        uint4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(mc)
                     : "memory");
        return v;
    }
};

__device__ __forceinline__ void mc_store(void *mc, uint4 v) {  // one store, delivered to every rank's copy
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

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

// A peer did not arrive: overwrite this CTA's part of the bucket with NaN so that nobody trains on a partial sum.
template <int E>
__device__ __forceinline__ void poison_range(float *bucket, size_t lo, size_t hi, size_t n) {
    const float nan = __int_as_float(0x7fc00000);
    for (size_t e = lo * E + threadIdx.x; e < hi * E && e < n; e += kCommThreads) bucket[e] = nan;
}

// ---------------------------------------------------------------------------------------------------------------------
// Memory-level parallelism.  A peer load over NVLink takes ~2-3 us; to keep 770 GB/s busy ~2 MB must be in flight per
// GPU.  With <= 296 x 256 threads that means several independent 16-byte loads per thread: every loop below gathers
// kU vectors x W ranks into registers before the first add (kU = 4 for W <= 2, 2 for W <= 4, 1 for W <= 8 keeps the
// register budget at ~32 data registers).
// ---------------------------------------------------------------------------------------------------------------------
template <int kWire, int kU>
__device__ __forceinline__ void pack_range(const float *bucket, uint4 *mine, size_t lo, size_t hi, size_t n, float scale) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    for (size_t g0 = lo + threadIdx.x; g0 < hi; g0 += (size_t)kCommThreads * kU) {
        float v[kU][E];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t g = g0 + (size_t)u * kCommThreads;
            if (g < hi) load_bucket<E>(bucket, g, n, scale, v[u]);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t g = g0 + (size_t)u * kCommThreads;
            if (g < hi) mine[g] = W::pack(v[u]);
        }
    }
}

// out(g) = sum over ranks of stage[r][g] for g in [lo, hi) (index space of the staging buffers, offset `goff`)
template <int kWire, int kU, class Sink>
__device__ __forceinline__ void reduce_range(const CommDev &c, int half, size_t lo, size_t hi, size_t goff, Sink sink) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    constexpr int kMaxW = DMLB_MAX_WORLD / kU;  // the host picks kU so that world <= kMaxW: kU x kMaxW = 8 vectors in flight
    for (size_t i0 = lo + threadIdx.x; i0 < hi; i0 += (size_t)kCommThreads * kU) {
        uint4 w[kU][kMaxW];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t i = i0 + (size_t)u * kCommThreads;
            if (i < hi) {
#pragma unroll
                for (int r = 0; r < kMaxW; ++r)
                    if (r < c.world) w[u][r] = ld_coherent_u4(reinterpret_cast<const uint4 *>(c.stage(r, half)) + goff + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t i = i0 + (size_t)u * kCommThreads;
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
// The metric CTA of the fused step exchange (block index == number of data CTAs).
// fold -> finalise (no reset) -> records into mstage[half] -> the collective's barrier 0 -> rank-ordered combine -> ring.
// kLL: the records travel as LL lines pushed into every rank's arena (header = lines 0,1; record i = lines 2+2i, 3+2i),
// so the exchange costs one one-way NVLink latency: no barrier, no peer loads (the LL all-reduce kernel's companion).
// ---------------------------------------------------------------------------------------------------------------------
template <bool kLL>
__device__ __noinline__ void metric_cta(const CommDev &c, uint32_t s, const dmlb_step_metrics &M) {
    __shared__ unsigned long long s_count;
    long long *cnt = reinterpret_cast<long long *>(M.cnt);
    if (threadIdx.x == 0) s_count = *reinterpret_cast<volatile unsigned long long *>(M.counter);
    __syncthreads();
    const unsigned long long count = s_count;
    const bool exchange = c.world > 1;
    const int half = s & 1;

    // 1. this step's values: warp w takes entries w, w + 8, ...
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const double *feed_slot = (M.feed && M.feed_slots > 0)
                                      ? M.feed + (size_t)(count % (unsigned long long)M.feed_slots) * 2 * DMLB_FEED_WIDTH
                                      : nullptr;
        for (int j = warp; j < M.n_folds; j += kCommThreads / 32) {
            dmlb_fold_entry e = M.folds[j];
            if (e.src_dtype == DMLB_SRC_FEED) {
                if (!feed_slot) continue;
                e.src = feed_slot;
            }
            fold_entry(M.acc, cnt, M.desc, e, lane, 32);
        }
    }
    __syncthreads();  // entries may target cells the finalisation below reads (same CTA: block-level ordering is enough)

    // 2. finalise
    unsigned char *slot = M.out_ring + (size_t)(count % (unsigned long long)M.ring_slots) *
                                           ((size_t)DMLB_METRIC_STATUS_SLOTS * 4 + 9 * (size_t)M.capacity);
    int *status = reinterpret_cast<int *>(slot);
    uint64_t *out_val = reinterpret_cast<uint64_t *>(slot + DMLB_METRIC_STATUS_SLOTS * 4);
    uint8_t *out_flag = slot + DMLB_METRIC_STATUS_SLOTS * 4 + 8 * (size_t)M.capacity;
    int n_glob = 0, n_all = 0;
    for (int j = 0; j < M.n_ranges; ++j) {
        const int len = M.ranges[j].end - M.ranges[j].begin;
        n_all += len;
        if (j < M.n_global_ranges) n_glob += len;
    }
    const dmlb_range *lr = M.ranges + M.n_global_ranges;
    for (int i = threadIdx.x; i < n_all - n_glob; i += kCommThreads) {
        const int cell = sel_to_cell(lr, M.n_ranges - M.n_global_ranges, i);
        uint64_t val;
        long long n;
        finalize_cell(M.acc, cnt, M.desc[cell], cell, val, n, false);
        out_val[cell] = val;
        out_flag[cell] = n > 0 ? 0 : 1;
    }
    uint64_t *rec_mine = reinterpret_cast<uint64_t *>(c.mstage(c.rank, half));
    for (int i = threadIdx.x; i < n_glob; i += kCommThreads) {
        const int cell = sel_to_cell(M.ranges, M.n_global_ranges, i);
        uint64_t val;
        long long n;
        finalize_cell(M.acc, cnt, M.desc[cell], cell, val, n, false);
        if (exchange && kLL) {
#pragma unroll
            for (int r = 0; r < DMLB_MAX_WORLD; ++r)
                if (r < c.world) {
                    uint4 *dst = c.ll_metric(r, half, c.rank) + 2 + 2 * i;
                    ll_store(dst, (uint32_t)val, (uint32_t)(val >> 32), s);
                    ll_store(dst + 1, (uint32_t)(uint64_t)n, (uint32_t)((uint64_t)n >> 32), s);
                }
        } else if (exchange) {
            rec_mine[2 + 2 * i] = val;
            rec_mine[3 + 2 * i] = (uint64_t)n;
        } else {
            out_val[cell] = val;
            out_flag[cell] = n > 0 ? 0 : 1;
        }
    }
    int st = DMLB_METRIC_OK;
    if (exchange && kLL) {
        if (threadIdx.x < c.world) {  // header lines to rank threadIdx.x
            uint4 *dst = c.ll_metric(threadIdx.x, half, c.rank);
            ll_store(dst, (uint32_t)M.layout_hash, (uint32_t)(M.layout_hash >> 32), s);
            ll_store(dst + 1, (uint32_t)n_glob, 0u, s);
        }
        // 3. every rank's header has to arrive and agree before any record index is trusted
        uint4 w[DMLB_MAX_WORLD];
        bool arrived = true;
        if (threadIdx.x == 0) {
            arrived = ll_wait_all(c, s, [&](int r) { return c.ll_metric(c.rank, half, r); }, w);
            if (arrived) {
                for (int r = 0; r < c.world; ++r)
                    if ((((uint64_t)w[r].z << 32) | w[r].x) != M.layout_hash) st = DMLB_METRIC_LAYOUT;
                arrived = ll_wait_all(c, s, [&](int r) { return c.ll_metric(c.rank, half, r) + 1; }, w);
                for (int r = 0; arrived && r < c.world; ++r)
                    if (w[r].x != (uint32_t)n_glob) st = DMLB_METRIC_LAYOUT;
            }
            if (!arrived) st = DMLB_METRIC_TIMEOUT;
        }
        const bool ok = __syncthreads_or(st != DMLB_METRIC_OK) == 0;
        // 4. combine in rank order (each thread polls the lines of its own cells)
        if (ok)
            for (int i = threadIdx.x; i < n_glob; i += kCommThreads) {
                const int cell = sel_to_cell(M.ranges, M.n_global_ranges, i);
                uint4 wv[DMLB_MAX_WORLD], wn[DMLB_MAX_WORLD];
                const bool got = ll_wait_all(c, s, [&](int r) { return c.ll_metric(c.rank, half, r) + 2 + 2 * i; }, wv) &&
                                 ll_wait_all(c, s, [&](int r) { return c.ll_metric(c.rank, half, r) + 3 + 2 * i; }, wn);
                if (!got) {
                    st = DMLB_METRIC_TIMEOUT;
                    break;
                }
                uint64_t out;
                uint8_t flag;
                auto rec = [&](int r, uint64_t &v, long long &n) {
                    v = ((uint64_t)wv[r].z << 32) | wv[r].x;
                    n = (long long)(((uint64_t)wn[r].z << 32) | wn[r].x);
                };
                combine_cell(M.desc[cell], c.world, rec, out, flag, st);
                out_val[cell] = out;
                out_flag[cell] = flag;
            }
    } else if (exchange) {
        if (threadIdx.x == 0) {
            rec_mine[0] = M.layout_hash;
            rec_mine[1] = (uint64_t)n_glob;
        }
        // 3. the collective's barrier 0 (this CTA owns flag slot blockIdx.x like any data CTA)
        const bool arrived = comm_barrier(c, 0, s);
        if (!arrived) st = DMLB_METRIC_TIMEOUT;
        if (arrived && threadIdx.x < c.world) {
            uint4 h = ld_coherent_u4(reinterpret_cast<const uint4 *>(c.mstage(threadIdx.x, half)));
            const uint64_t ph = ((uint64_t)h.y << 32) | h.x, pn = ((uint64_t)h.w << 32) | h.z;
            if (ph != M.layout_hash || pn != (uint64_t)n_glob) st = DMLB_METRIC_LAYOUT;
        }
        const bool ok = __syncthreads_or(st != DMLB_METRIC_OK) == 0;
        // 4. combine in rank order
        if (ok)
            for (int i = threadIdx.x; i < n_glob; i += kCommThreads) {
                const int cell = sel_to_cell(M.ranges, M.n_global_ranges, i);
                uint64_t out;
                uint8_t flag;
                auto rec = [&](int r, uint64_t &v, long long &n) {
                    uint4 w = ld_coherent_u4(reinterpret_cast<const uint4 *>(c.mstage(r, half)) + 1 + i);
                    v = ((uint64_t)w.y << 32) | w.x;
                    n = (long long)(((uint64_t)w.w << 32) | w.z);
                };
                combine_cell(M.desc[cell], c.world, rec, out, flag, st);
                out_val[cell] = out;
                out_flag[cell] = flag;
            }
    }
    int worst = DMLB_METRIC_OK;
    if (__syncthreads_or(st == DMLB_METRIC_TIMEOUT)) worst = DMLB_METRIC_TIMEOUT;
    else if (__syncthreads_or(st == DMLB_METRIC_LAYOUT)) worst = DMLB_METRIC_LAYOUT;
    else if (__syncthreads_or(st == DMLB_METRIC_SPLIT_VOTE)) worst = DMLB_METRIC_SPLIT_VOTE;
    // 5. publish: results first, then the stamp (the host trusts a slot only when its stamp matches)
    __syncthreads();
    if (threadIdx.x == 0) {
        status[0] = worst;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(slot + DMLB_METRIC_STATUS_SLOTS * 4 - 8) = count + 1ull;
        *reinterpret_cast<volatile unsigned long long *>(M.counter) = count + 1ull;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// one-shot
// ---------------------------------------------------------------------------------------------------------------------
template <int kWire, int kU>
__global__ void __launch_bounds__(kCommThreads, 2)
allreduce_oneshot_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t nvec, float scale,
                         double *sumsq_out, int n_data, const __grid_constant__ dmlb_step_metrics M) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    const uint32_t s = comm_begin(c);
    if ((int)blockIdx.x >= n_data) {
        metric_cta<false>(c, s, M);
        comm_end(c, s);
        return;
    }
    const int half = s & 1;
    const size_t per = (nvec + n_data - 1) / n_data;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(nvec, lo + per);
    double part = 0.0;
    const bool want_sumsq = sumsq_out != nullptr;

    if (c.world == 1) {
        // nobody to exchange with: round through the wire dtype in registers (what pack -> sum over one rank -> unpack
        // computes), no staging, no barrier
        for (size_t g = lo + threadIdx.x; g < hi; g += kCommThreads) {
            float v[E], acc[E];
            load_bucket<E>(bucket, g, n, scale, v);
            const uint4 w = W::pack(v);
#pragma unroll
            for (int j = 0; j < E; ++j) acc[j] = 0.0f;
            W::accumulate(acc, w);
            part += store_bucket<E>(bucket, g, n, acc, want_sumsq);
        }
    } else {
        pack_range<kWire, kU>(bucket, reinterpret_cast<uint4 *>(c.stage(c.rank, half)), lo, hi, n, scale);
        if (comm_barrier(c, 0, s)) {
            reduce_range<kWire, kU>(c, half, lo, hi, 0, [&](size_t g, const float *acc) {
                part += store_bucket<E>(bucket, g, n, acc, want_sumsq);
            });
        } else {
            poison_range<E>(bucket, lo, hi, n);
            part = __longlong_as_double(0x7ff8000000000000ll);
        }
    }
    if (sumsq_out) {
        double tot = block_sum(part);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sumsq_out, tot);
    }
    comm_end(c, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// one-shot, LL protocol (messages up to kLLMaxPayload wire bytes, W > 1): see peer_comm.cuh.  A line carries 8 payload bytes
// = 4 bf16 or 2 fp32 elements.  Every thread pushes its lines to all W ranks (its own included: the pull loop is uniform),
// then polls its own arena's lines of the same indices from all W sources and sums in rank order — bit-identical to the
// barrier one-shot and to the oracle.
// ---------------------------------------------------------------------------------------------------------------------
template <int kWire>
__global__ void __launch_bounds__(kCommThreads, 2)
allreduce_ll_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t n_lines, float scale,
                    double *sumsq_out, int n_data, const __grid_constant__ dmlb_step_metrics M) {
    constexpr int EL = kWire == DMLB_WIRE_BF16 ? 4 : 2;  // elements per line
    const uint32_t s = comm_begin(c);
    if ((int)blockIdx.x >= n_data) {
        metric_cta<true>(c, s, M);
        comm_end(c, s);
        return;
    }
    const int half = s & 1;
    const size_t per = (n_lines + n_data - 1) / n_data;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(n_lines, lo + per);
    // push
    for (size_t l = lo + threadIdx.x; l < hi; l += kCommThreads) {
        float v[EL];
        const size_t e0 = l * EL;
#pragma unroll
        for (int j = 0; j < EL; ++j) v[j] = (e0 + j < n) ? bucket[e0 + j] * scale : 0.0f;
        uint32_t d0, d1;
        if (kWire == DMLB_WIRE_BF16) {
            d0 = pack_bf16x2(v[0], v[1]);
            d1 = pack_bf16x2(v[EL - 2], v[EL - 1]);
        } else {
            d0 = __float_as_uint(v[0]);
            d1 = __float_as_uint(v[1]);
        }
#pragma unroll
        for (int r = 0; r < DMLB_MAX_WORLD; ++r)
            if (r < c.world) ll_store(c.ll(r, half, c.rank) + l, d0, d1, s);
    }
    // pull + sum in rank order
    double part = 0.0;
    const bool want_sumsq = sumsq_out != nullptr;
    bool ok = true;
    for (size_t l = lo + threadIdx.x; l < hi && ok; l += kCommThreads) {
        uint4 w[DMLB_MAX_WORLD];
        ok = ll_wait_all(c, s, [&](int r) { return c.ll(c.rank, half, r) + l; }, w);
        if (!ok) break;
        float acc[EL];
#pragma unroll
        for (int j = 0; j < EL; ++j) acc[j] = 0.0f;
#pragma unroll
        for (int r = 0; r < DMLB_MAX_WORLD; ++r)
            if (r < c.world) {
                if (kWire == DMLB_WIRE_BF16) {
                    acc[0] += bf16_lo(w[r].x), acc[1] += bf16_hi(w[r].x);
                    acc[EL - 2] += bf16_lo(w[r].z), acc[EL - 1] += bf16_hi(w[r].z);
                } else {
                    acc[0] += __uint_as_float(w[r].x), acc[1] += __uint_as_float(w[r].z);
                }
            }
        const size_t e0 = l * EL;
#pragma unroll
        for (int j = 0; j < EL; ++j)
            if (e0 + j < n) {
                bucket[e0 + j] = acc[j];
                if (want_sumsq) part += (double)acc[j] * acc[j];
            }
    }
    if (__syncthreads_or(!ok)) {  // a peer died: nobody trains on a partial sum
        const float nan = __int_as_float(0x7fc00000);
        for (size_t e = lo * EL + threadIdx.x; e < hi * EL && e < n; e += kCommThreads) bucket[e] = nan;
        part = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (sumsq_out) {
        double tot = block_sum(part);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sumsq_out, tot);
    }
    comm_end(c, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// two-shot: slice q (S wire vectors) is reduced by rank q.  CTA b owns vector range [b*per, (b+1)*per) of EVERY slice,
// so it only ever depends on what the peers' CTA b wrote (per-CTA barriers suffice).
// kNvls = 1: the reduce-scatter + all-gather pair is done by the switch — multimem.ld_reduce of my slice from the multicast
// mapping of all staging halves, multimem.st of the sum back into every rank's staging half (in place: between the two
// barriers only the owner touches slice q), then every rank widens its own, now reduced, staging half.
// kNvls = 2: only the reduce-scatter goes through the switch (one request stream per GPU instead of W-1); the reduced slices
// stay in their owner's result half and the all-gather is the peer-load phase of the plain two-shot, fused with K2.
// ---------------------------------------------------------------------------------------------------------------------
template <int kWire, int kU, int kNvls>
__global__ void __launch_bounds__(kCommThreads, 2)
allreduce_twoshot_kernel(const __grid_constant__ CommDev c, float *bucket, size_t n, size_t nvec, size_t S, float scale,
                         double *sumsq_out, int n_data, const __grid_constant__ dmlb_step_metrics M) {
    typedef Wire<kWire> W;
    constexpr int E = W::kElems;
    const uint32_t s = comm_begin(c);
    if ((int)blockIdx.x >= n_data) {
        metric_cta<false>(c, s, M);
        comm_end(c, s);
        return;
    }
    const int half = s & 1;
    const size_t per = (S + n_data - 1) / n_data;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = min(S, lo + per);

    // phase 1 (K1): scale + cast my whole bucket into my staging half, slice by slice
    uint4 *mine = reinterpret_cast<uint4 *>(c.stage(c.rank, half));
    for (int q = 0; q < c.world; ++q) {
        const size_t off = (size_t)q * S;
        if (off >= nvec) break;
        pack_range<kWire, kU>(bucket, mine, off + lo, min(off + hi, nvec), n, scale);
    }
    bool ok = comm_barrier(c, 0, s);

    // phase 2 (reduce-scatter): I reduce slice `rank`
    constexpr int kMaxW = DMLB_MAX_WORLD / kU;
    {
        const size_t off = (size_t)c.rank * S;
        const size_t lim = off < nvec ? min(hi, nvec - off) : 0;
        if (ok && lo < lim) {
            if (kNvls) {
                unsigned char *mcs = c.mc_stage(half) + off * 16;
                uint4 *res = reinterpret_cast<uint4 *>(c.result(c.rank, half));
                constexpr int kV = 4;  // independent in-switch reductions in flight per thread
                for (size_t i0 = lo + threadIdx.x; i0 < lim; i0 += (size_t)kCommThreads * kV) {
                    uint4 v[kV];
#pragma unroll
                    for (int u = 0; u < kV; ++u) {
                        const size_t i = i0 + (size_t)u * kCommThreads;
                        if (i < lim) v[u] = W::mc_reduce(mcs + i * 16);
                    }
#pragma unroll
                    for (int u = 0; u < kV; ++u) {
                        const size_t i = i0 + (size_t)u * kCommThreads;
                        if (i < lim) {
                            if (kNvls == 1) mc_store(mcs + i * 16, v[u]);  // broadcast by the switch into every staging half
                            else res[i] = v[u];                            // kept local: the peers pull it in phase 3
                        }
                    }
                }
            } else {
                uint4 *res = reinterpret_cast<uint4 *>(c.result(c.rank, half));
                reduce_range<kWire, kU>(c, half, lo, lim, off, [&](size_t i, const float *acc) { res[i] = W::pack(acc); });
            }
        }
    }
    ok = comm_barrier(c, 1, s) && ok;

    // phase 3 (all-gather + K2): W loads in flight per thread — from every rank's reduced slice over NVLink (pull), or
    // (NVLS) from my own staging half, which the owners' multicast stores have overwritten with the sums
    double part = 0.0;
    const bool want_sumsq = sumsq_out != nullptr;
    if (ok) {
        for (size_t i = lo + threadIdx.x; i < hi; i += kCommThreads) {
            uint4 w[kMaxW];
#pragma unroll
            for (int q = 0; q < kMaxW; ++q)
                if (q < c.world && (size_t)q * S + i < nvec)
                    w[q] = kNvls == 1 ? ld_coherent_u4(reinterpret_cast<const uint4 *>(mine) + (size_t)q * S + i)
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
    } else {
        for (int q = 0; q < c.world; ++q) poison_range<E>(bucket, (size_t)q * S + lo, min((size_t)q * S + hi, nvec), n);
        part = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (sumsq_out) {
        double tot = block_sum(part);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sumsq_out, tot);
    }
    comm_end(c, s);
}

__global__ void __launch_bounds__(kCommThreads) barrier_kernel(const __grid_constant__ CommDev c) {
    const uint32_t s = comm_begin(c);
    comm_barrier(c, 0, s);
    comm_end(c, s);
}

constexpr size_t kOneshotMaxBytes = 512 * 1024;
// NVLS pays off where it removes traffic: per GPU and direction it moves (1 + 1/W) M instead of 2 (W-1)/W M, i.e. nothing at
// W = 2 and 1.56x less at W = 8; multimem operations also have a longer latency than plain peer loads.  Auto-dispatch uses
// it from W = 8 and 8 MB upward (profiles/r2_nvls_probe_n{2,4,8}.json); algo = 3 forces it wherever multicast is bound.
constexpr size_t kNvlsMinBytes = 8 << 20;
constexpr int kNvlsMinWorld = 8;
static_assert(sizeof(dmlb_step_metrics) == 1880, "dmlb_step_metrics layout (dmlcloud_b200/_native.py StepMetrics mirrors it)");

}  // namespace dmlb

using namespace dmlb;

extern "C" {

size_t dmlb_comm_arena_bytes(size_t max_message_bytes) {
    size_t m = (max_message_bytes + 255) & ~(size_t)255;
    return kHeaderBytes + 4 * m + kLLBytes;
}

int dmlb_comm_create(void **comm, int world, int rank, void *const *arenas, size_t max_message_bytes) {
    if (!comm || !arenas || world < 1 || world > DMLB_MAX_WORLD || rank < 0 || rank >= world) return DMLB_EINVAL;
    Comm *c = new (std::nothrow) Comm();
    if (!c) return DMLB_EINVAL;
    c->dev.world = world;
    c->dev.rank = rank;
    c->dev.msg_cap = (max_message_bytes + 255) & ~(size_t)255;
    c->dev.timeout_ns = 600ull * 1000 * 1000 * 1000;  // 10 minutes, like NCCL's watchdog default
    c->dev.mc = nullptr;
    c->dev.host_err = nullptr;
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

int dmlb_comm_configure(void *comm, double timeout_seconds, uint32_t *host_error_word) {
    if (!comm) return DMLB_EINVAL;
    Comm *c = reinterpret_cast<Comm *>(comm);
    if (timeout_seconds > 0.0) c->dev.timeout_ns = (unsigned long long)(timeout_seconds * 1e9);
    c->dev.host_err = host_error_word;
    return DMLB_OK;
}

int dmlb_comm_set_multicast(void *comm, void *mc_base) {
    if (!comm) return DMLB_EINVAL;
    reinterpret_cast<Comm *>(comm)->dev.mc = reinterpret_cast<unsigned char *>(mc_base);
    return DMLB_OK;
}

int dmlb_comm_allreduce(void *comm, float *bucket, size_t n, int wire, float scale, double *sumsq, int algo,
                        const dmlb_step_metrics *metrics, void *stream) {
    if (!comm || (!bucket && n)) return DMLB_EINVAL;
    if (wire != DMLB_WIRE_F32 && wire != DMLB_WIRE_BF16) return DMLB_EINVAL;
    if ((uintptr_t)bucket & 15) return DMLB_EALIGN;
    if (n == 0 && !metrics) return DMLB_OK;
    Comm *c = reinterpret_cast<Comm *>(comm);
    const int W = c->dev.world;
    const int E = wire == DMLB_WIRE_BF16 ? 8 : 4;
    const size_t nvec = (n + E - 1) / E;
    const size_t bytes = nvec * 16;
    if (W > 1 && bytes > c->dev.msg_cap) return DMLB_ECAPACITY;
    static const dmlb_step_metrics kNoMetrics = {};
    if (metrics) {
        const dmlb_step_metrics &m = *metrics;
        if (!m.acc || !m.cnt || !m.desc || !m.counter || !m.out_ring || m.ring_slots < 1 || m.capacity < 1)
            return DMLB_EINVAL;
        if (m.n_folds < 0 || m.n_folds > DMLB_MAX_FOLD_ENTRIES || m.n_ranges < 0 || m.n_ranges > DMLB_MAX_RANGES ||
            m.n_global_ranges < 0 || m.n_global_ranges > m.n_ranges)
            return DMLB_ECAPACITY;
        long long n_glob = 0;
        for (int j = 0; j < m.n_ranges; ++j) {
            if (m.ranges[j].begin < 0 || m.ranges[j].end < m.ranges[j].begin || m.ranges[j].end > m.n_cells ||
                m.ranges[j].end > m.capacity)
                return DMLB_EINVAL;
            if (j < m.n_global_ranges) n_glob += m.ranges[j].end - m.ranges[j].begin;
        }
        if (n_glob > DMLB_STEP_METRIC_MAX_CELLS) return DMLB_ECAPACITY;
        for (int j = 0; j < m.n_folds; ++j) {
            const dmlb_fold_entry &e = m.folds[j];
            if (e.cell < 0 || e.lanes < 1 || e.k < 0 || e.steps < 1 || e.cell + e.lanes > m.n_cells) return DMLB_EINVAL;
            if (e.src_dtype == DMLB_SRC_FEED) {
                if (e.k >= DMLB_FEED_WIDTH || e.lanes != 1) return DMLB_EINVAL;
            } else if (e.src_dtype < DMLB_F32 || e.src_dtype > DMLB_U8 || e.k < 1) {
                return DMLB_EINVAL;
            }
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool nvls = W > 1 && c->dev.mc != nullptr &&
                      (algo == 3 || algo == 4 || (algo == 0 && W >= kNvlsMinWorld && bytes >= kNvlsMinBytes));
    if ((algo == 3 || algo == 4) && !nvls && W > 1) return DMLB_ESTATE;
    const bool nvls_rs_only = nvls && algo == 4;
    const bool oneshot = !nvls && (W == 1 || algo == 1 || algo == 5 || (algo == 0 && (bytes <= kOneshotMaxBytes || W <= 2)));
    // small messages at W > 1: the LL protocol (no barrier, no peer loads); algo 5 forces the barrier one-shot for A/B runs
    if (oneshot && W > 1 && algo != 5 && bytes <= kLLMaxPayload) {
        const int EL = wire == DMLB_WIRE_BF16 ? 4 : 2;
        const size_t n_lines = (n + EL - 1) / EL;
        size_t want = (n_lines + kCommThreads - 1) / kCommThreads;  // one line per thread while the grid can grow
        const size_t cap = (size_t)min(kMaxCtas, sm_count() * 2) - 1;
        if (want > cap) want = cap;
        const int n_data = n == 0 ? 0 : (int)(want < 1 ? 1 : want);
        const int grid = n_data + (metrics ? 1 : 0);
        const dmlb_step_metrics &Mll = metrics ? *metrics : kNoMetrics;
        if (wire == DMLB_WIRE_BF16)
            allreduce_ll_kernel<DMLB_WIRE_BF16><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, n_lines, scale, sumsq, n_data, Mll);
        else
            allreduce_ll_kernel<DMLB_WIRE_F32><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, n_lines, scale, sumsq, n_data, Mll);
        return launched();
    }
    const int kU = W <= 2 ? 4 : (W <= 4 ? 2 : 1);
    const size_t items = oneshot ? nvec : (nvec + W - 1) / W;  // vectors a CTA grid is spread over
    size_t want = (items + (size_t)kCommThreads * kU - 1) / ((size_t)kCommThreads * kU);
    // all CTAs co-resident (the per-CTA barriers need that); one slot is kept for the metric CTA
    size_t cap = (size_t)min(kMaxCtas, sm_count() * 2) - 1;
    if (want > cap) want = cap;
    const int n_data = n == 0 ? 0 : (int)(want < 1 ? 1 : want);
    const int grid = n_data + (metrics ? 1 : 0);
    const dmlb_step_metrics &M = metrics ? *metrics : kNoMetrics;
#define DMLB_LAUNCH_AR(WIRE, U)                                                                                          \
    do {                                                                                                                 \
        if (oneshot)                                                                                                     \
            allreduce_oneshot_kernel<WIRE, U><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, scale, sumsq,      \
                                                                             n_data, M);                                 \
        else if (nvls_rs_only)                                                                                           \
            allreduce_twoshot_kernel<WIRE, U, 2><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, items, scale,   \
                                                                                sumsq, n_data, M);                       \
        else if (nvls)                                                                                                   \
            allreduce_twoshot_kernel<WIRE, U, 1><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, items, scale,   \
                                                                                sumsq, n_data, M);                       \
        else                                                                                                             \
            allreduce_twoshot_kernel<WIRE, U, 0><<<grid, kCommThreads, 0, st>>>(c->dev, bucket, n, nvec, items, scale,   \
                                                                                sumsq, n_data, M);                       \
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
    // the error word lives in this rank's own arena (control block, word 2); a blocking 4-byte read
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
