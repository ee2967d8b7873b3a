// Peer-memory communicator shared by the fused gradient all-reduce, the fused step exchange and the metric-slab exchange.
//
// Every rank owns one ARENA (cudaMalloc + CUDA IPC, or cuMemCreate + POSIX-fd export when NVSwitch multicast is used),
// mapped by all peers over NVLink 5 / NVSwitch:
//
//   [0      ..  4 KB)   control: seq (u32), done counter (u32), error word (u32)      — touched only by the owner
//   [4 KB   .. 32 KB)   flags[2 barriers][kMaxCtas][8 ranks] u32                      — written by peers, read by owner
//   [32 KB  .. 64 KB)   mstage[2 halves][16 KB]     metric records of the fused step exchange — read by peers
//   [64 KB  .. +2*M )   stage[2 halves][M bytes]    this rank's scaled/cast message   — read by peers
//   [ ...   .. +2*M )   result[2 halves][M bytes]   two-shot: this rank's reduced slice — read by peers
//   [ ...   .. +8.5MB)  ll[2 halves][8 source ranks][544 KB]   LL lines PUSHED by the peers (see below) — read by owner
//
// Protocol (one kernel = one collective, sequence number s = seq+1, half = s & 1):
//   write own stage half  ->  per-CTA flag barrier (store s into every peer's flags[..][cta][me], spin until all 8 of my
//   flags[..][cta][*] >= s)  ->  read every peer's stage half.
// Safety with ONE barrier per collective comes from double buffering: a rank can only start collective s+2 (which
// rewrites half s&1) after finishing s+1, whose barrier needed every peer to have *arrived* at s+1, i.e. to have
// completed all reads of collective s.  Peers can run at most one collective ahead, hence the >= compare.
// CTA b on rank A pairs only with CTA b on the peers (it reads exactly the index range their CTA b wrote), so no
// grid-wide barrier is needed; all ranks must launch the same grid for the same collective (deterministic in n, W).
// A communicator must be driven from ONE stream at a time; the gradient path and the metric path own separate ones.
//
// LL ("low latency") one-shot for small messages — the protocol NCCL calls LL, on our arenas.  A 16-byte line carries 8 bytes
// of payload and the collective's sequence number twice: {data0, s, data1, s}.  A rank PUSHES its scaled/cast message as such
// lines into slot [its rank] of every peer's LL region (posted NVLink stores; each 8-byte half arrives atomically, so a
// half whose flag reads `s` is complete) and then polls its OWN region until every source rank's line shows `s` — no separate
// flag round trip, no peer loads: one one-way NVLink latency instead of a barrier plus a load round trip.  The region is
// written by LL kernels only, so a line's flag words are always sequence numbers (never payload that could alias one), and
// the two halves (s & 1) give the same write-after-read guarantee as for the staging buffers.
//
// Dead peers: a barrier gives up after `timeout_ns` (default 10 minutes, like NCCL's watchdog; dmlb_comm_configure), sets
// the sticky error word in the arena AND — when configured — a word in device-mapped pinned host memory that the host
// polls every step without any synchronisation, and the kernel POISONS its outputs (NaN gradients, TIMEOUT metric status)
// instead of writing a plausible partial sum.
#pragma once
#include "dmlb_common.cuh"

namespace dmlb {

constexpr int kMaxCtas = 296;  // 2 per SM on 148 SMs
constexpr size_t kCtrlBytes = 4096;
constexpr int kFlagRegions = 2;  // the two per-collective barriers
constexpr size_t kFlagBytes = (size_t)kFlagRegions * kMaxCtas * DMLB_MAX_WORLD * sizeof(uint32_t);
constexpr size_t kMetricStageOff = 32768;
constexpr size_t kMetricStageBytes = 16384;  // per half: 16-byte header + 16 B per exchanged cell
constexpr size_t kHeaderBytes = 65536;
static_assert(kCtrlBytes + kFlagBytes <= kMetricStageOff, "arena header: flags");
static_assert(kMetricStageOff + 2 * kMetricStageBytes <= kHeaderBytes, "arena header: metric staging");
constexpr int kCommThreads = 256;
constexpr int kStepMetricMaxCells = (int)(kMetricStageBytes / 16) - 1;
constexpr size_t kLLMaxPayload = 256 * 1024;                                  // largest message (wire bytes) the LL kernel takes
constexpr size_t kLLMetricBytes = 2 * (16 + 16 * (size_t)(kStepMetricMaxCells + 1));  // LL lines of the metric CTA's records
constexpr size_t kLLSlotBytes = 2 * kLLMaxPayload + kLLMetricBytes;          // lines of ONE source rank in one half
constexpr size_t kLLBytes = 2 * DMLB_MAX_WORLD * kLLSlotBytes;               // whole LL region of an arena

struct CommDev {
    int world, rank;
    size_t msg_cap;  // M: bytes per stage half
    unsigned char *arena[DMLB_MAX_WORLD];
    unsigned char *mc;  // multicast mapping of the arenas (NVSwitch in-switch reduction), or nullptr
    unsigned long long timeout_ns;
    uint32_t *host_err;  // device address of a mapped pinned host word (or nullptr)

    __device__ __forceinline__ uint32_t *seq() const { return reinterpret_cast<uint32_t *>(arena[rank]); }
    __device__ __forceinline__ uint32_t *done() const { return reinterpret_cast<uint32_t *>(arena[rank]) + 1; }
    __device__ __forceinline__ uint32_t *err() const { return reinterpret_cast<uint32_t *>(arena[rank]) + 2; }
    __device__ __forceinline__ uint32_t *flags(int r, int barrier, int cta) const {
        return reinterpret_cast<uint32_t *>(arena[r] + kCtrlBytes) +
               ((size_t)barrier * kMaxCtas + cta) * DMLB_MAX_WORLD;
    }
    __device__ __forceinline__ unsigned char *mstage(int r, int half) const {
        return arena[r] + kMetricStageOff + (size_t)half * kMetricStageBytes;
    }
    __device__ __forceinline__ unsigned char *stage(int r, int half) const {
        return arena[r] + kHeaderBytes + (size_t)half * msg_cap;
    }
    __device__ __forceinline__ unsigned char *result(int r, int half) const {
        return arena[r] + kHeaderBytes + 2 * msg_cap + (size_t)half * msg_cap;
    }
    __device__ __forceinline__ unsigned char *mc_stage(int half) const {
        return mc + kHeaderBytes + (size_t)half * msg_cap;
    }
    // LL lines that rank `src` pushed into rank `dst`'s arena for the collective using `half`
    __device__ __forceinline__ uint4 *ll(int dst, int half, int src) const {
        return reinterpret_cast<uint4 *>(arena[dst] + kHeaderBytes + 4 * msg_cap +
                                         ((size_t)half * DMLB_MAX_WORLD + src) * kLLSlotBytes);
    }
    __device__ __forceinline__ uint4 *ll_metric(int dst, int half, int src) const {
        return reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(ll(dst, half, src)) + 2 * kLLMaxPayload);
    }
};

struct Comm {  // host handle
    CommDev dev;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// LL line access: 16-byte volatile (relaxed, system-coherent, L1-bypassing) store / load
__device__ __forceinline__ void ll_store(uint4 *p, uint32_t d0, uint32_t d1, uint32_t s) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(d0), "r"(s), "r"(d1), "r"(s) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// Poll one line of each of the W source ranks until all carry sequence number s (all W loads are in flight together; only
// the lines that were not ready are re-read).  Returns false on timeout / after a peer failure (error word raised).
template <class LineOf>
__device__ __forceinline__ bool ll_wait_all(const CommDev &c, uint32_t s, LineOf line_of, uint4 (&w)[DMLB_MAX_WORLD]) {
    uint32_t ready = 0;
    const uint32_t all = (1u << c.world) - 1u;
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (true) {
#pragma unroll
        for (int r = 0; r < DMLB_MAX_WORLD; ++r)
            if (r < c.world && !(ready >> r & 1u)) w[r] = ll_load(line_of(r));
#pragma unroll
        for (int r = 0; r < DMLB_MAX_WORLD; ++r)
            if (r < c.world && !(ready >> r & 1u) && w[r].y == s && w[r].w == s) ready |= 1u << r;
        if (ready == all) return true;
        if ((++spins & 255u) == 0u) {
            if (t0 == 0) t0 = globaltimer_ns();
            const bool dead = *reinterpret_cast<volatile uint32_t *>(c.err()) != 0u;
            if (dead || globaltimer_ns() - t0 > c.timeout_ns) {
                atomicExch(c.err(), 1u);
                if (c.host_err) {
                    *reinterpret_cast<volatile uint32_t *>(c.host_err) = 1u;
                    __threadfence_system();
                }
                return false;
            }
        }
    }
}

// Sequence number of the collective this kernel performs; every thread of the CTA gets it.
__device__ __forceinline__ uint32_t comm_begin(const CommDev &c) {
    __shared__ uint32_t s_seq;
    if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile uint32_t *>(c.seq()) + 1u;
    __syncthreads();
    return s_seq;
}

// Per-CTA barrier `which` (0 or 1) of collective s.  All threads of the CTA must call.  On entry every thread's prior
// global writes are published to the peers; on exit the peers' writes (made before their arrival) are visible.
// Returns false (to every thread of the CTA) when a peer did not arrive in time: the caller must poison its outputs.
__device__ __forceinline__ bool comm_barrier(const CommDev &c, int which, uint32_t s) {
    __syncthreads();
    int failed = 0;
    if (threadIdx.x < c.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(c.flags(peer, which, blockIdx.x) + c.rank, s);
        const uint32_t *mine = c.flags(c.rank, which, blockIdx.x) + peer;
        const unsigned long long t0 = globaltimer_ns();
        unsigned int spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - s) < 0) {
            if ((++spins & 1023u) == 0u) {
                // a peer that already failed never arrives: give up as soon as the sticky error word is set
                const bool dead = *reinterpret_cast<volatile uint32_t *>(c.err()) != 0u;
                if (dead || globaltimer_ns() - t0 > c.timeout_ns) {
                    atomicExch(c.err(), 1u);
                    if (c.host_err) {
                        *reinterpret_cast<volatile uint32_t *>(c.host_err) = 1u;
                        __threadfence_system();
                    }
                    failed = 1;
                    break;
                }
            }
        }
    }
    return __syncthreads_or(failed) == 0;
}

// Last CTA out publishes seq = s for the next collective on this stream.
__device__ __forceinline__ void comm_end(const CommDev &c, uint32_t s) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        uint32_t d = atomicAdd(c.done(), 1u);
        if (d == gridDim.x - 1) {
            *c.done() = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(c.seq()) = s;
        }
    }
}

}  // namespace dmlb
