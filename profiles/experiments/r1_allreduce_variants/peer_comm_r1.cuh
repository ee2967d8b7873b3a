// Peer-memory communicator shared by the fused gradient all-reduce and the metric-slab exchange.
//
// Every rank owns one ARENA (cudaMalloc, exported with CUDA IPC, mapped by all peers over NVLink 5 / NVSwitch):
//
//   [0      ..  4 KB)   control: seq (u32), done counter (u32), error word (u32)      — touched only by the owner
//   [4 KB   .. 64 KB)   flags[4 regions][kMaxCtas][8 ranks] u32                       — written by peers, read by owner
//   [64 KB  .. +2*M )   stage[2 halves][M bytes]    this rank's scaled/cast message   — read by peers
//   [ ...   .. +2*M )   result[2 halves][M bytes]   two-shot: this rank's reduced slice — read by peers
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
#pragma once
#include "dmlb_common.cuh"

namespace dmlb {

constexpr int kMaxCtas = 296;  // 2 per SM on 148 SMs
constexpr size_t kCtrlBytes = 4096;
constexpr int kFlagRegions = 4;  // 0, 1: the per-collective barriers; 2, 3: per-chunk flags of the pipelined all-reduces
constexpr size_t kFlagBytes = (size_t)kFlagRegions * kMaxCtas * DMLB_MAX_WORLD * sizeof(uint32_t);
constexpr size_t kHeaderBytes = 65536;
static_assert(kCtrlBytes + kFlagBytes <= kHeaderBytes, "arena header");
constexpr int kCommThreads = 256;

struct CommDev {
    int world, rank;
    size_t msg_cap;  // M: bytes per stage half
    unsigned char *arena[DMLB_MAX_WORLD];
    unsigned long long timeout_ns;

    __device__ __forceinline__ uint32_t *seq() const { return reinterpret_cast<uint32_t *>(arena[rank]); }
    __device__ __forceinline__ uint32_t *done() const { return reinterpret_cast<uint32_t *>(arena[rank]) + 1; }
    __device__ __forceinline__ uint32_t *err() const { return reinterpret_cast<uint32_t *>(arena[rank]) + 2; }
    __device__ __forceinline__ uint32_t *flags(int r, int barrier, int cta) const {
        return reinterpret_cast<uint32_t *>(arena[r] + kCtrlBytes) +
               ((size_t)barrier * kMaxCtas + cta) * DMLB_MAX_WORLD;
    }
    __device__ __forceinline__ unsigned char *stage(int r, int half) const {
        return arena[r] + kHeaderBytes + (size_t)half * msg_cap;
    }
    __device__ __forceinline__ unsigned char *result(int r, int half) const {
        return arena[r] + kHeaderBytes + 2 * msg_cap + (size_t)half * msg_cap;
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

// Sequence number of the collective this kernel performs; every thread of the CTA gets it.
__device__ __forceinline__ uint32_t comm_begin(const CommDev &c) {
    __shared__ uint32_t s_seq;
    if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile uint32_t *>(c.seq()) + 1u;
    __syncthreads();
    return s_seq;
}

// Per-CTA barrier `which` (0 or 1) of collective s.  All threads of the CTA must call.  On entry every thread's prior
// global writes are published to the peers; on exit the peers' writes (made before their arrival) are visible.
__device__ __forceinline__ void comm_barrier(const CommDev &c, int which, uint32_t s) {
    __syncthreads();
    if (threadIdx.x < c.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(c.flags(peer, which, blockIdx.x) + c.rank, s);
        const uint32_t *mine = c.flags(c.rank, which, blockIdx.x) + peer;
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - s) < 0) {
            if (globaltimer_ns() - t0 > c.timeout_ns) {  // a peer died: record it and fall through instead of hanging
                atomicExch(c.err(), 1u);
                break;
            }
        }
    }
    __syncthreads();
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
