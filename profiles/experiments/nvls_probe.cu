// NVLS feasibility probe (round 2): is NVSwitch multicast (cuMulticast* + multimem.ld_reduce / multimem.st) usable on the
// gpurun box, and what does an in-switch all-reduce of the ResNet-18 gradient set cost?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o nvls_probe nvls_probe.cu -lcuda
//   ./nvls_probe            (single process, drives every visible GPU; prints one JSON object)
//
// One process owns all GPUs here (no FD passing needed) — the point is to learn (1) whether the driver / fabric manager
// in the container allows multicast objects at all, (2) whether POSIX-FD export of the handles works (what the
// multi-process product path needs), (3) numerics + speed of:  pack fp32->bf16 -> flag barrier -> multimem.ld_reduce of
// my 1/W slice + multimem.st of the sum (in place) -> flag barrier -> unpack bf16->fp32.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include <vector>

#define CU(x)                                                                    \
    do {                                                                         \
        CUresult _r = (x);                                                       \
        if (_r != CUDA_SUCCESS) {                                                \
            const char *s = nullptr;                                             \
            cuGetErrorString(_r, &s);                                            \
            printf("{\"ok\": false, \"where\": \"%s\", \"line\": %d, \"cuerr\": %d, \"msg\": \"%s\"}\n", #x, __LINE__, (int)_r, s ? s : "?"); \
            exit(0);                                                             \
        }                                                                        \
    } while (0)
#define RT(x)                                                                    \
    do {                                                                         \
        cudaError_t _e = (x);                                                    \
        if (_e != cudaSuccess) {                                                 \
            printf("{\"ok\": false, \"where\": \"%s\", \"line\": %d, \"msg\": \"%s\"}\n", #x, __LINE__, cudaGetErrorString(_e)); \
            exit(0);                                                             \
        }                                                                        \
    } while (0)

constexpr int kMaxW = 8;
constexpr size_t kHeader = 65536;  // flags live in the first 64 KB of every rank's buffer
constexpr int kMaxCtas = 592;

struct Dev {
    int world, rank;
    unsigned char *uc[kMaxW];  // unicast VA of every rank's buffer
    unsigned char *mc;         // multicast VA (same offsets)
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void barrier(const Dev &c, int which, uint32_t s) {
    __syncthreads();
    if (threadIdx.x < c.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        uint32_t *theirs = reinterpret_cast<uint32_t *>(c.uc[peer]) + ((size_t)which * kMaxCtas + blockIdx.x) * kMaxW + c.rank;
        st_release_sys(theirs, s);
        const uint32_t *mine = reinterpret_cast<uint32_t *>(c.uc[c.rank]) + ((size_t)which * kMaxCtas + blockIdx.x) * kMaxW + peer;
        while ((int32_t)(ld_acquire_sys(mine) - s) < 0) {
        }
    }
    __syncthreads();
}

__device__ __forceinline__ uint4 mm_ld_reduce_bf16(const void *mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint4 mm_ld_reduce_f32(const void *mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mm_st(void *mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// kBf16: wire is bf16 (8 el / 16 B vector) else fp32 (4 el / vector).  n multiple of the vector width (probe only).
template <bool kBf16>
__global__ void __launch_bounds__(512, 2)
nvls_allreduce(const __grid_constant__ Dev c, float *bucket, size_t nvec, float scale, uint32_t s, int phases) {
    constexpr int E = kBf16 ? 8 : 4;
    uint4 *stage = reinterpret_cast<uint4 *>(c.uc[c.rank] + kHeader);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    // K1: pack
    if (phases & 1)
        for (size_t g = tid; g < nvec; g += nth) {
            const float4 *src = reinterpret_cast<const float4 *>(bucket + g * E);
            float4 a = src[0];
            uint4 o;
            if (kBf16) {
                float4 b = src[1];
                o.x = pack2(a.x * scale, a.y * scale), o.y = pack2(a.z * scale, a.w * scale);
                o.z = pack2(b.x * scale, b.y * scale), o.w = pack2(b.z * scale, b.w * scale);
            } else {
                o.x = __float_as_uint(a.x * scale), o.y = __float_as_uint(a.y * scale);
                o.z = __float_as_uint(a.z * scale), o.w = __float_as_uint(a.w * scale);
            }
            stage[g] = o;
        }
    barrier(c, 0, s);
    // in-switch reduce of my slice, broadcast of the sum (in place)
    if (phases & 2) {
        const size_t S = (nvec + c.world - 1) / c.world;
        const size_t lo = (size_t)c.rank * S, hi = min(nvec, lo + S);
        unsigned char *mcs = c.mc + kHeader;
        constexpr int U = 4;  // independent in-switch reductions in flight per thread
        for (size_t g0 = lo + tid; g0 < hi; g0 += nth * U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t g = g0 + (size_t)u * nth;
                if (g < hi) v[u] = kBf16 ? mm_ld_reduce_bf16(mcs + g * 16) : mm_ld_reduce_f32(mcs + g * 16);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t g = g0 + (size_t)u * nth;
                if (g < hi) mm_st(mcs + g * 16, v[u]);
            }
        }
    }
    if (phases & 24) {  // 8: in-switch reduce only (result stored to my own staging, unicast); 16: multicast store only
        const size_t S = (nvec + c.world - 1) / c.world;
        const size_t lo = (size_t)c.rank * S, hi = min(nvec, lo + S);
        unsigned char *mcs = c.mc + kHeader;
        constexpr int U = 4;
        for (size_t g0 = lo + tid; g0 < hi; g0 += nth * U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t g = g0 + (size_t)u * nth;
                if (g < hi) v[u] = (phases & 8) ? (kBf16 ? mm_ld_reduce_bf16(mcs + g * 16) : mm_ld_reduce_f32(mcs + g * 16)) : stage[g];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t g = g0 + (size_t)u * nth;
                if (g < hi) {
                    if (phases & 8) stage[g] = v[u];
                    else mm_st(mcs + g * 16, v[u]);
                }
            }
        }
    }
    barrier(c, 1, s);
    // K2: unpack from my own (now reduced) staging
    if (phases & 4)
        for (size_t g = tid; g < nvec; g += nth) {
            uint4 w = stage[g];
            float4 *dst = reinterpret_cast<float4 *>(bucket + g * E);
            if (kBf16) {
                dst[0] = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                                     __uint_as_float(w.y & 0xffff0000u));
                dst[1] = make_float4(__uint_as_float(w.z << 16), __uint_as_float(w.z & 0xffff0000u), __uint_as_float(w.w << 16),
                                     __uint_as_float(w.w & 0xffff0000u));
            } else {
                dst[0] = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
            }
        }
}

__global__ void fill(float *p, size_t n, int rank) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = (float)((i * 2654435761u >> 20) & 63) * (1.0f / 16.0f) - 2.0f + 0.25f * rank;  // exactly representable in bf16
}

int main() {
    CU(cuInit(0));
    int n = 0;
    RT(cudaGetDeviceCount(&n));
    if (n > kMaxW) n = kMaxW;
    const char *env = getenv("NVLS_WORLD");
    if (env && atoi(env) >= 1 && atoi(env) < n) n = atoi(env);
    int supported = 1;
    for (int d = 0; d < n; ++d) {
        RT(cudaSetDevice(d));
        RT(cudaFree(0));
        int v = 0;
        CUdevice dev;
        CU(cuDeviceGet(&dev, d));
        CU(cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
        supported &= v;
    }
    if (!supported || n < 2) {
        printf("{\"ok\": false, \"n_gpus\": %d, \"multicast_supported\": %d}\n", n, supported);
        return 0;
    }
    for (int d = 0; d < n; ++d)
        for (int e = 0; e < n; ++e)
            if (d != e) {
                RT(cudaSetDevice(d));
                cudaDeviceEnablePeerAccess(e, 0);
                cudaGetLastError();
            }
    const size_t payload = 64ull << 20;
    CUmulticastObjectProp mp = {};
    mp.numDevices = n;
    mp.size = kHeader + payload;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0, gmin = 0;
    CU(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    CU(cuMulticastGetGranularity(&gmin, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
    size_t size = ((mp.size + gran - 1) / gran) * gran;
    mp.size = size;
    CUmemGenericAllocationHandle mch;
    CU(cuMulticastCreate(&mch, &mp));
    int fd = -1;
    CUresult fr = cuMemExportToShareableHandle(&fd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    int fd_ok = fr == CUDA_SUCCESS && fd >= 0;
    for (int d = 0; d < n; ++d) {
        CUdevice dev;
        CU(cuDeviceGet(&dev, d));
        CU(cuMulticastAddDevice(mch, dev));
    }
    std::vector<CUmemAccessDesc> acc(n);
    for (int d = 0; d < n; ++d) {
        acc[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc[d].location.id = d;
        acc[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    Dev base = {};
    base.world = n;
    int mem_fd_ok = 1;
    std::vector<CUmemGenericAllocationHandle> mem(n);
    for (int d = 0; d < n; ++d) {
        RT(cudaSetDevice(d));
        CUmemAllocationProp ap = {};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = d;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CU(cuMemCreate(&mem[d], size, &ap, 0));
        int mfd = -1;
        if (cuMemExportToShareableHandle(&mfd, mem[d], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || mfd < 0)
            mem_fd_ok = 0;
        else
            close(mfd);
        CU(cuMulticastBindMem(mch, 0, mem[d], 0, size, 0));
        CUdeviceptr va;
        CU(cuMemAddressReserve(&va, size, gran, 0, 0));
        CU(cuMemMap(va, size, 0, mem[d], 0));
        CU(cuMemSetAccess(va, size, acc.data(), n));
        base.uc[d] = (unsigned char *)va;
        RT(cudaMemset((void *)va, 0, kHeader));
    }
    CUdeviceptr mcva;
    CU(cuMemAddressReserve(&mcva, size, gran, 0, 0));
    CU(cuMemMap(mcva, size, 0, mch, 0));
    CU(cuMemSetAccess(mcva, size, acc.data(), n));
    base.mc = (unsigned char *)mcva;

    const size_t N = 11689512 / 8 * 8;  // ResNet-18 gradient set (rounded to the vector width)
    std::vector<float *> bucket(n);
    std::vector<cudaStream_t> st(n);
    std::vector<cudaEvent_t> e0(n), e1(n);
    for (int d = 0; d < n; ++d) {
        RT(cudaSetDevice(d));
        RT(cudaMalloc(&bucket[d], N * 4));
        RT(cudaStreamCreate(&st[d]));
        RT(cudaEventCreate(&e0[d]));
        RT(cudaEventCreate(&e1[d]));
    }
    uint32_t seq = 0;
    auto run = [&](bool bf16, int phases, int grid, int reps, float *us_out) {
        const size_t nvec = N / (bf16 ? 8 : 4);
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            RT(cudaEventRecord(e0[d], st[d]));
        }
        for (int r = 0; r < reps; ++r) {
            ++seq;
            for (int d = 0; d < n; ++d) {
                RT(cudaSetDevice(d));
                Dev c = base;
                c.rank = d;
                if (bf16)
                    nvls_allreduce<true><<<grid, 512, 0, st[d]>>>(c, bucket[d], nvec, 1.0f / n, seq, phases);
                else
                    nvls_allreduce<false><<<grid, 512, 0, st[d]>>>(c, bucket[d], nvec, 1.0f / n, seq, phases);
            }
        }
        float worst = 0;
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            RT(cudaEventRecord(e1[d], st[d]));
        }
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            RT(cudaEventSynchronize(e1[d]));
            float ms;
            RT(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            if (ms > worst) worst = ms;
        }
        *us_out = worst * 1e3f / reps;
    };
    // ---- numerics: inputs exactly representable in bf16, scale 1/n exact for n = 2,4,8 -> result exact ----
    int bad[2] = {0, 0};
    for (int pass = 0; pass < 2; ++pass) {
        const bool bf16 = pass == 0;
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            fill<<<296, 512, 0, st[d]>>>(bucket[d], N, d);
            RT(cudaStreamSynchronize(st[d]));
        }
        float us;
        run(bf16, 7, 148, 1, &us);
        std::vector<float> h(4096);
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            const size_t off = (N / 3) & ~(size_t)7;
            RT(cudaMemcpy(h.data(), bucket[d] + off, h.size() * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < h.size(); ++i) {
                const size_t gi = off + i;
                float want = 0;
                for (int r = 0; r < n; ++r) want += ((float)((gi * 2654435761u >> 20) & 63) * (1.0f / 16.0f) - 2.0f + 0.25f * r) / n;
                float tol = bf16 ? 0.02f : 1e-5f;  // the bf16 all-gather half rounds the sum to bf16
                if (fabsf(h[i] - want) > tol) ++bad[pass];
            }
        }
    }
    // ---- timing ----
    printf("{\"ok\": true, \"n_gpus\": %d, \"multicast_supported\": 1, \"granularity\": %zu, \"granularity_min\": %zu, \"mc_fd_export\": %d, "
           "\"mem_fd_export\": %d, \"mismatch_bf16\": %d, \"mismatch_f32\": %d, \"elements\": %zu, \"timings\": [",
           n, gran, gmin, fd_ok, mem_fd_ok, bad[0], bad[1], N);
    bool first = true;
    for (int pass = 0; pass < 2; ++pass) {
        const bool bf16 = pass == 0;
        const double wire_bytes = (double)N * (bf16 ? 2 : 4);
        for (int grid : {148, 296, 592}) {
            if (grid > 296) continue;  // co-residency: at most two 512-thread CTAs per SM
            for (int phases : {7, 2, 5, 8, 16}) {
                float us;
                run(bf16, phases, grid, 3, &us);
                run(bf16, phases, grid, 20, &us);
                double bus = 2.0 * (n - 1) / n * wire_bytes / (us * 1e-6) / 1e9;
                printf("%s{\"wire\": \"%s\", \"grid\": %d, \"phases\": %d, \"us\": %.2f, \"busbw_GBps\": %.1f}", first ? "" : ", ",
                       bf16 ? "bf16" : "fp32", grid, phases, us, bus);
                first = false;
            }
        }
    }
    printf("]}\n");
    return 0;
}
