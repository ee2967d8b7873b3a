// Experiment (not product code): does the store policy change what K2 (bf16 -> fp32 widen, 2 B read + 4 B written per
// element) reaches on a cold 1 GiB bucket?  DESIGN.md claims K2's ~0.87 of the copy peak is a DRAM read/write-mix
// property; this binary tests the alternatives that could refute it.  Same loop shape as bucket_kernels.cu's
// stream_kernel<UnpackBf16> (512 threads, 592 CTAs, grid-stride, 4 independent vectors per thread per iteration).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/experiments/k2_store_policy profiles/experiments/k2_store_policy.cu
//   ./profiles/experiments/k2_store_policy > gpurun_out/k2_store_policy.json
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

enum Policy { kDefault, kCs, kNoAllocL1, kWt, kEvictFirstStore, kEvictFirstBoth, kV8, kV8EvictFirst, kNumPolicies };
static const char *kNames[] = {"st.global (default write-back)",
                               "st.global.cs (streaming)",
                               "st.global.L1::no_allocate",
                               "st.global.wt (write-through)",
                               "st.global.L2::cache_hint evict_first",
                               "ld+st L2::cache_hint evict_first",
                               "256-bit st.global.v8.f32 (8 el/thread-vector)",
                               "256-bit store + L2 evict_first"};

__device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int P>
__device__ __forceinline__ uint2 load8(const uint2 *p, uint64_t pol) {
    uint2 v;
    if (P == kEvictFirstBoth)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

template <int P>
__device__ __forceinline__ void store16(float *p, float4 v, uint64_t pol) {
    if (P == kCs)
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (P == kNoAllocL1)
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (P == kWt)
        asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (P == kEvictFirstStore || P == kEvictFirstBoth)
        asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    else
        *reinterpret_cast<float4 *>(p) = v;
}

template <int P>
__global__ void __launch_bounds__(512) widen4(const uint2 *__restrict__ src, float *__restrict__ dst, size_t nvec, float s) {
    const uint64_t pol = (P == kEvictFirstStore || P == kEvictFirstBoth) ? evict_first_policy() : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
        uint2 w[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < nvec) w[u] = load8<P>(src + i0 + u * stride, pol);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < nvec)
                store16<P>(dst + 4 * (i0 + u * stride), make_float4(lo(w[u].x) * s, hi(w[u].x) * s, lo(w[u].y) * s, hi(w[u].y) * s), pol);
    }
}

// 8 elements per vector: one 128-bit load, one 256-bit store
template <int P>
__global__ void __launch_bounds__(512) widen8(const uint4 *__restrict__ src, float *__restrict__ dst, size_t nvec, float s) {
    const uint64_t pol = P == kV8EvictFirst ? evict_first_policy() : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int U = 2;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
        uint4 w[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < nvec)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w) : "l"(src + i0 + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < nvec) {
                float *p = dst + 8 * (i0 + u * stride);
                if (P == kV8EvictFirst)
                    asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(p),
                                 "f"(lo(w[u].x) * s), "f"(hi(w[u].x) * s), "f"(lo(w[u].y) * s), "f"(hi(w[u].y) * s),
                                 "f"(lo(w[u].z) * s), "f"(hi(w[u].z) * s), "f"(lo(w[u].w) * s), "f"(hi(w[u].w) * s), "l"(pol) : "memory");
                else
                    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
                                 "f"(lo(w[u].x) * s), "f"(hi(w[u].x) * s), "f"(lo(w[u].y) * s), "f"(hi(w[u].y) * s),
                                 "f"(lo(w[u].z) * s), "f"(hi(w[u].z) * s), "f"(lo(w[u].w) * s), "f"(hi(w[u].w) * s) : "memory");
            }
    }
}

template <int P>
static void launch(const void *src, float *dst, size_t n, int grid) {
    if (P == kV8 || P == kV8EvictFirst)
        widen8<P><<<grid, 512>>>(reinterpret_cast<const uint4 *>(src), dst, n / 8, 1.0f);
    else
        widen4<P><<<grid, 512>>>(reinterpret_cast<const uint2 *>(src), dst, n / 4, 1.0f);
}

typedef void (*LaunchFn)(const void *, float *, size_t, int);
static LaunchFn kLaunch[] = {launch<kDefault>, launch<kCs>, launch<kNoAllocL1>, launch<kWt>,
                             launch<kEvictFirstStore>, launch<kEvictFirstBoth>, launch<kV8>, launch<kV8EvictFirst>};

int main() {
    const size_t n = (size_t)256 << 20;  // 268,435,456 elements: 0.5 GiB bf16 in, 1 GiB fp32 out
    const double peak = 6575.1;          // MEASURED_PEAKS.json hbm copy GB/s
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4;
    void *src;
    float *dst;
    if (cudaMalloc(&src, n * 2) != cudaSuccess || cudaMalloc(&dst, n * 4) != cudaSuccess) return 1;
    cudaMemset(src, 0x3f, n * 2);
    cudaMemset(dst, 0, n * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int reps = 12;
    std::vector<std::vector<float>> us(kNumPolicies);
    for (int p = 0; p < kNumPolicies; ++p) kLaunch[p](src, dst, n, grid);  // warm-up, also surfaces launch errors
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fprintf(stderr, "warm-up failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 2;
    }
    for (int r = 0; r < reps; ++r)
        for (int p = 0; p < kNumPolicies; ++p) {  // interleaved so clock / thermal drift hits every variant alike
            cudaEventRecord(a);
            kLaunch[p](src, dst, n, grid);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            us[p].push_back(ms * 1e3f);
        }
    // spot check of the last variant's output: 0x3f3f bf16 == 0.74609375
    float h[8];
    cudaMemcpy(h, dst + n - 8, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"experiment\": \"K2 store policy, 268435456 elements cold (1.61 GB of traffic)\", \"grid\": %d, \"threads\": 512, "
           "\"peak_GBps\": %.1f, \"tail_value\": %.8f, \"variants\": [\n", grid, peak, h[7]);
    for (int p = 0; p < kNumPolicies; ++p) {
        std::sort(us[p].begin(), us[p].end());
        const double med = us[p][reps / 2], best = us[p][0];
        printf("  {\"store\": \"%s\", \"median_us\": %.2f, \"best_us\": %.2f, \"GBps\": %.1f, \"frac_of_peak\": %.4f}%s\n", kNames[p], med,
               best, n * 6.0 / med / 1e3, n * 6.0 / med / 1e3 / peak, p + 1 < kNumPolicies ? "," : "");
    }
    printf("]}\n");
    return 0;
}
