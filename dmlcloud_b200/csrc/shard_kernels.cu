// Device-resident data-shard iterator kernels.
//
// Reference: util/data.py:11-30 shard_indices (`indices[rank::world]` after the optional MT19937 shuffle and tail drop)
// and examples/mnist.py:16-21 (torchvision ToTensor + Normalize((0.1307,), (0.3081,)) + DataLoader batching, on the host).
// Here the whole uint8 dataset stays in HBM (MNIST: 47 MB of 180 GB); the epoch permutation is uploaded once per epoch
// (bit-exact host computation — the same numpy MT19937 stream the reference uses) and each step is one gather kernel:
//   out[i, :] = (float(images[idx[i], :]) / 255 - mean) / std                      784 B read, 3136 (fp32) B written / sample
// HBM-bound byte work: 16 pixels (one 128-bit load) per thread, 4x 128-bit stores, rows found through idx[] (L2-resident).
#include "dmlb_common.cuh"

namespace dmlb {

__device__ __forceinline__ float norm_px(uint32_t byte, float mean, float std) {
    // exactly torchvision's arithmetic order: ToTensor -> x/255 ; Normalize -> (x - mean) / std   (IEEE fp32 div/sub/div)
    return ((float)byte / 255.0f - mean) / std;
}

template <bool kBf16>
__global__ void __launch_bounds__(256)
shard_gather_u8_kernel(const uint8_t *__restrict__ images, const long long *__restrict__ idx, long long batch,
                       long long row_elems, long long vec_per_row, float mean, float std, void *out) {
    const long long total = batch * vec_per_row;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / vec_per_row, v = t - i * vec_per_row;
        const long long row = idx[i];
        const uint4 px = ld_stream_u4(reinterpret_cast<const uint4 *>(images + row * row_elems) + v);
        const uint32_t w[4] = {px.x, px.y, px.z, px.w};
        float f[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[4 * k + 0] = norm_px(w[k] & 0xff, mean, std);
            f[4 * k + 1] = norm_px((w[k] >> 8) & 0xff, mean, std);
            f[4 * k + 2] = norm_px((w[k] >> 16) & 0xff, mean, std);
            f[4 * k + 3] = norm_px(w[k] >> 24, mean, std);
        }
        if (kBf16) {
            uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(out) + i * row_elems) + 2 * v;
            o[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            o[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                              pack_bf16x2(f[14], f[15]));
        } else {
            float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + i * row_elems) + 4 * v;
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
        }
    }
}

template <bool kBf16>
__global__ void __launch_bounds__(256)
shard_gather_u8_scalar_kernel(const uint8_t *__restrict__ images, const long long *__restrict__ idx, long long batch,
                              long long row_elems, float mean, float std, void *out) {
    const long long total = batch * row_elems;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / row_elems, e = t - i * row_elems;
        float f = norm_px(images[idx[i] * row_elems + e], mean, std);
        if (kBf16)
            reinterpret_cast<uint16_t *>(out)[t] = f32_to_bf16(f);
        else
            reinterpret_cast<float *>(out)[t] = f;
    }
}

__global__ void shard_gather_i64_kernel(const long long *__restrict__ labels, const long long *__restrict__ idx,
                                        long long batch, long long *out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < batch; i += (long long)gridDim.x * blockDim.x)
        out[i] = labels[idx[i]];
}

__global__ void shard_slice_kernel(const long long *__restrict__ perm, long long first, long long count, long long rank,
                                   long long world, long long *out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = perm[(first + i) * world + rank];
}

static inline int grid_for(long long work, int threads) {
    long long g = (work + threads - 1) / threads;
    long long cap = (long long)sm_count() * 8;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

}  // namespace dmlb

using namespace dmlb;

extern "C" {

int dmlb_shard_gather_u8(const uint8_t *images, const int64_t *idx, int64_t batch, int64_t row_elems, float mean,
                         float std, void *out, int out_bf16, void *stream) {
    if (!images || !idx || !out || batch < 0 || row_elems <= 0 || std == 0.0f) return DMLB_EINVAL;
    if (batch == 0) return DMLB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (row_elems % 16 == 0) && (((uintptr_t)images & 15) == 0) && (((uintptr_t)out & 15) == 0);
    if (vec) {
        long long vpr = row_elems / 16;
        int grid = grid_for(batch * vpr, 256);
        if (out_bf16)
            shard_gather_u8_kernel<true><<<grid, 256, 0, st>>>(images, (const long long *)idx, batch, row_elems, vpr, mean, std, out);
        else
            shard_gather_u8_kernel<false><<<grid, 256, 0, st>>>(images, (const long long *)idx, batch, row_elems, vpr, mean, std, out);
    } else {
        int grid = grid_for(batch * row_elems, 256);
        if (out_bf16)
            shard_gather_u8_scalar_kernel<true><<<grid, 256, 0, st>>>(images, (const long long *)idx, batch, row_elems, mean, std, out);
        else
            shard_gather_u8_scalar_kernel<false><<<grid, 256, 0, st>>>(images, (const long long *)idx, batch, row_elems, mean, std, out);
    }
    return launched();
}

int dmlb_shard_gather_i64(const int64_t *labels, const int64_t *idx, int64_t batch, int64_t *labels_out, void *stream) {
    if (!labels || !idx || !labels_out || batch < 0) return DMLB_EINVAL;
    if (batch == 0) return DMLB_OK;
    shard_gather_i64_kernel<<<grid_for(batch, 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long *)labels, (const long long *)idx, batch, (long long *)labels_out);
    return launched();
}

int dmlb_shard_slice(const int64_t *perm, int64_t first, int64_t count, int64_t rank, int64_t world, int64_t *idx_out,
                     void *stream) {
    if (!perm || !idx_out || first < 0 || count < 0 || world < 1 || rank < 0 || rank >= world) return DMLB_EINVAL;
    if (count == 0) return DMLB_OK;
    shard_slice_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>((const long long *)perm, first, count, rank,
                                                                               world, (long long *)idx_out);
    return launched();
}

}  // extern "C"
