// Device-side pieces of the metric slab shared by metric_kernels.cu (stand-alone fold / reduce launches) and
// peer_comm.cu (the fused step exchange: a metric CTA rides along with the gradient all-reduce).
//
// Reference arithmetic being replaced (dmlcloud/metrics.py): MetricReducer.append 66-73, reduce_locally 107-119,
// reduce_globally 121-141 — see metric_kernels.cu for the mapping.
#pragma once
#include <math_constants.h>

#include "dmlb_common.cuh"

namespace dmlb {

__device__ __forceinline__ int desc_op(uint32_t d) { return d & 3; }
__device__ __forceinline__ bool desc_int(uint32_t d) { return (d >> 2) & 1; }
__device__ __forceinline__ bool desc_global(uint32_t d) { return (d >> 3) & 1; }
__device__ __forceinline__ bool desc_f64(uint32_t d) { return (d >> 4) & 1; }

__device__ __forceinline__ uint64_t identity_bits(uint32_t d) {
    const int op = desc_op(d);
    if (desc_int(d)) {
        if (op == DMLB_MIN) return (uint64_t)INT64_MAX;
        if (op == DMLB_MAX) return (uint64_t)INT64_MIN;
        return 0ull;
    }
    if (op == DMLB_MIN) return (uint64_t)__double_as_longlong(CUDART_INF);
    if (op == DMLB_MAX) return (uint64_t)__double_as_longlong(-CUDART_INF);
    return 0ull;
}

// torch.amin/amax propagate NaN; fmin/fmax would drop it
__device__ __forceinline__ double nan_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ double nan_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

__device__ __forceinline__ double combine_f(int op, double a, double b) {
    if (op == DMLB_MIN) return nan_min(a, b);
    if (op == DMLB_MAX) return nan_max(a, b);
    return a + b;
}
__device__ __forceinline__ long long combine_i(int op, long long a, long long b) {
    if (op == DMLB_MIN) return a < b ? a : b;
    if (op == DMLB_MAX) return a > b ? a : b;
    return a + b;
}

__device__ __forceinline__ double load_as_f64(const void *p, int dtype, size_t i) {
    switch (dtype) {
        case DMLB_F32: return (double)reinterpret_cast<const float *>(p)[i];
        case DMLB_F64: return reinterpret_cast<const double *>(p)[i];
        case DMLB_F16: return (double)__half2float(reinterpret_cast<const __half *>(p)[i]);
        case DMLB_BF16: return (double)bf16_to_f32(reinterpret_cast<const uint16_t *>(p)[i]);
        case DMLB_I64: return (double)reinterpret_cast<const long long *>(p)[i];
        case DMLB_I32: return (double)reinterpret_cast<const int *>(p)[i];
        default: return (double)reinterpret_cast<const unsigned char *>(p)[i];
    }
}
__device__ __forceinline__ long long load_as_i64(const void *p, int dtype, size_t i) {
    switch (dtype) {
        case DMLB_I64: return reinterpret_cast<const long long *>(p)[i];
        case DMLB_I32: return (long long)reinterpret_cast<const int *>(p)[i];
        case DMLB_U8: return (long long)reinterpret_cast<const unsigned char *>(p)[i];
        default: return (long long)load_as_f64(p, dtype, i);
    }
}

// Fold ONE entry into the slab with `nthreads` cooperating threads (tid in [0, nthreads), nthreads a multiple of 32;
// every thread of those warps must call).  A value is [lanes, k] row-major (optionally a stack [steps, lanes, k]):
//   steps*k >= 32 : one warp per cell, lanes stride the folded elements, __shfl_xor tree   (batch-style metrics)
//   steps*k <  32 : one thread per cell, sequential                                        (scalars: lanes = k = 1)
// Immediates (src == NULL): `imm` is the pre-combined value of `steps` host scalars (the host merges python scalars that
// hit the same cell between two launches), so acc = op(acc, imm), cnt += steps.
// Feed entries (src_dtype == DMLB_SRC_FEED): src points at one slot of the mapped host feed ring — DMLB_FEED_WIDTH pairs
// {pre-combined value, count} of doubles; entry index = k; count 0 = nothing this step.
__device__ __forceinline__ void fold_entry(uint64_t *acc, long long *cnt, const uint32_t *desc, const dmlb_fold_entry &e,
                                           int tid, int nthreads) {
    const uint32_t d = desc[e.cell];
    const int op = desc_op(d);
    const bool is_int = desc_int(d);
    if (e.src == nullptr || e.src_dtype == DMLB_SRC_FEED) {
        if (tid == 0) {
            double fv = 0.0;
            long long iv = 0, n = e.steps;
            if (e.src == nullptr) {
                fv = __longlong_as_double((long long)e.imm);
                iv = (long long)e.imm;
            } else {  // one 16-byte load = one PCIe read: the row interleaves {value, count} pairs
                const double2 vc = __ldcv(reinterpret_cast<const double2 *>(e.src) + e.k);
                n = (long long)vc.y;
                fv = vc.x;
                iv = (long long)fv;
            }
            if (n > 0) {
                if (is_int)
                    acc[e.cell] = (uint64_t)combine_i(op, (long long)acc[e.cell], iv);
                else
                    acc[e.cell] = (uint64_t)__double_as_longlong(combine_f(op, __longlong_as_double((long long)acc[e.cell]), fv));
                cnt[e.cell] += n;
            }
        }
        return;
    }
    const int k = e.k, steps = e.steps;
    const long long per_cell = (long long)steps * k;  // elements folded into each cell by this entry
    const size_t step_stride = (size_t)e.lanes * k;
    if (per_cell >= 32) {
        const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
        for (int c = warp; c < e.lanes; c += nwarps) {
            const size_t base = (size_t)c * k;
            if (is_int) {
                long long v = (long long)identity_bits(d);
                for (long long t = lane; t < per_cell; t += 32) {
                    const long long st = t / k, j = t - st * k;
                    v = combine_i(op, v, load_as_i64(e.src, e.src_dtype, st * step_stride + base + j));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = combine_i(op, v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) acc[e.cell + c] = (uint64_t)combine_i(op, (long long)acc[e.cell + c], v);
            } else {
                double v = __longlong_as_double((long long)identity_bits(d));
                for (long long t = lane; t < per_cell; t += 32) {
                    const long long st = t / k, j = t - st * k;
                    v = combine_f(op, v, load_as_f64(e.src, e.src_dtype, st * step_stride + base + j));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = combine_f(op, v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0)
                    acc[e.cell + c] = (uint64_t)__double_as_longlong(
                        combine_f(op, __longlong_as_double((long long)acc[e.cell + c]), v));
            }
            if (lane == 0) cnt[e.cell + c] += per_cell;
        }
    } else {
        for (int c = tid; c < e.lanes; c += nthreads) {
            const size_t base = (size_t)c * k;
            if (is_int) {
                long long v = (long long)acc[e.cell + c];
                for (int st = 0; st < steps; ++st)
                    for (int j = 0; j < k; ++j)
                        v = combine_i(op, v, load_as_i64(e.src, e.src_dtype, st * step_stride + base + j));
                acc[e.cell + c] = (uint64_t)v;
            } else {
                double v = __longlong_as_double((long long)acc[e.cell + c]);
                for (int st = 0; st < steps; ++st)
                    for (int j = 0; j < k; ++j)
                        v = combine_f(op, v, load_as_f64(e.src, e.src_dtype, st * step_stride + base + j));
                acc[e.cell + c] = (uint64_t)__double_as_longlong(v);
            }
            cnt[e.cell + c] += per_cell;
        }
    }
}

// selection index -> cell over up to DMLB_MAX_RANGES cell ranges
__device__ __forceinline__ int sel_to_cell(const dmlb_range *r, int n, int i) {
    for (int j = 0; j < n; ++j) {
        int len = r[j].end - r[j].begin;
        if (i < len) return r[j].begin + i;
        i -= len;
    }
    return -1;
}

// local finalisation of one cell -> (value bits, count); optionally resets the cell
__device__ __forceinline__ void finalize_cell(uint64_t *acc, long long *cnt, uint32_t d, int c, uint64_t &val,
                                              long long &n, bool reset) {
    const int op = desc_op(d);
    n = cnt[c];
    uint64_t a = acc[c];
    if (desc_int(d)) {
        val = a;  // (MEAN on integer metrics is rejected on the host, as torch.mean would be)
    } else {
        double v = __longlong_as_double((long long)a);
        if (op == DMLB_MEAN) v = n > 0 ? v / (double)n : 0.0;
        if (!desc_f64(d)) v = (double)(float)v;  // the metric's dtype is fp32: one rounding, like the reference's result
        val = (uint64_t)__double_as_longlong(v);
    }
    if (reset) {
        acc[c] = identity_bits(d);
        cnt[c] = 0;
    }
}

// combine W records of one cell in rank order.  rec(r) -> (val, cnt)
template <class Rec>
__device__ __forceinline__ void combine_cell(uint32_t d, int world, Rec rec, uint64_t &out, uint8_t &flag, int &status) {
    const int op = desc_op(d);
    int empty = 0;
    uint64_t v0;
    long long n0;
    rec(0, v0, n0);
    empty += n0 <= 0;
    if (desc_int(d)) {
        long long a = (long long)v0;
        for (int r = 1; r < world; ++r) {
            uint64_t v;
            long long n;
            rec(r, v, n);
            empty += n <= 0;
            a = combine_i(op == DMLB_MEAN ? DMLB_SUM : op, a, (long long)v);
        }
        out = (uint64_t)a;
    } else if (desc_f64(d)) {
        double a = __longlong_as_double((long long)v0);
        for (int r = 1; r < world; ++r) {
            uint64_t v;
            long long n;
            rec(r, v, n);
            empty += n <= 0;
            a = combine_f(op == DMLB_MEAN ? DMLB_SUM : op, a, __longlong_as_double((long long)v));
        }
        if (op == DMLB_MEAN) a /= (double)world;
        out = (uint64_t)__double_as_longlong(a);
    } else {  // fp32 metric: the cross-rank arithmetic is fp32, like gloo's all_reduce + `tensor /= W`
        float a = (float)__longlong_as_double((long long)v0);
        for (int r = 1; r < world; ++r) {
            uint64_t v;
            long long n;
            rec(r, v, n);
            empty += n <= 0;
            float b = (float)__longlong_as_double((long long)v);
            if (op == DMLB_MIN)
                a = (float)nan_min(a, b);
            else if (op == DMLB_MAX)
                a = (float)nan_max(a, b);
            else
                a = a + b;
        }
        if (op == DMLB_MEAN) a = a / (float)world;
        out = (uint64_t)__double_as_longlong((double)a);
    }
    flag = empty == world ? 1 : 0;
    if (empty != 0 && empty != world) status = DMLB_METRIC_SPLIT_VOTE;
}

}  // namespace dmlb
