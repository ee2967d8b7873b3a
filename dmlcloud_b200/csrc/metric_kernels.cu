// K3 / K4 — device-resident metric slab (sm_100a).
//
// Reference path being replaced (dmlcloud/metrics.py):
//   track() / MetricReducer.append (232-247, 66-73): D2H copy + stream sync per tracked CUDA value, python list append
//   reduce_locally (107-119):  torch.stack(values) then mean/sum/amin/amax over the step axis (+ user dims)
//   reduce_globally (121-141): per metric: all_gather_object emptiness vote (2 gloo all_gathers) + all_reduce  =>
//                              O(#metrics) sequential collectives, 1089 ms for 1024 metrics at W=2
// Here: values never leave the device.  Each metric owns a run of cells {acc, cnt}; a tracked value is folded into its
// cells by one tiny launch (warp-shuffle reduction over the reduced elements), and ONE kernel per reduce_all()
// finalises the local values, exchanges every selected cell with all peers through NVLink peer memory (the vote is the
// comparison of the count lanes that travel with the values), combines in fixed rank order, writes the results and
// resets the cells.  Latency budget: one launch, one flag barrier, 16 B per cell per peer.
//
// Numerics: int64 cells are exact; MIN/MAX exact; float SUM/MEAN accumulate in fp64 locally (error <= the
// reference's fp32 pairwise sum), are rounded to the metric's dtype, then combined across ranks in that dtype in rank
// order — the reference's "mean of per-rank means" (metrics.py:136-138), not a sample-weighted mean.
#include <math_constants.h>

#include "peer_comm.cuh"

namespace dmlb {

constexpr int kFoldThreads = 128;

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

struct FoldParams {
    dmlb_fold_entry e[DMLB_MAX_FOLD_ENTRIES];
};

// grid.x = entry.  A value is [lanes, k] row-major: k elements fold into each of `lanes` cells.
// (optionally a stack [steps, lanes, k] of such values).
//   steps*k >= 32 : one warp per cell, lanes stride the folded elements, __shfl_xor tree   (batch-style metrics)
//   steps*k <  32 : one thread per cell, sequential                                        (scalars: lanes = k = 1)
__global__ void __launch_bounds__(kFoldThreads)
metric_fold_kernel(uint64_t *__restrict__ acc, long long *__restrict__ cnt, const uint32_t *__restrict__ desc,
                   const __grid_constant__ FoldParams P) {
    const dmlb_fold_entry &e = P.e[blockIdx.x];
    const uint32_t d = desc[e.cell];
    const int op = desc_op(d);
    const bool is_int = desc_int(d);
    if (e.src == nullptr) {  // immediate host scalar
        if (threadIdx.x == 0) {
            if (is_int)
                acc[e.cell] = (uint64_t)combine_i(op, (long long)acc[e.cell], (long long)e.imm);
            else
                acc[e.cell] = (uint64_t)__double_as_longlong(
                    combine_f(op, __longlong_as_double((long long)acc[e.cell]), __longlong_as_double((long long)e.imm)));
            cnt[e.cell] += 1;
        }
        return;
    }
    const int k = e.k, steps = e.steps;
    const long long per_cell = (long long)steps * k;  // elements folded into each cell by this entry
    const size_t step_stride = (size_t)e.lanes * k;
    if (per_cell >= 32) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kFoldThreads >> 5;
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
        for (int c = threadIdx.x; c < e.lanes; c += kFoldThreads) {
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

__global__ void metric_reset_kernel(uint64_t *acc, long long *cnt, const uint32_t *desc, int begin, int end) {
    for (int c = begin + blockIdx.x * blockDim.x + threadIdx.x; c < end; c += gridDim.x * blockDim.x) {
        acc[c] = identity_bits(desc[c]);
        cnt[c] = 0;
    }
}

struct RangeParams {
    dmlb_range r[DMLB_MAX_RANGES];
    int n;
};

__device__ __forceinline__ int sel_to_cell(const RangeParams &R, int i) {
    for (int j = 0; j < R.n; ++j) {
        int len = R.r[j].end - R.r[j].begin;
        if (i < len) return R.r[j].begin + i;
        i -= len;
    }
    return -1;
}

// local finalisation of one cell -> (value bits, count); resets the cell
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

// The fused reduce: finalise -> (W>1: peer exchange) -> combine -> results.  Record layout in the staging half:
//   u64[0] layout hash, u64[1] n_sel, then per selected cell {val, cnt}.
__global__ void __launch_bounds__(kCommThreads, 2)
metric_reduce_kernel(const __grid_constant__ CommDev c, bool has_comm, uint64_t *acc, long long *cnt, const uint32_t *__restrict__ desc,
                     const __grid_constant__ RangeParams R, int n_sel, uint64_t layout_hash, bool reset,
                     uint64_t *out_val, uint8_t *out_flag, int *status) {
    const bool exchange = has_comm && c.world > 1;
    uint32_t s = 0;
    int half = 0;
    if (exchange) {
        s = comm_begin(c);
        half = s & 1;
    }
    const int per = (n_sel + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per;
    const int hi = min(n_sel, lo + per);
    uint64_t *rec_mine = exchange ? reinterpret_cast<uint64_t *>(c.stage(c.rank, half)) : nullptr;

    for (int i = lo + threadIdx.x; i < hi; i += kCommThreads) {
        const int cell = sel_to_cell(R, i);
        const uint32_t d = desc[cell];
        uint64_t val;
        long long n;
        finalize_cell(acc, cnt, d, cell, val, n, reset);
        if (exchange && desc_global(d)) {
            rec_mine[2 + 2 * i] = val;
            rec_mine[3 + 2 * i] = (uint64_t)n;
        } else {
            out_val[cell] = val;
            out_flag[cell] = n > 0 ? 0 : 1;
        }
    }
    if (!exchange) return;  // nothing can go wrong locally: the slot keeps whatever this reduce has recorded so far
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rec_mine[0] = layout_hash;
        rec_mine[1] = (uint64_t)n_sel;
    }
    comm_barrier(c, 0, s);

    int st = DMLB_METRIC_OK;
    if (blockIdx.x == 0 && threadIdx.x < c.world) {
        const uint64_t *peer = reinterpret_cast<const uint64_t *>(c.stage(threadIdx.x, half));
        uint4 h = ld_coherent_u4(reinterpret_cast<const uint4 *>(peer));
        // the hash covers the globally-reduced cells only (they lead the selection, so record indices agree); the
        // rank-local tail of the selection may legitimately differ between ranks
        uint64_t ph = ((uint64_t)h.y << 32) | h.x;
        if (ph != layout_hash) st = DMLB_METRIC_LAYOUT;
    }
    for (int i = lo + threadIdx.x; i < hi; i += kCommThreads) {
        const int cell = sel_to_cell(R, i);
        const uint32_t d = desc[cell];
        if (!desc_global(d)) continue;
        uint64_t out;
        uint8_t flag;
        auto rec = [&](int r, uint64_t &v, long long &n) {
            uint4 w = ld_coherent_u4(reinterpret_cast<const uint4 *>(c.stage(r, half)) + 1 + i);
            v = ((uint64_t)w.y << 32) | w.x;
            n = (long long)(((uint64_t)w.w << 32) | w.z);
        };
        combine_cell(d, c.world, rec, out, flag, st);
        out_val[cell] = out;
        out_flag[cell] = flag;
    }
    // One status slot per CTA (DMLB_METRIC_STATUS_SLOTS of them), no atomics.  Slots are sticky (max with what is there)
    // so that a reduce split over several launches keeps an error of an earlier launch; the caller zeroes them per reduce.
    st = __syncthreads_or(st == DMLB_METRIC_LAYOUT) ? DMLB_METRIC_LAYOUT
                                                     : (__syncthreads_or(st == DMLB_METRIC_SPLIT_VOTE) ? DMLB_METRIC_SPLIT_VOTE : DMLB_METRIC_OK);
    if (threadIdx.x == 0 && st != DMLB_METRIC_OK && st > status[blockIdx.x]) status[blockIdx.x] = st;  // clean run: untouched
    comm_end(c, s);
}

// split variant for an external exchange (torch.distributed all_gather): finalize -> [caller gathers] -> combine
__global__ void __launch_bounds__(kCommThreads)
metric_finalize_kernel(uint64_t *acc, long long *cnt, const uint32_t *__restrict__ desc,
                       const __grid_constant__ RangeParams R, int n_sel, uint64_t layout_hash, bool reset,
                       uint64_t *record) {
    for (int i = blockIdx.x * kCommThreads + threadIdx.x; i < n_sel; i += gridDim.x * kCommThreads) {
        const int cell = sel_to_cell(R, i);
        uint64_t val;
        long long n;
        finalize_cell(acc, cnt, desc[cell], cell, val, n, reset);
        record[2 + 2 * i] = val;
        record[3 + 2 * i] = (uint64_t)n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        record[0] = layout_hash;
        record[1] = (uint64_t)n_sel;
    }
}

__global__ void __launch_bounds__(kCommThreads)
metric_combine_kernel(const uint64_t *__restrict__ gathered, int world, int rank, const uint32_t *__restrict__ desc,
                      const __grid_constant__ RangeParams R, int n_sel, uint64_t *out_val, uint8_t *out_flag,
                      int *status) {
    const size_t words = 2 + 2 * (size_t)n_sel;
    int st = DMLB_METRIC_OK;
    if (blockIdx.x == 0 && threadIdx.x < world) {
        const uint64_t *peer = gathered + threadIdx.x * words;
        if (peer[0] != gathered[0] || peer[1] != (uint64_t)n_sel) st = DMLB_METRIC_LAYOUT;
    }
    for (int i = blockIdx.x * kCommThreads + threadIdx.x; i < n_sel; i += gridDim.x * kCommThreads) {
        const int cell = sel_to_cell(R, i);
        const uint32_t d = desc[cell];
        uint64_t out;
        uint8_t flag;
        if (desc_global(d)) {
            auto rec = [&](int r, uint64_t &v, long long &n) {
                v = gathered[r * words + 2 + 2 * i];
                n = (long long)gathered[r * words + 3 + 2 * i];
            };
            combine_cell(d, world, rec, out, flag, st);
        } else {  // local-only metric (globally=False): this rank's own record
            out = gathered[rank * words + 2 + 2 * i];
            flag = (long long)gathered[rank * words + 3 + 2 * i] > 0 ? 0 : 1;
        }
        out_val[cell] = out;
        out_flag[cell] = flag;
    }
    st = __syncthreads_or(st == DMLB_METRIC_LAYOUT) ? DMLB_METRIC_LAYOUT
                                                     : (__syncthreads_or(st == DMLB_METRIC_SPLIT_VOTE) ? DMLB_METRIC_SPLIT_VOTE : DMLB_METRIC_OK);
    if (threadIdx.x == 0 && st != DMLB_METRIC_OK && st > status[blockIdx.x]) status[blockIdx.x] = st;  // clean run: untouched
}

static int fill_ranges(RangeParams &R, const dmlb_range *ranges, int n_ranges, int n_cells, int &n_sel) {
    if (n_ranges < 0 || n_ranges > DMLB_MAX_RANGES || (n_ranges > 0 && !ranges)) return DMLB_ECAPACITY;
    n_sel = 0;
    R.n = n_ranges;
    for (int j = 0; j < n_ranges; ++j) {
        if (ranges[j].begin < 0 || ranges[j].end < ranges[j].begin || (n_cells >= 0 && ranges[j].end > n_cells))
            return DMLB_EINVAL;
        R.r[j] = ranges[j];
        n_sel += ranges[j].end - ranges[j].begin;
    }
    return DMLB_OK;
}

}  // namespace dmlb

using namespace dmlb;

extern "C" {

int dmlb_metric_reset(uint64_t *acc, int64_t *cnt, const uint32_t *desc, int begin, int end, void *stream) {
    if (!acc || !cnt || !desc || begin < 0 || end < begin) return DMLB_EINVAL;
    if (end == begin) return DMLB_OK;
    int grid = (end - begin + 255) / 256;
    if (grid > 148) grid = 148;
    metric_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(acc, (long long *)cnt, desc, begin, end);
    return launched();
}

int dmlb_metric_fold(uint64_t *acc, int64_t *cnt, const uint32_t *desc, const dmlb_fold_entry *entries, int n_entries,
                     void *stream) {
    if (!acc || !cnt || !desc || !entries || n_entries < 0) return DMLB_EINVAL;
    if (n_entries == 0) return DMLB_OK;
    if (n_entries > DMLB_MAX_FOLD_ENTRIES) return DMLB_ECAPACITY;
    FoldParams P;
    for (int i = 0; i < n_entries; ++i) {
        const dmlb_fold_entry &e = entries[i];
        if (e.cell < 0 || e.lanes < 1 || e.k < 1 || e.steps < 1) return DMLB_EINVAL;
        if (e.src == nullptr && (e.lanes != 1 || e.k != 1 || e.steps != 1)) return DMLB_EINVAL;
        if (e.src_dtype < DMLB_F32 || e.src_dtype > DMLB_U8) return DMLB_EINVAL;
        P.e[i] = e;
    }
    metric_fold_kernel<<<n_entries, kFoldThreads, 0, (cudaStream_t)stream>>>(acc, (long long *)cnt, desc, P);
    return launched();
}

int dmlb_metric_reduce(void *comm, uint64_t *acc, int64_t *cnt, const uint32_t *desc, int n_cells,
                       const dmlb_range *ranges, int n_ranges, uint64_t layout_hash, int reset, uint64_t *out_val,
                       uint8_t *out_flag, int32_t *status, void *stream) {
    if (!acc || !cnt || !desc || !out_val || !out_flag || !status) return DMLB_EINVAL;
    RangeParams R;
    int n_sel = 0;
    int rc = fill_ranges(R, ranges, n_ranges, n_cells, n_sel);
    if (rc != DMLB_OK) return rc;
    if (n_sel == 0 && comm == nullptr) return DMLB_OK;  // with a communicator an empty selection still exchanges headers
    CommDev dev{};
    bool has = comm != nullptr;
    if (has) {
        dev = reinterpret_cast<Comm *>(comm)->dev;
        if ((size_t)(2 + 2 * (size_t)n_sel) * 8 > dev.msg_cap) return DMLB_ECAPACITY;
    } else {
        dev.world = 1;
    }
    int grid = (n_sel + kCommThreads - 1) / kCommThreads;
    if (grid > DMLB_METRIC_STATUS_SLOTS) grid = DMLB_METRIC_STATUS_SLOTS;
    if (grid < 1) grid = 1;
    metric_reduce_kernel<<<grid, kCommThreads, 0, (cudaStream_t)stream>>>(dev, has, acc, (long long *)cnt, desc, R, n_sel,
                                                                           layout_hash, reset != 0, out_val, out_flag, status);
    return launched();
}

size_t dmlb_metric_record_words(int n_sel) { return 2 + 2 * (size_t)(n_sel < 0 ? 0 : n_sel); }

int dmlb_metric_finalize(uint64_t *acc, int64_t *cnt, const uint32_t *desc, const dmlb_range *ranges, int n_ranges,
                         uint64_t layout_hash, int reset, uint64_t *record, void *stream) {
    if (!acc || !cnt || !desc || !record) return DMLB_EINVAL;
    RangeParams R;
    int n_sel = 0;
    int rc = fill_ranges(R, ranges, n_ranges, -1, n_sel);
    if (rc != DMLB_OK) return rc;
    int grid = (n_sel + kCommThreads - 1) / kCommThreads;
    grid = grid < 1 ? 1 : (grid > 32 ? 32 : grid);
    metric_finalize_kernel<<<grid, kCommThreads, 0, (cudaStream_t)stream>>>(acc, (long long *)cnt, desc, R, n_sel,
                                                                             layout_hash, reset != 0, record);
    return launched();
}

int dmlb_metric_combine(const uint64_t *gathered, int world, int rank, const uint32_t *desc, const dmlb_range *ranges,
                        int n_ranges, uint64_t *out_val, uint8_t *out_flag, int32_t *status, void *stream) {
    if (!gathered || !desc || !out_val || !out_flag || !status || world < 1 || rank < 0 || rank >= world)
        return DMLB_EINVAL;
    RangeParams R;
    int n_sel = 0;
    int rc = fill_ranges(R, ranges, n_ranges, -1, n_sel);
    if (rc != DMLB_OK) return rc;
    if (n_sel == 0) return DMLB_OK;
    int grid = (n_sel + kCommThreads - 1) / kCommThreads;
    grid = grid > DMLB_METRIC_STATUS_SLOTS ? DMLB_METRIC_STATUS_SLOTS : grid;
    metric_combine_kernel<<<grid, kCommThreads, 0, (cudaStream_t)stream>>>(gathered, world, rank, desc, R, n_sel, out_val,
                                                                            out_flag, status);
    return launched();
}

}  // extern "C"
