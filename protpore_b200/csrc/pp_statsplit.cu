// pp_statsplit.cu -- batched FastStatSplit segmenter for sm_100a: raw current -> breakpoints -> segment means.
//
// The step immediately before the HMM path: PyPore's FastStatSplit.parse
// (/root/reference/PyPore/cparsers.pyx:103-118) builds the cumulative sums of the current and its square,
// splits [0, n) recursively at the position of largest log-variance gain inside a sliding window
// (_best_split_stepwise 157-178, _recursive_split 180-203) and hands the segment means to the HMM
// (peptide_hmm_v0.0.1.py:443-466).  Here one CTA owns one event:
//
//   k_cumsum      cumulative sums.  numpy.cumsum is a SEQUENTIAL float64 accumulation and every gain is a difference
//                 of its entries, so the order of the additions is part of the result: a chain cannot be
//                 parallelised, but the 2 x n_events chains are independent.  One LANE owns one chain (a warp takes
//                 16 events: lanes 0-15 accumulate c, lanes 16-31 c2 of the same events); 16 x 32-sample tiles are
//                 staged through shared memory so that every global load and store is a coalesced 256-byte row.
//   k_statsplit   (one CTA per event) the recursion as an explicit interval stack in shared memory.  A window scan is
//                          parallel over the candidate positions: float64 gain with explicitly rounded
//                          operations (no FMA contraction) and a (gain, position) arg-max whose tie rule --
//                          lowest position among equal gains, gain strictly above the threshold -- is what the
//                          reference's ascending strict-'>' scan leaves.  Splits go to an unordered per-event list.
//                 (a forced split may cut off a piece shorter than min_width: at most 2 * (n / min_width) segments)
//   k_segments    rank-sorts the list (ascending = the reference's in-order concatenation) and reduces every segment to
//                 mean / std with one warp per segment (Segment.mean / .std, PyPore/core.py:209-214).
//
// Float64 `log` here is CUDA's (<= 1 ulp), the reference's is glibc's: a window whose two best gains agree to
// ~1e-13 relative could split elsewhere.  The CPU oracle reports that margin per event; tests accept a different
// breakpoint list only when it is below 1e-9 (none observed).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "pp_common.h"

namespace pp {
namespace {

constexpr int kSegThreads = 256;
constexpr int kCsWarps = 4;            // warps per CTA of k_cumsum
constexpr int kCsEvents = 16;          // events per warp of k_cumsum (lane = event + 16 * {c, c2})
constexpr int kSegStack = 1024;        // pending intervals per event (depth-first: bounded by the recursion depth)

struct SegParams {
    int min_width, max_width, window_width, step;
    double min_gain;
};

__device__ __forceinline__ bool better(double g, int i, double bg, int bi) {
    return g > bg || (g == bg && i >= 0 && i < bi);
}

// a / b for an integer-valued b with y = RN(1 / b): q0 = RN(a y), r = a - q0 b (exact in an FMA), q = RN(q0 + r y) is the
// correctly rounded quotient (Markstein) -- 3 instructions instead of the ~35 of a float64 divide.  Checked bit for bit
// against IEEE division on the GPU (tests/test_statsplit.py::test_gpu_division_by_reciprocal_is_exact).
__device__ __forceinline__ double div_rcp(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    const double r = __fma_rn(-q0, b, a);
    return __fma_rn(r, y, q0);
}

// cparsers.pyx:157-178 -- all threads of the CTA call this with the same arguments and get the same answer.
// TABLE: s_rcp[n] = RN(1 / n) for n <= window_width is resident in shared memory (every segment length of a scan is
// an integer below the window width); otherwise the reciprocal is computed per position.
template <bool TABLE>
__device__ int best_split_stepwise(const SegParams& P, const double* __restrict__ c, const double* __restrict__ c2,
                                   int start, int end, double* s_gain, int* s_idx, const double* s_rcp) {
    if (end - start <= 2 * P.min_width) return -1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // var_c (31-38) with c[-1] = c2[-1] = 0: x - 0.0 and n - 0 are exact, so the start == 0 branch is the same arithmetic
    const double cb = start > 0 ? c[start - 1] : 0.0, c2b = start > 0 ? c2[start - 1] : 0.0;
    const double ce = c[end - 1], c2e = c2[end - 1];
    const double nt = static_cast<double>(end - start);
    const double mt = __ddiv_rn(__dsub_rn(ce, cb), nt);
    const double vt = __dsub_rn(__ddiv_rn(__dsub_rn(c2e, c2b), nt), __dmul_rn(mt, mt));
    const double var_summed = __dmul_rn(nt, log(vt));
    double bg = P.min_gain;
    int bi = -1;
    for (int i = start + P.min_width + tid; i < end + 1 - P.min_width; i += kSegThreads) {
        const double ci = c[i - 1], c2i = c2[i - 1];
        const double nl = static_cast<double>(i - start), nh = static_cast<double>(end - i);
        const double yl = TABLE ? s_rcp[i - start] : __drcp_rn(nl);
        const double yh = TABLE ? s_rcp[end - i] : __drcp_rn(nh);
        const double ml = div_rcp(__dsub_rn(ci, cb), nl, yl);
        const double vl = __dsub_rn(div_rcp(__dsub_rn(c2i, c2b), nl, yl), __dmul_rn(ml, ml));
        const double mh = div_rcp(__dsub_rn(ce, ci), nh, yh);
        const double vh = __dsub_rn(div_rcp(__dsub_rn(c2e, c2i), nh, yh), __dmul_rn(mh, mh));
        const double low = __dmul_rn(nl, log(vl));
        const double high = __dmul_rn(nh, log(vh));
        const double gain = __dsub_rn(var_summed, __dadd_rn(low, high));
        if (gain > bg) { bg = gain; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double og = __shfl_down_sync(0xffffffffu, bg, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (better(og, oi, bg, bi)) { bg = og; bi = oi; }
    }
    if (lane == 0) { s_gain[warp] = bg; s_idx[warp] = bi; }
    __syncthreads();
    double rg = s_gain[0];
    int ri = s_idx[0];
#pragma unroll
    for (int w = 1; w < kSegThreads / 32; ++w)
        if (better(s_gain[w], s_idx[w], rg, ri)) { rg = s_gain[w]; ri = s_idx[w]; }
    __syncthreads();                       // s_gain / s_idx are reused by the next scan
    return ri;
}

// test hook: q[i] = a[i] / b[i] by IEEE division and by div_rcp; returns the number of mismatching bit patterns
__global__ void k_div_check(const double* __restrict__ a, const int* __restrict__ b, int n, unsigned long long* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double bd = static_cast<double>(b[i]);
    const double q = __ddiv_rn(a[i], bd), p = div_rcp(a[i], bd, __drcp_rn(bd));
    if (__double_as_longlong(q) != __double_as_longlong(p) && !(q == 0.0 && p == 0.0)) {   // a = -0.0 gives +0.0: log(+-0) is the same -inf
        atomicAdd(bad, 1ULL); atomicMin(bad + 1, static_cast<unsigned long long>(i)); }
}

// c = cumsum(x), c2 = cumsum(x * x), sequential float64 per event (cparsers.pyx:110-111).
// Shared memory per warp: [c | c2][16 events][32 samples + 1]; the staging pass writes x and x * x, every lane then
// accumulates ITS row in place, and the rows go back to global memory as coalesced 256-byte stores.
__global__ void __launch_bounds__(kCsWarps * 32)
k_cumsum(const double* __restrict__ current, const int64_t* __restrict__ offsets, int n_events, double* __restrict__ cs,
         double* __restrict__ cs2) {
    constexpr int ROW = 33;                                    // 32 samples + 1: lanes of different events hit different banks
    constexpr int ARR = kCsEvents * ROW + 1;                   // c2 rows shifted by one bank pair against the c rows
    __shared__ double s_tile[kCsWarps][2 * ARR];
    __shared__ int64_t s_off[kCsWarps][kCsEvents];
    __shared__ int s_n[kCsWarps][kCsEvents];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* t = s_tile[warp];
    const int ev0 = (blockIdx.x * kCsWarps + warp) * kCsEvents;
    if (ev0 >= n_events) return;
    const int my_e = lane & (kCsEvents - 1), arr = lane >> 4;
    int64_t my_off = 0;
    int my_n = 0;
    if (ev0 + my_e < n_events) {
        my_off = offsets[ev0 + my_e];
        my_n = static_cast<int>(offsets[ev0 + my_e + 1] - my_off);
    }
    int nmax = my_n, nmin = my_n;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        nmin = min(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    }
    if (lane < kCsEvents) { s_off[warp][lane] = my_off; s_n[warp][lane] = my_n; }
    __syncwarp();
    int64_t offe[kCsEvents];
#pragma unroll
    for (int e = 0; e < kCsEvents; ++e) offe[e] = s_off[warp][e] + lane;
    // software pipeline: the 16 row loads of tile t + 1 are in flight while the lanes walk tile t
    double pre[kCsEvents];
#pragma unroll
    for (int e = 0; e < kCsEvents; ++e) pre[e] = lane < s_n[warp][e] ? current[offe[e]] : 0.0;
    double s = 0.0;
    double* r = t + arr * ARR + my_e * ROW;                    // this lane's row
    for (int t0 = 0; t0 < nmax; t0 += 32) {
        const bool full = t0 + 32 <= nmin;                     // no row ends inside this tile (warp-uniform)
#pragma unroll
        for (int e = 0; e < kCsEvents; ++e) {
            t[e * ROW + lane] = pre[e];
            t[ARR + e * ROW + lane] = __dmul_rn(pre[e], pre[e]);
        }
        __syncwarp();
        if (t0 + 64 <= nmin) {
#pragma unroll
            for (int e = 0; e < kCsEvents; ++e) pre[e] = current[offe[e] + t0 + 32];
        } else if (t0 + 32 < nmax) {
            const int i = t0 + 32 + lane;
#pragma unroll
            for (int e = 0; e < kCsEvents; ++e) pre[e] = i < s_n[warp][e] ? current[offe[e] + t0 + 32] : 0.0;
        }
        if (full && t0 > 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) { s = __dadd_rn(s, r[k]); r[k] = s; }
        } else {
            const int cnt = min(32, my_n - t0);
            int k = 0;
            if (t0 == 0 && cnt > 0) { s = r[0]; k = 1; }       // c[0] = x[0] (no 0.0 + x: keeps a -0.0)
            for (; k < cnt; ++k) { s = __dadd_rn(s, r[k]); r[k] = s; }
        }
        __syncwarp();
        if (full) {
#pragma unroll
            for (int e = 0; e < kCsEvents; ++e) {              // 2 x 16 rows back, coalesced
                cs[offe[e] + t0] = t[e * ROW + lane];
                cs2[offe[e] + t0] = t[ARR + e * ROW + lane];
            }
        } else {
            const int i = t0 + lane;
#pragma unroll
            for (int e = 0; e < kCsEvents; ++e) {
                if (i < s_n[warp][e]) {
                    cs[offe[e] + t0] = t[e * ROW + lane];
                    cs2[offe[e] + t0] = t[ARR + e * ROW + lane];
                }
            }
        }
        __syncwarp();
    }
}

template <bool TABLE>
__global__ void __launch_bounds__(kSegThreads)
k_statsplit(const SegParams P, const int64_t* __restrict__ offsets, int n_events,
            const double* __restrict__ cs, const double* __restrict__ cs2, int32_t* __restrict__ splits,
            int32_t* __restrict__ seg_count, int32_t* __restrict__ status) {
    __shared__ int2 s_stack[kSegStack];
    __shared__ double s_gain[kSegThreads / 32];
    __shared__ int s_idx[kSegThreads / 32];
    __shared__ int s_sp, s_fail;
    extern __shared__ double s_rcp[];
    const int tid = threadIdx.x;
    if (TABLE) {
        for (int k = tid; k <= P.window_width; k += kSegThreads) s_rcp[k] = k > 0 ? __drcp_rn(static_cast<double>(k)) : 0.0;
        __syncthreads();
    }
    for (int ev = blockIdx.x; ev < n_events; ev += gridDim.x) {
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        const double* c = cs + off;
        const double* c2 = cs2 + off;
        // Split list of this event.  A forced split (198-201) may cut off a piece shorter than min_width, but then its sibling
        // is exactly min_width long and final, so an event has at most 2 * (n / min_width) segments; the base below is safe
        // because sum_j floor(n_j / w) <= floor(off / w).  The bound is checked, not trusted.
        int32_t* sl = splits + (2 * (off / P.min_width) + ev);
        const int sl_cap = 2 * (n / P.min_width) + 1;

        if (tid == 0) { s_stack[0] = make_int2(0, n); s_sp = 1; s_fail = 0; }
        int n_splits = 0;                   // thread 0 only
        __syncthreads();

        // _recursive_split (180-203) with an explicit stack, left part first
        while (true) {
            const int sp = s_sp;
            if (sp == 0) break;
            const int2 iv = s_stack[sp - 1];
            __syncthreads();
            const int start = iv.x, end = iv.y;
            int split_at = -1;
            bool right_only = false;
            for (long long ps = start; ps < static_cast<long long>(end) - 2LL * P.min_width; ps += P.step) {
                if (ps > static_cast<long long>(start) + P.max_width) {                     // 188-190
                    split_at = static_cast<int>(min(static_cast<long long>(start) + P.max_width,
                                                    static_cast<long long>(end) - P.min_width));
                    right_only = true;
                    break;
                }
                const int pe = static_cast<int>(min(static_cast<long long>(end), ps + P.window_width));
                split_at = best_split_stepwise<TABLE>(P, c, c2, static_cast<int>(ps), pe, s_gain, s_idx, s_rcp);
                if (split_at >= 0) break;
            }
            if (split_at == -1 && static_cast<long long>(end) - start > P.max_width)          // 198-201
                split_at = static_cast<int>(min(static_cast<long long>(start) + P.max_width,
                                                static_cast<long long>(end) - P.min_width));
            if (tid == 0) {
                int top = sp - 1;
                if (split_at >= 0) {
                    if (split_at <= start || split_at >= end || top + 2 > kSegStack || n_splits >= sl_cap) {
                        s_fail = 1;                                  // stack exhausted or a split that cannot terminate
                        top = 0;
                    } else {
                        sl[n_splits++] = split_at;
                        s_stack[top++] = make_int2(split_at, end);
                        if (!right_only) s_stack[top++] = make_int2(start, split_at);
                    }
                }
                s_sp = top;
            }
            __syncthreads();
        }
        if (tid == 0) {
            seg_count[ev] = n_splits + 1;
            if (s_fail) atomicExch(status, 1);
        }
        __syncthreads();
    }
}

// unordered split list -> ascending segment starts (rank sort; the splits of an event are distinct), then mean / std of
// every segment with one warp per segment
constexpr int kSegSortCap = 2048;
__global__ void __launch_bounds__(kSegThreads)
k_segments(const double* __restrict__ current, const int64_t* __restrict__ offsets, int n_events, int min_width,
           const int32_t* __restrict__ splits, const int64_t* __restrict__ seg_offsets,
           int32_t* __restrict__ seg_start, double* __restrict__ seg_mean, double* __restrict__ seg_std,
           float* __restrict__ seg_mean_f32) {
    __shared__ int s_split[kSegSortCap];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int ev = blockIdx.x; ev < n_events; ev += gridDim.x) {
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        const int64_t so = seg_offsets[ev];
        const int n_seg = static_cast<int>(seg_offsets[ev + 1] - so);
        const int k = n_seg - 1;
        const int32_t* sl = splits + (2 * (off / min_width) + ev);
        const bool in_smem = k <= kSegSortCap;
        if (in_smem) for (int j = tid; j < k; j += kSegThreads) s_split[j] = sl[j];
        if (tid == 0) seg_start[so] = 0;
        __syncthreads();
        for (int j = tid; j < k; j += kSegThreads) {
            const int v = in_smem ? s_split[j] : sl[j];
            int rank = 0;
            if (in_smem) for (int q = 0; q < k; ++q) rank += s_split[q] < v;
            else for (int q = 0; q < k; ++q) rank += sl[q] < v;
            seg_start[so + 1 + rank] = v;
        }
        __threadfence_block();
        __syncthreads();
        for (int j = warp; j < n_seg; j += kSegThreads / 32) {
            const int a = seg_start[so + j];
            const int b = j + 1 < n_seg ? seg_start[so + j + 1] : n;
            double s = 0.0;
            for (int i = a + lane; i < b; i += 32) s += current[off + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const double mean = s / static_cast<double>(b - a);
            double q = 0.0;
            if (seg_std) {
                for (int i = a + lane; i < b; i += 32) { const double d = current[off + i] - mean; q += d * d; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (lane == 0) {
                if (seg_mean) seg_mean[so + j] = mean;
                if (seg_std) seg_std[so + j] = sqrt(q / static_cast<double>(b - a));
                if (seg_mean_f32) seg_mean_f32[so + j] = static_cast<float>(mean);
            }
        }
        __syncthreads();
    }
}

// per-device workspace, reused across calls
struct SegCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int sm_count = 0;
    DevBuf d_cur, d_off, d_c, d_c2, d_marks, d_count, d_status, d_segoff, d_start, d_mean, d_std, d_mean32;
    int64_t launches = 0;
    float last_ms[3] = {0.f, 0.f, 0.f};
};
std::mutex g_seg_mutex;
std::vector<SegCtx*> g_seg_ctx;

int seg_ctx(int device, SegCtx** out) {
    for (SegCtx* c : g_seg_ctx)
        if (c->device == device) { *out = c; return PP_OK; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_error("no CUDA device %d (the segmenter has no CPU fallback)", device);
        return PP_ERR_CUDA;
    }
    PP_CUDA(cudaSetDevice(device));
    SegCtx* c = new SegCtx();
    c->device = device;
    PP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& e : c->ev) PP_CUDA(cudaEventCreate(&e));
    PP_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    g_seg_ctx.push_back(c);
    *out = c;
    return PP_OK;
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" {

double pp_statsplit_min_gain(int64_t window_width, double min_gain_per_sample, double false_positive_rate,
                             double prior_segments_per_second, double sampling_freq, double cutoff_freq) {
    // FastStatSplit.__init__ (cparsers.pyx:56-101); 0 stands for an argument the caller left out (`if not x`)
    if (false_positive_rate == 0.0) false_positive_rate = sampling_freq;
    if (prior_segments_per_second == 0.0) prior_segments_per_second = sampling_freq / 2.;
    double g;
    if (min_gain_per_sample != 0.0) {
        g = min_gain_per_sample * static_cast<double>(window_width);
    } else {
        const double k = cutoff_freq != 0.0 ? cutoff_freq / (0.5 * sampling_freq) : 1.0;
        const double sps = prior_segments_per_second;
        g = (-std::log(sps / (sampling_freq - sps)) - std::log(false_positive_rate / sampling_freq)) / k;
    }
    return g * 2;
}

int pp_statsplit_batch(const pp_statsplit_params* prm, int device, const double* current, const int64_t* offsets,
                       int64_t n_events, int32_t mem, int64_t* seg_offsets, int32_t* seg_start, double* seg_mean,
                       double* seg_std, float* seg_mean_f32, int64_t capacity, int64_t* seg_total) {
    if (!prm || !current || !offsets || !seg_offsets || !seg_total || n_events < 0 || capacity < 0) {
        set_error("null or negative argument");
        return PP_ERR_ARG;
    }
    // the reference asserts the last two (cparsers.pyx:69-72); min_width >= 1 keeps every split inside its interval
    if (prm->min_width < 1 || prm->max_width < prm->min_width || prm->window_width < 2 * prm->min_width ||
        prm->window_width / 2 < 1 || prm->max_width > 0x3fffffff || prm->window_width > 0x3fffffff) {
        set_error("bad segmenter parameters (need 1 <= min_width <= max_width, window_width >= 2 * min_width)");
        return PP_ERR_ARG;
    }
    if (n_events > 0x7fffffff) { set_error("more than 2^31 events in one call"); return PP_ERR_ARG; }
    std::lock_guard<std::mutex> lock(g_seg_mutex);
    SegCtx* C = nullptr;
    int rc = seg_ctx(device, &C);
    if (rc != PP_OK) return rc;
    PP_CUDA(cudaSetDevice(device));
    cudaStream_t s = C->stream;
    const int ne = static_cast<int>(n_events);
    *seg_total = 0;
    if (ne == 0) {
        const int64_t zero = 0;
        if (mem == PP_MEM_HOST) seg_offsets[0] = 0;
        else PP_CUDA(cudaMemcpy(seg_offsets, &zero, 8, cudaMemcpyHostToDevice));
        return PP_OK;
    }

    // offsets on the host (validation, output offsets) and on the device
    std::vector<int64_t> h_off(static_cast<size_t>(ne) + 1);
    const int64_t* d_off = offsets;
    if (mem == PP_MEM_HOST) {
        std::memcpy(h_off.data(), offsets, h_off.size() * 8);
        PP_CUDA(C->d_off.reserve(h_off.size() * 8));
        PP_CUDA(cudaMemcpyAsync(C->d_off.p, h_off.data(), h_off.size() * 8, cudaMemcpyHostToDevice, s));
        d_off = C->d_off.as<int64_t>();
    } else {
        PP_CUDA(cudaMemcpy(h_off.data(), offsets, h_off.size() * 8, cudaMemcpyDeviceToHost));
    }
    if (h_off[0] != 0) { set_error("offsets[0] must be 0"); return PP_ERR_ARG; }
    for (int e = 0; e < ne; ++e) {
        const int64_t len = h_off[e + 1] - h_off[e];
        if (len < 1) { set_error("event %d is empty", e); return PP_ERR_ARG; }
        if (len > 0x7ffffff0) { set_error("event %d too long", e); return PP_ERR_ARG; }
    }
    const int64_t total = h_off[ne];
    const double* d_cur = current;
    if (mem == PP_MEM_HOST) {
        PP_CUDA(C->d_cur.reserve(static_cast<size_t>(total) * 8));
        PP_CUDA(cudaMemcpyAsync(C->d_cur.p, current, static_cast<size_t>(total) * 8, cudaMemcpyHostToDevice, s));
        d_cur = C->d_cur.as<double>();
    }
    PP_CUDA(C->d_c.reserve(static_cast<size_t>(total) * 8));
    PP_CUDA(C->d_c2.reserve(static_cast<size_t>(total) * 8));
    PP_CUDA(C->d_marks.reserve(static_cast<size_t>(2 * (total / prm->min_width) + ne + 1) * 4));       // split lists
    PP_CUDA(C->d_count.reserve(static_cast<size_t>(ne) * 4));
    PP_CUDA(C->d_status.reserve(16));
    PP_CUDA(cudaMemsetAsync(C->d_status.p, 0, 16, s));

    SegParams P;
    P.min_width = static_cast<int>(prm->min_width);
    P.max_width = static_cast<int>(prm->max_width);
    P.window_width = static_cast<int>(prm->window_width);
    P.step = static_cast<int>(prm->window_width / 2);
    P.min_gain = prm->min_gain;
    // reciprocal table in shared memory while it leaves room for >= 5 CTAs per SM
    const bool table = prm->window_width <= 4096;
    const size_t dyn = table ? static_cast<size_t>(prm->window_width + 1) * 8 : 0;
    int per_sm = 4;
    if (table) PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_statsplit<true>, kSegThreads, dyn));
    else PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_statsplit<false>, kSegThreads, dyn));
    const int grid = std::min(ne, C->sm_count * std::max(1, per_sm));
    PP_CUDA(cudaEventRecord(C->ev[0], s));
    k_cumsum<<<(ne + kCsWarps * kCsEvents - 1) / (kCsWarps * kCsEvents), kCsWarps * 32, 0, s>>>(
        d_cur, d_off, ne, C->d_c.as<double>(), C->d_c2.as<double>());
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaEventRecord(C->ev[4], s));
    ++C->launches;
    if (table)
        k_statsplit<true><<<grid, kSegThreads, dyn, s>>>(P, d_off, ne, C->d_c.as<double>(), C->d_c2.as<double>(),
                                                         C->d_marks.as<int32_t>(), C->d_count.as<int32_t>(),
                                                         C->d_status.as<int32_t>());
    else
        k_statsplit<false><<<grid, kSegThreads, 0, s>>>(P, d_off, ne, C->d_c.as<double>(), C->d_c2.as<double>(),
                                                        C->d_marks.as<int32_t>(), C->d_count.as<int32_t>(),
                                                        C->d_status.as<int32_t>());
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaEventRecord(C->ev[1], s));
    ++C->launches;

    // segment counts -> offsets (host scan: 4 bytes per event down, 8 up)
    std::vector<int32_t> h_count(ne);
    int32_t h_status = 0;
    PP_CUDA(cudaMemcpyAsync(h_count.data(), C->d_count.p, static_cast<size_t>(ne) * 4, cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaMemcpyAsync(&h_status, C->d_status.p, 4, cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    if (h_status != 0) {
        set_error("an event needs more than %d pending intervals (recursion too deep for the shared-memory stack) or more "
                  "splits than 2 * (n / min_width)", kSegStack);
        return PP_ERR_UNSUPPORTED;
    }
    std::vector<int64_t> h_segoff(static_cast<size_t>(ne) + 1);
    h_segoff[0] = 0;
    for (int e = 0; e < ne; ++e) h_segoff[e + 1] = h_segoff[e] + h_count[e];
    const int64_t n_seg = h_segoff[ne];
    *seg_total = n_seg;
    int64_t* d_segoff = seg_offsets;
    if (mem == PP_MEM_HOST) {
        std::memcpy(seg_offsets, h_segoff.data(), h_segoff.size() * 8);
        PP_CUDA(C->d_segoff.reserve(h_segoff.size() * 8));
        d_segoff = C->d_segoff.as<int64_t>();
    }
    PP_CUDA(cudaMemcpyAsync(d_segoff, h_segoff.data(), h_segoff.size() * 8, cudaMemcpyHostToDevice, s));
    if (n_seg > capacity || !seg_start) {
        PP_CUDA(cudaStreamSynchronize(s));
        if (!seg_start) return PP_OK;                                  // counting pass only
        set_error("segment buffers hold %lld entries, %lld needed", (long long)capacity, (long long)n_seg);
        return PP_ERR_CAPACITY;
    }
    int32_t* d_start = seg_start;
    double* d_mean = seg_mean;
    double* d_std = seg_std;
    float* d_mean32 = seg_mean_f32;
    if (mem == PP_MEM_HOST) {
        PP_CUDA(C->d_start.reserve(static_cast<size_t>(n_seg) * 4));
        d_start = C->d_start.as<int32_t>();
        if (seg_mean) { PP_CUDA(C->d_mean.reserve(static_cast<size_t>(n_seg) * 8)); d_mean = C->d_mean.as<double>(); }
        if (seg_std) { PP_CUDA(C->d_std.reserve(static_cast<size_t>(n_seg) * 8)); d_std = C->d_std.as<double>(); }
        if (seg_mean_f32) { PP_CUDA(C->d_mean32.reserve(static_cast<size_t>(n_seg) * 4)); d_mean32 = C->d_mean32.as<float>(); }
    }
    PP_CUDA(cudaEventRecord(C->ev[2], s));
    k_segments<<<grid, kSegThreads, 0, s>>>(d_cur, d_off, ne, P.min_width, C->d_marks.as<int32_t>(), d_segoff, d_start, d_mean,
                                            d_std, d_mean32);
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaEventRecord(C->ev[3], s));
    ++C->launches;
    if (mem == PP_MEM_HOST) {
        PP_CUDA(cudaMemcpyAsync(seg_start, d_start, static_cast<size_t>(n_seg) * 4, cudaMemcpyDeviceToHost, s));
        if (seg_mean) PP_CUDA(cudaMemcpyAsync(seg_mean, d_mean, static_cast<size_t>(n_seg) * 8, cudaMemcpyDeviceToHost, s));
        if (seg_std) PP_CUDA(cudaMemcpyAsync(seg_std, d_std, static_cast<size_t>(n_seg) * 8, cudaMemcpyDeviceToHost, s));
        if (seg_mean_f32) PP_CUDA(cudaMemcpyAsync(seg_mean_f32, d_mean32, static_cast<size_t>(n_seg) * 4, cudaMemcpyDeviceToHost, s));
    }
    PP_CUDA(cudaStreamSynchronize(s));
    PP_CUDA(cudaEventElapsedTime(&C->last_ms[2], C->ev[0], C->ev[4]));
    PP_CUDA(cudaEventElapsedTime(&C->last_ms[0], C->ev[4], C->ev[1]));
    PP_CUDA(cudaEventElapsedTime(&C->last_ms[1], C->ev[2], C->ev[3]));
    return PP_OK;
}

// Test hook (not in the public header): counts a[i] / b[i] results of the reciprocal division that differ from IEEE division.
int pp_debug_div_check(int device, const double* a_host, const int32_t* b_host, int64_t n, int64_t* mismatches) {
    if (!a_host || !b_host || !mismatches || n < 0) { set_error("null argument"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(device));
    double* da = nullptr; int* db = nullptr; unsigned long long* dbad = nullptr; unsigned long long bad[2] = {0, ~0ULL};
    PP_CUDA(cudaMalloc(&da, n * 8 + 8)); PP_CUDA(cudaMalloc(&db, n * 4 + 8)); PP_CUDA(cudaMalloc(&dbad, 16));
    PP_CUDA(cudaMemcpy(da, a_host, n * 8, cudaMemcpyHostToDevice));
    PP_CUDA(cudaMemcpy(db, b_host, n * 4, cudaMemcpyHostToDevice));
    PP_CUDA(cudaMemcpy(dbad, bad, 16, cudaMemcpyHostToDevice));
    if (n > 0) k_div_check<<<static_cast<unsigned>((n + 255) / 256), 256>>>(da, db, static_cast<int>(n), dbad);
    PP_CUDA(cudaMemcpy(bad, dbad, 16, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dbad);
    mismatches[0] = static_cast<int64_t>(bad[0]);
    mismatches[1] = static_cast<int64_t>(bad[1]);
    return PP_OK;
}

int pp_statsplit_profile(int device, double* out4) {
    double* out3 = out4;
    if (!out3) { set_error("null argument"); return PP_ERR_ARG; }
    std::lock_guard<std::mutex> lock(g_seg_mutex);
    out3[0] = out3[1] = out3[2] = out3[3] = 0.0;
    for (SegCtx* c : g_seg_ctx)
        if (c->device == device) {
            out3[0] = c->last_ms[0];
            out3[1] = c->last_ms[1];
            out3[2] = static_cast<double>(c->launches);
            out3[3] = c->last_ms[2];
        }
    return PP_OK;
}

}  // extern "C"
