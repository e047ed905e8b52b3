// pp_statsplit.cu -- batched FastStatSplit segmenter for sm_100a: raw current -> breakpoints -> segment means.
//
// The step immediately before the HMM path: PyPore's FastStatSplit.parse
// (/root/reference/PyPore/cparsers.pyx:103-118) builds the cumulative sums of the current and its square,
// splits [0, n) recursively at the position of largest log-variance gain inside a sliding window
// (_best_split_stepwise 157-178, _recursive_split 180-203) and hands the segment means to the HMM
// (peptide_hmm_v0.0.1.py:443-466).  Here one CTA owns one event:
//
//   k_statsplit   phase 1  cumulative sums.  numpy.cumsum is a SEQUENTIAL float64 accumulation and every gain
//                          is a difference of its entries, so the order of the additions is part of the result:
//                          tiles of the current are staged in shared memory with coalesced loads, ONE thread
//                          accumulates c and one thread (of another warp) c2 in the reference's order, and the
//                          tile goes back to HBM/L2 with coalesced stores.  The serial part costs 2 of 256
//                          threads; the other CTAs resident on the SM fill the issue slots meanwhile.
//                 phase 2  the recursion as an explicit interval stack in shared memory.  A window scan is
//                          parallel over the candidate positions: float64 gain with explicitly rounded
//                          operations (no FMA contraction) and a (gain, position) arg-max whose tie rule --
//                          lowest position among equal gains, gain strictly above the threshold -- is what the
//                          reference's ascending strict-'>' scan leaves.  Splits are marked in a byte map, which
//                          makes the output order (ascending = the reference's in-order concatenation) free.
//   k_segments    compacts the byte map into segment starts (block scan) and reduces every segment to mean / std
//                 with one warp per segment (Segment.mean / .std, PyPore/core.py:209-214).
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
constexpr int kSegTile = 512;          // samples staged per cumulative-sum tile
constexpr int kSegStack = 1024;        // pending intervals per event (depth-first: bounded by the recursion depth)

struct SegParams {
    int min_width, max_width, window_width, step;
    double min_gain;
};

__device__ __forceinline__ bool better(double g, int i, double bg, int bi) {
    return g > bg || (g == bg && i >= 0 && i < bi);
}

// cparsers.pyx:157-178 -- all threads of the CTA call this with the same arguments and get the same answer
__device__ int best_split_stepwise(const SegParams& P, const double* __restrict__ c, const double* __restrict__ c2,
                                   int start, int end, double* s_gain, int* s_idx) {
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
        const double ml = __ddiv_rn(__dsub_rn(ci, cb), nl);
        const double vl = __dsub_rn(__ddiv_rn(__dsub_rn(c2i, c2b), nl), __dmul_rn(ml, ml));
        const double mh = __ddiv_rn(__dsub_rn(ce, ci), nh);
        const double vh = __dsub_rn(__ddiv_rn(__dsub_rn(c2e, c2i), nh), __dmul_rn(mh, mh));
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

__global__ void __launch_bounds__(kSegThreads)
k_statsplit(const SegParams P, const double* __restrict__ current, const int64_t* __restrict__ offsets, int n_events,
            double* __restrict__ cs, double* __restrict__ cs2, unsigned char* __restrict__ marks,
            int32_t* __restrict__ seg_count, int32_t* __restrict__ status) {
    __shared__ double s_x[kSegTile], s_c[kSegTile], s_c2[kSegTile];
    __shared__ int2 s_stack[kSegStack];
    __shared__ double s_gain[kSegThreads / 32];
    __shared__ int s_idx[kSegThreads / 32];
    __shared__ int s_sp, s_fail;
    const int tid = threadIdx.x;
    for (int ev = blockIdx.x; ev < n_events; ev += gridDim.x) {
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        const double* x = current + off;
        double* c = cs + off;
        double* c2 = cs2 + off;
        unsigned char* mk = marks + off;

        // ---- phase 1: c = cumsum(x), c2 = cumsum(x * x), sequential float64 (cparsers.pyx:110-111)
        double run = 0.0;                   // thread 0 carries c, thread 32 carries c2
        for (int t0 = 0; t0 < n; t0 += kSegTile) {
            const int cnt = min(kSegTile, n - t0);
            for (int k = tid; k < cnt; k += kSegThreads) s_x[k] = x[t0 + k];
            __syncthreads();
            if (tid == 0) {
                double s = run;
                int k = 0;
                if (t0 == 0) { s = s_x[0]; s_c[0] = s; k = 1; }
#pragma unroll 8
                for (; k < cnt; ++k) { s = __dadd_rn(s, s_x[k]); s_c[k] = s; }
                run = s;
            } else if (tid == 32) {
                double s = run;
                int k = 0;
                if (t0 == 0) { s = __dmul_rn(s_x[0], s_x[0]); s_c2[0] = s; k = 1; }
#pragma unroll 8
                for (; k < cnt; ++k) { s = __dadd_rn(s, __dmul_rn(s_x[k], s_x[k])); s_c2[k] = s; }
                run = s;
            }
            __syncthreads();
            for (int k = tid; k < cnt; k += kSegThreads) { c[t0 + k] = s_c[k]; c2[t0 + k] = s_c2[k]; }
        }
        if (tid == 0) { s_stack[0] = make_int2(0, n); s_sp = 1; s_fail = 0; }
        int n_splits = 0;                   // thread 0 only
        __syncthreads();                    // the sums are visible to the whole CTA from here on

        // ---- phase 2: _recursive_split (180-203) with an explicit stack, left part first
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
                split_at = best_split_stepwise(P, c, c2, static_cast<int>(ps), pe, s_gain, s_idx);
                if (split_at >= 0) break;
            }
            if (split_at == -1 && static_cast<long long>(end) - start > P.max_width)          // 198-201
                split_at = static_cast<int>(min(static_cast<long long>(start) + P.max_width,
                                                static_cast<long long>(end) - P.min_width));
            if (tid == 0) {
                int top = sp - 1;
                if (split_at >= 0) {
                    if (split_at <= start || split_at >= end || top + 2 > kSegStack) {
                        s_fail = 1;                                  // stack exhausted or a split that cannot terminate
                        top = 0;
                    } else {
                        mk[split_at] = 1;
                        ++n_splits;
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

// byte map -> segment starts, then mean / std of every segment (one warp per segment)
__global__ void __launch_bounds__(kSegThreads)
k_segments(const double* __restrict__ current, const int64_t* __restrict__ offsets, int n_events,
           const unsigned char* __restrict__ marks, const int64_t* __restrict__ seg_offsets,
           int32_t* __restrict__ seg_start, double* __restrict__ seg_mean, double* __restrict__ seg_std,
           float* __restrict__ seg_mean_f32) {
    __shared__ int s_warp[kSegThreads / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int ev = blockIdx.x; ev < n_events; ev += gridDim.x) {
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        const int64_t so = seg_offsets[ev];
        const int n_seg = static_cast<int>(seg_offsets[ev + 1] - so);
        const unsigned char* mk = marks + off;
        if (tid == 0) { s_base = 1; seg_start[so] = 0; }
        __syncthreads();
        for (int t0 = 0; t0 < n; t0 += kSegThreads) {
            const int i = t0 + tid;
            const bool m = i < n && mk[i] != 0;
            const unsigned b = __ballot_sync(0xffffffffu, m);
            if (lane == 0) s_warp[warp] = __popc(b);
            __syncthreads();
            int before = s_base;
            for (int w = 0; w < warp; ++w) before += s_warp[w];
            if (m) seg_start[so + before + __popc(b & ((1u << lane) - 1u))] = i;
            __syncthreads();
            if (tid == 0) {
                int t = 0;
                for (int w = 0; w < kSegThreads / 32; ++w) t += s_warp[w];
                s_base += t;
            }
            __syncthreads();
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
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int sm_count = 0;
    DevBuf d_cur, d_off, d_c, d_c2, d_marks, d_count, d_status, d_segoff, d_start, d_mean, d_std, d_mean32;
    int64_t launches = 0;
    float last_ms[2] = {0.f, 0.f};
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
    PP_CUDA(C->d_marks.reserve(static_cast<size_t>(total)));
    PP_CUDA(C->d_count.reserve(static_cast<size_t>(ne) * 4));
    PP_CUDA(C->d_status.reserve(16));
    PP_CUDA(cudaMemsetAsync(C->d_marks.p, 0, static_cast<size_t>(total), s));
    PP_CUDA(cudaMemsetAsync(C->d_status.p, 0, 16, s));

    SegParams P;
    P.min_width = static_cast<int>(prm->min_width);
    P.max_width = static_cast<int>(prm->max_width);
    P.window_width = static_cast<int>(prm->window_width);
    P.step = static_cast<int>(prm->window_width / 2);
    P.min_gain = prm->min_gain;
    const int grid = std::min(ne, C->sm_count * 8);
    PP_CUDA(cudaEventRecord(C->ev[0], s));
    k_statsplit<<<grid, kSegThreads, 0, s>>>(P, d_cur, d_off, ne, C->d_c.as<double>(), C->d_c2.as<double>(),
                                             C->d_marks.as<unsigned char>(), C->d_count.as<int32_t>(),
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
        set_error("an event needs more than %d pending intervals (recursion too deep for the shared-memory stack)", kSegStack);
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
    k_segments<<<grid, kSegThreads, 0, s>>>(d_cur, d_off, ne, C->d_marks.as<unsigned char>(), d_segoff, d_start, d_mean,
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
    PP_CUDA(cudaEventElapsedTime(&C->last_ms[0], C->ev[0], C->ev[1]));
    PP_CUDA(cudaEventElapsedTime(&C->last_ms[1], C->ev[2], C->ev[3]));
    return PP_OK;
}

int pp_statsplit_profile(int device, double* out3) {
    if (!out3) { set_error("null argument"); return PP_ERR_ARG; }
    std::lock_guard<std::mutex> lock(g_seg_mutex);
    out3[0] = out3[1] = out3[2] = 0.0;
    for (SegCtx* c : g_seg_ctx)
        if (c->device == device) {
            out3[0] = c->last_ms[0];
            out3[1] = c->last_ms[1];
            out3[2] = static_cast<double>(c->launches);
        }
    return PP_OK;
}

}  // extern "C"
