// pp_segalign.cu -- batched cSegmentAligner: the Karplus semi-local alignment of (mean, std, duration)
// segments to a model with skip and back-slip penalties, /root/reference/PyPore/calignment.pyx:20-101 (caller:
// PyPore/alignment.py:26-47).
//
// Like the HMM sweeps this is a per-sequence dynamic programme whose two running maxima -- the skip chain (ascending
// model position) and the back-slip chain (descending) -- are first-order recurrences max(chain, score) - penalty that
// cannot be re-associated without changing bits, and the backtrace DECIDES by comparing float64 values with a 1e-6
// relative tolerance.  So the mapping is the lane-per-sequence one: a warp takes 32 sequences (length-sorted), lane l
// walks sequence l with exactly the reference's float64 operations in the reference's order (explicitly rounded, no
// FMA contraction: the reference is compiled for baseline x86-64), and the three tables the backtrace reads -- score,
// skip_score, backslip_score -- live in HBM as [row][model position][lane], so every access of a warp is one coalesced
// 256-byte line.  Parallelism comes from the batch: 32 sequences per warp, thousands of warps.
//
// Two places where the reference is undefined are DEFINED here and reported per sequence (status bits, see
// protpore_b200.h); the CPU restatement oracle/segalign_oracle.c defines them identically.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

#include "pp_common.h"

namespace pp {

constexpr double kSegNegInf = -99999999.0;            // calignment.pyx:38

static double g_seg_profile[4] = {0, 0, 0, 0};        // last call: kernel ms, launches, workspace bytes, groups

struct SegCache {
    std::mutex mu;
    DevBuf b[12];
    bool keep = true;                                 // PP_SEGALIGN_CACHE=0 frees everything at the end of a call
};
static SegCache& seg_cache(int device) {
    static SegCache caches[64];
    static const bool keep = [] { const char* e = std::getenv("PP_SEGALIGN_CACHE"); return !(e && std::atoi(e) == 0); }();
    SegCache& c = caches[device < 0 || device >= 64 ? 0 : device];
    c.keep = keep;
    return c;
}

struct SegAlignArgs {
    const double* model;          // [4][m]: means, stds, durs, cumulative durs
    int32_t m;
    double skip_penalty, backslip_penalty;
    const double *seq_means, *seq_stds, *seq_durs;
    const int64_t* offsets;
    const int32_t* order;         // [n_groups * 32] sequence per (group, lane), -1 = padding
    const int64_t* group_ws;      // [n_groups] first double of the group's tables
    const int32_t* group_len;     // [n_groups] longest sequence
    double* ws;
    double* score_out;
    int32_t* order_out;
    int32_t* status_out;
    int32_t n_groups;
};

// numpy.add.reduce over float64 (pairwise summation): what numpy.sum(seq_durs) returns (calignment.pyx:100)
__device__ double seg_pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] = __dadd_rn(r[k], a[i + k]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(seg_pairwise_sum(a, n2), seg_pairwise_sum(a + n2, n - n2));
}

__device__ __forceinline__ double seg_max(double a, double b) { return a >= b ? a : b; }      // double_max, 8

__global__ void __launch_bounds__(128) k_segment_align(const SegAlignArgs A) {
    extern __shared__ double s_model[];                       // [4][m]
    const int m = A.m;
    for (int i = threadIdx.x; i < 4 * m; i += blockDim.x) s_model[i] = A.model[i];
    __syncthreads();
    const double *mean = s_model, *stdv = s_model + m, *dur = s_model + 2 * m, *cdur = s_model + 3 * m;
    const int lane = threadIdx.x & 31;
    // the previous score row of the warp's 32 sequences stays in shared memory ([position][lane]): the chains read it at
    // shared-memory latency instead of waiting for DRAM (the rows of all resident warps exceed the L2)
    double* const prow = s_model + 4 * m + static_cast<size_t>(threadIdx.x >> 5) * 2 * m * 32 + lane;
    double* const brow = prow + m * 32;                       // the back-slip chain of the current row, read back by the ascending sweep
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= A.n_groups) return;
    const int sq = A.order[g * 32 + lane];
    if (sq < 0) return;
    const int64_t off = A.offsets[sq];
    const int s = static_cast<int>(A.offsets[sq + 1] - off);
    const int smax = A.group_len[g];
    const double *xm = A.seq_means + off, *xs = A.seq_stds + off, *xd = A.seq_durs + off;
    double* T = A.ws + A.group_ws[g] + lane;                  // [table][row][position][lane]
    const size_t tstride = static_cast<size_t>(smax) * m * 32;
#define S_(i, j) T[(static_cast<size_t>(i) * m + (j)) * 32]
#define K_(i, j) T[tstride + (static_cast<size_t>(i) * m + (j)) * 32]
#define B_(i, j) T[2 * tstride + (static_cast<size_t>(i) * m + (j)) * 32]
    const double sp = A.skip_penalty, bp = A.backslip_penalty;
    auto match = [&](int i, int j) {                          // 47: -(x - mu)^2 / (std_x * std_model)
        const double d = __dsub_rn(xm[i], mean[j]);
        return __ddiv_rn(-__dmul_rn(d, d), __dmul_rn(xs[i], stdv[j]));
    };
    {
        const double d0 = xd[0];
        for (int j = 0; j < m; ++j) {
            const double v = __dsub_rn(__dmul_rn(match(0, j), d0), __dmul_rn(sp, __dsub_rn(cdur[j], dur[j])));   // 50
            S_(0, j) = v;
            prow[j * 32] = v;
        }
    }
    for (int i = 1; i < s; ++i) {
        const double di = xd[i], xmi = xm[i], xsi = xs[i];
        // disjoint rows: the compiler may hoist loads above the stores of row i (through one base pointer every link of the
        // chains waited for a memory round trip)
        double* __restrict__ Sc = &S_(i, 0);
        double* __restrict__ Kc = &K_(i, 0);
        double* __restrict__ Bc = &B_(i, 0);
        double c = kSegNegInf;
        Bc[static_cast<size_t>(m - 1) * 32] = c;
        brow[(m - 1) * 32] = c;
#pragma unroll 8
        for (int j = m - 2; j >= 0; --j) {                     // 58-59: the back-slip chain
            c = __dsub_rn(seg_max(c, prow[(j + 1) * 32]), __dmul_rn(dur[j + 1], bp));
            Bc[static_cast<size_t>(j) * 32] = c;
            brow[j * 32] = c;
        }
        // 55 and 61-67 in one ascending sweep: K(i, j - 1) and S(i - 1, j - 1) are carried in registers
        auto match_i = [&](int j) {
            const double d = __dsub_rn(xmi, mean[j]);
            return __ddiv_rn(-__dmul_rn(d, d), __dmul_rn(xsi, stdv[j]));
        };
        double kc = kSegNegInf, sjm1 = prow[0];
        Kc[0] = kc;
        {
            double prev = sjm1;
            if (0 < m - 1) prev = seg_max(prev, brow[0]);
            const double v = __dadd_rn(prev, __dmul_rn(match_i(0), di));
            Sc[0] = v;
            prow[0] = v;
        }
#pragma unroll 8
        for (int j = 1; j < m; ++j) {
            const double sj = prow[j * 32];
            double prev = sj;
            if (sjm1 > prev) prev = sjm1;
            if (kc > prev) prev = kc;
            if (j < m - 1) prev = seg_max(prev, brow[j * 32]);
            const double v = __dadd_rn(prev, __dmul_rn(match_i(j), di));
            Sc[static_cast<size_t>(j) * 32] = v;
            prow[j * 32] = v;                                                   // row i replaces row i - 1 in place (sjm1 keeps the old value)
            kc = __dsub_rn(seg_max(kc, sjm1), __dmul_rn(dur[j], sp));           // the skip chain: K(i, j)
            Kc[static_cast<size_t>(j) * 32] = kc;
            sjm1 = sj;
        }
    }
    int status = 0;
    int j = -1;
    {   // double_argmax of the last row (11-18)
        double maximum = -1;
        for (int t = 0; t < m; ++t) {
            const double v = S_(s - 1, t);
            if (v > maximum) { j = t; maximum = v; }
        }
        if (j < 0) {                                           // undefined in the reference: the first maximum
            status |= 1;
            j = 0;
            double best = S_(s - 1, 0);
            for (int t = 1; t < m; ++t) {
                const double v = S_(s - 1, t);
                if (v > best) { best = v; j = t; }
            }
        }
    }
    int32_t* out = A.order_out + off;
    for (int i = s - 1; i > 0; --i) {                           // 72-97
        out[i] = j;
        const double prev_test = __dsub_rn(S_(i, j), __dmul_rn(match(i, j), xd[i]));
        const double max_err = __dmul_rn(1e-6, fabs(prev_test));
        if (j == 0) status |= 2;                               // the reference reads score[i-1, -1u] here and raises
        if (j > 0 && fabs(__dsub_rn(prev_test, S_(i - 1, j - 1))) <= max_err) { j -= 1; continue; }
        if (fabs(__dsub_rn(prev_test, S_(i - 1, j))) <= max_err) continue;
        if (j < m - 1) {
            int k = j;
            double back_test = prev_test;
            while (k < m - 1 && fabs(__dsub_rn(back_test, B_(i, k))) <= max_err) {
                k += 1;
                back_test = __dadd_rn(back_test, __dmul_rn(dur[k], bp));
            }
            if (k > j) { j = k; continue; }
        }
        if (j > 0) {
            int k = j;
            double skip_test = prev_test;
            while (k >= 1 && fabs(__dsub_rn(skip_test, K_(i, k - 1))) <= max_err) {
                k -= 1;
                skip_test = __dadd_rn(skip_test, __dmul_rn(dur[k], sp));
            }
            if (k < j) {
                if (k == 0) { status |= 2; j = 0; } else j = k - 1;
                continue;
            }
        }
    }
    out[0] = j;
    A.score_out[sq] = __ddiv_rn(S_(s - 1, m - 1), seg_pairwise_sum(xd, s));      // 100
    A.status_out[sq] = status;
#undef S_
#undef K_
#undef B_
}

}  // namespace pp

using namespace pp;

extern "C" int pp_segment_align_batch(const pp_segment_model* model, int device, const double* seq_means, const double* seq_stds,
                                      const double* seq_durs, const int64_t* offsets, int64_t n_seqs, int32_t mem, double* score_out,
                                      int32_t* order_out, int32_t* status_out) {
    if (!model || !model->means || !model->stds || !model->durs || !seq_means || !seq_stds || !seq_durs || !offsets || !score_out ||
        !order_out || !status_out) { set_error("null argument"); return PP_ERR_ARG; }
    if (model->n < 2) { set_error("the model needs at least two segments (calignment.pyx:59 indexes position 1)"); return PP_ERR_ARG; }
    if (mem != PP_MEM_HOST && mem != PP_MEM_DEVICE) { set_error("bad mem"); return PP_ERR_ARG; }
    if (n_seqs < 0 || n_seqs > 0x7fffffff - 64) { set_error("bad n_seqs"); return PP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device: protpore_b200 has no CPU fallback"); return PP_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return PP_ERR_ARG; }
    if (n_seqs == 0) return PP_OK;
    PP_CUDA(cudaSetDevice(device));
    const int m = model->n;
    const bool host = mem == PP_MEM_HOST;
    std::vector<int64_t> h_off(static_cast<size_t>(n_seqs) + 1);
    if (host) std::memcpy(h_off.data(), offsets, h_off.size() * sizeof(int64_t));
    else PP_CUDA(cudaMemcpy(h_off.data(), offsets, h_off.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < n_seqs; ++e)
        if (h_off[e + 1] - h_off[e] < 1) { set_error("sequence %lld has no segments", (long long)e); return PP_ERR_ARG; }
    const int64_t total = h_off[n_seqs] - h_off[0];

    // model: means | stds | durs | numpy.cumsum(durs) (sequential float64 sums, calignment.pyx:30)
    std::vector<double> hm(4 * static_cast<size_t>(m));
    double acc = 0.0;
    for (int j = 0; j < m; ++j) {
        hm[j] = model->means[j]; hm[m + j] = model->stds[j]; hm[2 * m + j] = model->durs[j];
        acc += model->durs[j];
        hm[3 * m + j] = acc;
    }
    // sequences by length, longest first, 32 per warp
    std::vector<int32_t> idx(static_cast<size_t>(n_seqs));
    {   // stable, descending by length: a counting sort when the lengths are small numbers (they are: segments per event)
        int64_t longest = 0;
        for (int64_t e = 0; e < n_seqs; ++e) longest = std::max(longest, h_off[e + 1] - h_off[e]);
        if (longest <= (int64_t(1) << 22)) {
            std::vector<int64_t> first(static_cast<size_t>(longest) + 2, 0);
            for (int64_t e = 0; e < n_seqs; ++e) ++first[static_cast<size_t>(longest - (h_off[e + 1] - h_off[e])) + 1];
            for (size_t k = 1; k < first.size(); ++k) first[k] += first[k - 1];
            for (int64_t e = 0; e < n_seqs; ++e) idx[static_cast<size_t>(first[static_cast<size_t>(longest - (h_off[e + 1] - h_off[e]))]++)] = static_cast<int32_t>(e);
        } else {
            std::iota(idx.begin(), idx.end(), 0);
            std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return h_off[a + 1] - h_off[a] > h_off[b + 1] - h_off[b]; });
        }
    }
    const int32_t ng = static_cast<int32_t>((n_seqs + 31) / 32);
    std::vector<int32_t> order(static_cast<size_t>(ng) * 32, -1), glen(ng);
    std::copy(idx.begin(), idx.end(), order.begin());
    for (int32_t g = 0; g < ng; ++g) glen[g] = static_cast<int32_t>(h_off[idx[static_cast<size_t>(g) * 32] + 1] - h_off[idx[static_cast<size_t>(g) * 32]]);

    // device buffers live in a per-device cache between calls (the aligner has no handle: PyPore builds one per call);
    // cudaMalloc / cudaFree of the 30 GB of tables cost more than the kernel.  Calls on one device take turns.
    SegCache& C = seg_cache(device);
    std::lock_guard<std::mutex> cache_lock(C.mu);
    DevBuf &d_model = C.b[0], &d_means = C.b[1], &d_stds = C.b[2], &d_durs = C.b[3], &d_off = C.b[4], &d_order = C.b[5], &d_gws = C.b[6],
           &d_glen = C.b[7], &d_ws = C.b[8], &d_score = C.b[9], &d_ord = C.b[10], &d_status = C.b[11];
    auto release = [&] { if (!C.keep) for (DevBuf& b : C.b) b.release(); };
#define SEG_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error("%s failed: %s", #x, cudaGetErrorString(e_)); release(); return PP_ERR_CUDA; } } while (0)
    SEG_CUDA(upload(&d_model, hm, nullptr));
    SEG_CUDA(upload(&d_order, order, nullptr));
    SEG_CUDA(upload(&d_glen, glen, nullptr));
    const double *p_means = seq_means, *p_stds = seq_stds, *p_durs = seq_durs;
    const int64_t* p_off = offsets;
    double* p_score = score_out;
    int32_t *p_ord = order_out, *p_status = status_out;
    if (host) {
        const size_t nb = static_cast<size_t>(total) * sizeof(double);
        SEG_CUDA(d_means.reserve(nb)); SEG_CUDA(d_stds.reserve(nb)); SEG_CUDA(d_durs.reserve(nb));
        SEG_CUDA(cudaMemcpy(d_means.p, seq_means + h_off[0], nb, cudaMemcpyHostToDevice));
        SEG_CUDA(cudaMemcpy(d_stds.p, seq_stds + h_off[0], nb, cudaMemcpyHostToDevice));
        SEG_CUDA(cudaMemcpy(d_durs.p, seq_durs + h_off[0], nb, cudaMemcpyHostToDevice));
        SEG_CUDA(upload(&d_off, h_off, nullptr));
        SEG_CUDA(d_score.reserve(n_seqs * sizeof(double))); SEG_CUDA(d_ord.reserve(total * sizeof(int32_t))); SEG_CUDA(d_status.reserve(n_seqs * sizeof(int32_t)));
        p_means = d_means.as<double>() - h_off[0]; p_stds = d_stds.as<double>() - h_off[0]; p_durs = d_durs.as<double>() - h_off[0];
        p_off = d_off.as<int64_t>();
        p_score = d_score.as<double>(); p_ord = d_ord.as<int32_t>() - h_off[0]; p_status = d_status.as<int32_t>();
    }
    // the three tables of a group take 3 x rows x m x 32 doubles: launch as many groups as fit the workspace
    size_t fr = 0, tot = 0;
    SEG_CUDA(cudaMemGetInfo(&fr, &tot));
    const size_t limit = std::max<size_t>(size_t(256) << 20, static_cast<size_t>((fr + d_ws.cap) * 0.6)) / sizeof(double);   // the cached tables count as free
    // shared memory: the model and one score row (m x 32 doubles) per warp; as many warps per block (<= 4) as fit
    struct { size_t sharedMemPerBlockOptin; } prop;
    {
        int optin = 0;
        SEG_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        prop.sharedMemPerBlockOptin = static_cast<size_t>(optin);
    }
    const size_t row_bytes = 2 * static_cast<size_t>(m) * 32 * sizeof(double), model_bytes = 4 * static_cast<size_t>(m) * sizeof(double);
    if (model_bytes + row_bytes > prop.sharedMemPerBlockOptin) { set_error("model of %d segments does not fit shared memory", m); release(); return PP_ERR_UNSUPPORTED; }
    // blocks of one warp pack an SM best (16 KB each for m = 64); four warps share one copy of the model when rows are small
    const int wpb = row_bytes >= 4096 ? 1 : static_cast<int>(std::min<size_t>(4, (prop.sharedMemPerBlockOptin - model_bytes) / row_bytes));
    const size_t smem = model_bytes + wpb * row_bytes;
    cudaError_t e = cudaFuncSetAttribute(k_segment_align, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("model of %d segments does not fit shared memory", m); release(); return PP_ERR_UNSUPPORTED; }
    std::vector<int64_t> gws(ng);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    double kernel_ms = 0.0, ws_bytes = 0.0;
    int launches = 0;
    for (int32_t g0 = 0; g0 < ng;) {
        size_t used = 0;
        int32_t g1 = g0;
        while (g1 < ng) {
            const size_t need = static_cast<size_t>(3) * glen[g1] * m * 32;
            if (g1 > g0 && used + need > limit) break;
            gws[g1] = static_cast<int64_t>(used);
            used += need;
            ++g1;
        }
        if (used > limit && g1 == g0 + 1 && used * sizeof(double) > fr) { set_error("one sequence needs %zu bytes of workspace", used * sizeof(double)); release(); return PP_ERR_NOMEM; }
        SEG_CUDA(d_ws.reserve(used * sizeof(double)));
        SEG_CUDA(upload(&d_gws, gws, nullptr));
        SegAlignArgs A{d_model.as<double>(), m, model->skip_penalty, model->backslip_penalty, p_means, p_stds, p_durs, p_off,
                       d_order.as<int32_t>() + static_cast<size_t>(g0) * 32, d_gws.as<int64_t>() + g0, d_glen.as<int32_t>() + g0,
                       d_ws.as<double>(), p_score, p_ord, p_status, g1 - g0};
        if (ev0) cudaEventRecord(ev0, 0);
        k_segment_align<<<(g1 - g0 + wpb - 1) / wpb, wpb * 32, smem>>>(A);
        if (ev1) cudaEventRecord(ev1, 0);
        SEG_CUDA(cudaGetLastError());
        SEG_CUDA(cudaDeviceSynchronize());
        float ms = 0.f;
        if (ev0 && ev1 && cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) kernel_ms += ms;
        ++launches;
        ws_bytes = std::max(ws_bytes, static_cast<double>(used) * sizeof(double));
        g0 = g1;
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    g_seg_profile[0] = kernel_ms; g_seg_profile[1] = launches; g_seg_profile[2] = ws_bytes; g_seg_profile[3] = ng;
    if (host) {
        SEG_CUDA(cudaMemcpy(score_out, d_score.p, n_seqs * sizeof(double), cudaMemcpyDeviceToHost));
        SEG_CUDA(cudaMemcpy(status_out, d_status.p, n_seqs * sizeof(int32_t), cudaMemcpyDeviceToHost));
        SEG_CUDA(cudaMemcpy(order_out + h_off[0], d_ord.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    release();
    return PP_OK;
#undef SEG_CUDA
}

extern "C" int pp_segment_align_profile(double* out4) {
    if (!out4) { pp::set_error("null argument"); return PP_ERR_ARG; }
    for (int i = 0; i < 4; ++i) out4[i] = pp::g_seg_profile[i];
    return PP_OK;
}

extern "C" int pp_segment_align_trim(int device) {
    pp::SegCache& C = pp::seg_cache(device);
    std::lock_guard<std::mutex> lock(C.mu);
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); pp::set_error("bad device"); return PP_ERR_ARG; }
    for (pp::DevBuf& b : C.b) b.release();
    return PP_OK;
}
