// pp_api.cu -- C ABI of protpore_b200 (include/protpore_b200.h): model handle, event bucketing,
// batched Viterbi + traceback and batched forward over the lane-per-event kernels, event sampler.
// There is no CPU fallback anywhere in this file: without a CUDA device every entry point that
// computes returns PP_ERR_CUDA.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#define PP_WITH_TRACEBACK_KERNELS                   // pp_device_common.cuh: k_traceback* are instantiated here only
#include "pp_lane_types.h"
#include "pp_model.h"
#include "pp_profile_sched.h"

namespace pp {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

int profile_viterbi_launch(const ProfSched& S, const LaneBatchAny& B, const VitOut& O, int n_groups, cudaStream_t stream);
int profile_forward_launch(const ProfSched& S, const LaneBatchAny& B, const FwdOut& O, int n_groups, cudaStream_t stream);
void lane_viterbi4_invalidate(const void*);
void lane_viterbi4g_invalidate(const void*);
void lane_forwardg_invalidate(const void*);
void lane_forward_invalidate(const void*);
void lane_fb_invalidate(const void*);
void lane_arena_invalidate(const void* owner) {
    lane_fb_invalidate(owner);
    lane_viterbi4_invalidate(owner);
    lane_viterbi4g_invalidate(owner);
    lane_forwardg_invalidate(owner);
    lane_forward_invalidate(owner);
}

// ---- event bucketing on the device ----------------------------------------------------------------------------
// The same schedule build_schedule makes on the host -- events sorted by length, longest first, stable; 32 per group;
// group g owns group_len[g] + 1 rows of the pointer table from group_row[g] -- built with a CUB radix sort and a scan, so
// a batch of a million events costs ~0.3 ms of GPU time instead of ~8 ms of host time in front of every sweep.
static __global__ void k_event_lengths(const int64_t* __restrict__ off, int64_t e0, int64_t n, uint32_t* __restrict__ len,
                                       int32_t* __restrict__ idx, long long* __restrict__ minmax) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    long long lo = 0x7fffffffffffffffLL, hi = -0x7fffffffffffffffLL;
    if (i < n) {
        const long long l = off[e0 + i + 1] - off[e0 + i];
        len[i] = static_cast<uint32_t>(l < 0 ? 0 : (l > 0x7ffffff0LL ? 0x7ffffff0LL : l));
        idx[i] = static_cast<int32_t>(e0 + i);
        lo = hi = l;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0 && hi >= lo) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}
static __global__ void k_schedule_groups(const uint32_t* __restrict__ len_sorted, const int32_t* __restrict__ idx_sorted, int64_t n,
                                         int32_t n_groups, int32_t* __restrict__ order, int32_t* __restrict__ group_len,
                                         int64_t* __restrict__ group_rows) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < static_cast<int64_t>(n_groups) * kLanes) order[i] = i < n ? idx_sorted[i] : -1;
    if (i < n_groups) {
        const int32_t l = static_cast<int32_t>(len_sorted[i * kLanes]);
        group_len[i] = l;
        group_rows[i] = l + 1;
    }
}
static __global__ void k_schedule_total(const int64_t* __restrict__ group_row, const int32_t* __restrict__ group_len, int32_t n_groups,
                                        long long* __restrict__ minmax) {
    minmax[2] = group_row[n_groups - 1] + group_len[n_groups - 1] + 1;
}
static __global__ void k_first_bad_event(const int64_t* __restrict__ off, int64_t n, long long* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        const long long l = off[i + 1] - off[i];
        if (l < 1 || l > 0x7ffffff0LL) atomicMin(out, static_cast<long long>(i));
    }
}

struct DevSchedule {
    int32_t n_groups = 0;
    cudaEvent_t ready = nullptr;       // the read-back of {min length, max length, total rows} has landed in h_pin
};
static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
// Enqueues the bucketing of events [e0, e1) into the buffers of pipeline slot `slot` on `stream`; the three scalars go to
// h_pin[16 + 4 slot ...].  Nothing is synchronised here.
static int device_schedule_enqueue(pp_model* M, int slot, const int64_t* d_off, int64_t e0, int64_t e1, cudaStream_t stream,
                                   DevSchedule* out) {
    pp_model::PipeSlot& P = M->ps[slot];
    const int64_t n = e1 - e0;
    const int32_t ng = static_cast<int32_t>((n + kLanes - 1) / kLanes);
    out->n_groups = ng;
    size_t sort_tmp = 0, scan_tmp = 0;
    PP_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_tmp, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                                      static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), static_cast<int>(n), 0, 32, stream));
    PP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, static_cast<const int64_t*>(nullptr), static_cast<int64_t*>(nullptr), ng, stream));
    const size_t b_len = align256(static_cast<size_t>(n) * 4), b_grp = align256(static_cast<size_t>(ng) * 8), b_tmp = align256(std::max(sort_tmp, scan_tmp));
    PP_CUDA(P.sched.reserve(4 * b_len + b_grp + b_tmp + 256));
    unsigned char* base = P.sched.as<unsigned char>();
    uint32_t* len = reinterpret_cast<uint32_t*>(base);
    uint32_t* len_sorted = reinterpret_cast<uint32_t*>(base + b_len);
    int32_t* idx = reinterpret_cast<int32_t*>(base + 2 * b_len);
    int32_t* idx_sorted = reinterpret_cast<int32_t*>(base + 3 * b_len);
    int64_t* rows = reinterpret_cast<int64_t*>(base + 4 * b_len);
    void* tmp = base + 4 * b_len + b_grp;
    long long* scal = reinterpret_cast<long long*>(base + 4 * b_len + b_grp + b_tmp);
    PP_CUDA(P.order.reserve(static_cast<size_t>(ng) * kLanes * sizeof(int32_t) + 16));
    PP_CUDA(P.group_len.reserve(static_cast<size_t>(ng) * sizeof(int32_t) + 16));
    PP_CUDA(P.group_row.reserve(static_cast<size_t>(ng) * sizeof(int64_t) + 16));
    const long long init[4] = {0x7fffffffffffffffLL, -0x7fffffffffffffffLL, 0, 0};
    PP_CUDA(cudaMemcpyAsync(scal, init, sizeof init, cudaMemcpyHostToDevice, stream));
    const unsigned nb = static_cast<unsigned>((n + 255) / 256);
    k_event_lengths<<<nb, 256, 0, stream>>>(d_off, e0, n, len, idx, scal);
    PP_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, sort_tmp, len, len_sorted, idx, idx_sorted, static_cast<int>(n), 0, 32, stream));
    k_schedule_groups<<<static_cast<unsigned>((static_cast<int64_t>(ng) * kLanes + 255) / 256), 256, 0, stream>>>(
        len_sorted, idx_sorted, n, ng, P.order.as<int32_t>(), P.group_len.as<int32_t>(), rows);
    PP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, rows, P.group_row.as<int64_t>(), ng, stream));
    k_schedule_total<<<1, 1, 0, stream>>>(P.group_row.as<int64_t>(), P.group_len.as<int32_t>(), ng, scal);
    PP_CUDA(cudaGetLastError());
    M->launches += 5;
    PP_CUDA(cudaMemcpyAsync(M->h_pin + 16 + 4 * slot, scal, 3 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    return PP_OK;
}
static int bad_event_error(pp_model* M, const int64_t* d_off, int64_t n_events, cudaStream_t stream) {
    long long* d = nullptr;
    long long first = 0x7fffffffffffffffLL;
    if (cudaMalloc(&d, sizeof first) == cudaSuccess) {
        cudaMemcpyAsync(d, &first, sizeof first, cudaMemcpyHostToDevice, stream);
        k_first_bad_event<<<static_cast<unsigned>((n_events + 255) / 256), 256, 0, stream>>>(d_off, n_events, d);
        cudaMemcpyAsync(&first, d, sizeof first, cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        cudaFree(d);
    }
    // the reference raises on empty sequences (yahmm.pyx:2836)
    set_error("event %lld has no observations (or is too long); every event needs at least one", first);
    return PP_ERR_ARG;
}

static int check_lane_model(pp_model* M) {
    if (!M->low.lane_ok) { set_error("model cannot run on the lane-per-event kernels: %s", M->low.lane_why.c_str()); return PP_ERR_UNSUPPORTED; }
    if (M->arena.size() > static_cast<size_t>(kArenaDoubles)) {
        set_error("model schedule needs %zu bytes of constant memory (limit %d)", M->arena.size() * 8, kArenaDoubles * 8);
        return PP_ERR_UNSUPPORTED;
    }
    return PP_OK;
}

// ---- event bucketing ------------------------------------------------------------------------
void build_schedule(const int64_t* offsets, int64_t e0, int64_t e1, Schedule* out) {
    const int64_t n = e1 - e0;
    int64_t maxlen = 0;
    for (int64_t e = e0; e < e1; ++e) maxlen = std::max(maxlen, offsets[e + 1] - offsets[e]);
    std::vector<int32_t> sorted(static_cast<size_t>(n));
    if (maxlen <= (1 << 22)) {                       // counting sort, longest first, stable
        std::vector<int64_t> count(static_cast<size_t>(maxlen) + 2, 0);
        for (int64_t e = e0; e < e1; ++e) ++count[static_cast<size_t>(maxlen - (offsets[e + 1] - offsets[e])) + 1];
        for (size_t i = 1; i < count.size(); ++i) count[i] += count[i - 1];
        for (int64_t e = e0; e < e1; ++e)
            sorted[static_cast<size_t>(count[static_cast<size_t>(maxlen - (offsets[e + 1] - offsets[e]))]++)] =
                static_cast<int32_t>(e);
    } else {
        for (int64_t i = 0; i < n; ++i) sorted[static_cast<size_t>(i)] = static_cast<int32_t>(e0 + i);
        std::stable_sort(sorted.begin(), sorted.end(), [&](int32_t a, int32_t b) {
            return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b];
        });
    }
    const int32_t ng = static_cast<int32_t>((n + kLanes - 1) / kLanes);
    out->n_groups = ng;
    out->order.assign(static_cast<size_t>(ng) * kLanes, -1);
    std::copy(sorted.begin(), sorted.end(), out->order.begin());
    out->group_len.resize(ng);
    out->group_row.resize(ng);
    int64_t rows = 0;
    for (int32_t g = 0; g < ng; ++g) {
        const int32_t first = sorted[static_cast<size_t>(g) * kLanes];
        const int32_t len = static_cast<int32_t>(offsets[first + 1] - offsets[first]);
        out->group_len[g] = len;
        out->group_row[g] = rows;
        rows += len + 1;
    }
    out->total_rows = rows;
}

int upload_model(pp_model* M) {
    cudaStream_t s = M->stream;
    const Lowered& L = M->low;
    for (int q = 0; q < 2; ++q) {
        PP_CUDA(upload(&M->d_edges_logp[q], L.edges_logp[q], s));
        PP_CUDA(upload(&M->d_edges_prob[q], L.edges_prob[q], s));
    }
    PP_CUDA(upload(&M->d_emis, L.emis, s));
    PP_CUDA(upload(&M->d_warp_run_ranges, L.warp_run_ranges, s));
    {   // host image of the constant arena (uploaded to c_arena right before a launch, see bind_arena)
        static_assert(sizeof(RunRec) % 8 == 0, "RunRec size");
        const size_t run_d = L.runs.size() * sizeof(RunRec) / 8, n_low = L.low_src.size();
        M->arena_emis = static_cast<int32_t>(run_d);                 // log(state.weight) per emitting state
        M->arena.assign(run_d + static_cast<size_t>(L.ne), 0.0);
        if (!L.runs.empty()) std::memcpy(M->arena.data(), L.runs.data(), L.runs.size() * sizeof(RunRec));
        for (int k = 0; k < L.ne; ++k) M->arena[run_d + k] = L.emis[k].logw;
        lane_arena_invalidate(M);                                    // force a re-upload
        // global copies the kernels stage into shared memory: weights in lowered-edge order and
        // 4-double emission records per kernel family
        std::vector<double> wl(n_low), wp(n_low), ev(5 * static_cast<size_t>(L.ne)), ef(4 * static_cast<size_t>(L.ne));
        for (size_t i = 0; i < n_low; ++i) { wl[i] = L.edges_logp[0][i].val; wp[i] = L.edges_prob[0][i].val; }
        for (int k = 0; k < L.ne; ++k) {
            const EmisRec& R = L.emis[k];
            double* a = &ev[5 * static_cast<size_t>(k)];
            double* b = &ef[4 * static_cast<size_t>(k)];
            a[0] = R.a0; a[1] = R.a1; a[2] = R.a2; a[3] = R.a3; a[4] = R.logw;
            b[0] = R.a0; b[1] = R.kind == EMIS_NORMAL ? R.k2 : R.a1; b[2] = R.c2; b[3] = R.logw * 1.4426950408889634074;
        }
        PP_CUDA(upload(&M->d_wtab_logp, wl, s));
        PP_CUDA(upload(&M->d_wtab_prob, wp, s));
        PP_CUDA(upload(&M->d_etab_vit, ev, s));
        PP_CUDA(upload(&M->d_etab_fwd, ef, s));
    }
    if (L.bw_ok) {                                                   // forward-backward lane kernel: forward runs | backward runs
        const size_t nf = L.runs.size() * sizeof(RunRec) / 8, nb = L.bw_runs.size() * sizeof(RunRec) / 8;
        M->arena_fb.assign(nf + nb, 0.0);
        if (nf) std::memcpy(M->arena_fb.data(), L.runs.data(), nf * 8);
        if (nb) std::memcpy(M->arena_fb.data() + nf, L.bw_runs.data(), nb * 8);
        PP_CUDA(upload(&M->d_bw_wtab, L.bw_wtab, s));
        PP_CUDA(upload(&M->d_acc_map, L.acc_map, s));
        PP_CUDA(upload(&M->d_acc_log2, L.acc_log2, s));
    } else {
        M->arena_fb.clear();
    }
    PP_CUDA(upload(&M->d_state_edge_begin, L.state_edge_begin, s));
    PP_CUDA(upload(&M->d_low_src, L.low_src, s));
    PP_CUDA(upload(&M->d_state_pword, L.state_pword, s));
    PP_CUDA(upload(&M->d_state_pshift, L.state_pshift, s));
    PP_CUDA(upload(&M->d_state_pmul, L.state_pmul, s));
    PP_CUDA(upload(&M->d_state_pmask, std::vector<int32_t>(static_cast<size_t>(M->m), (1 << L.ptr_bits) - 1), s));
    {   // traceback lookup of the run-based lane kernels: one record per state + the lowered in-edge sources
        std::vector<TraceRec> rec(static_cast<size_t>(M->m));
        for (int k = 0; k < M->m; ++k)
            rec[k] = TraceRec{static_cast<uint32_t>(L.state_edge_begin[k]), static_cast<uint32_t>(std::max(0, L.state_pword.empty() ? 0 : L.state_pword[k])),
                              static_cast<uint32_t>(L.state_pmul.empty() ? 1 : L.state_pmul[k]),
                              static_cast<uint32_t>(((L.state_pshift.empty() ? 0 : L.state_pshift[k]) << 4) | ((1 << L.ptr_bits) - 1 & 15))};
        std::vector<uint16_t> src(L.low_src.begin(), L.low_src.end());
        PP_CUDA(upload(&M->d_trace_rec[0], rec, s));
        PP_CUDA(upload(&M->d_trace_src[0], src, s));
        M->n_trace_src[0] = static_cast<int32_t>(src.size());
    }
    if (M->plan.ok) {
        const ProfilePlan& F = M->plan;
        PP_CUDA(upload(&M->d_prof_rec_v, F.rec_v, s));
        PP_CUDA(upload(&M->d_prof_rec_f, F.rec_f, s));
        PP_CUDA(upload(&M->d_prof_end_v, F.end_v, s));
        PP_CUDA(upload(&M->d_prof_end_f, F.end_f, s));
        PP_CUDA(upload(&M->d_prof_unit_off, F.unit_off, s));
        PP_CUDA(upload(&M->d_prof_unit_bytes, F.unit_bytes, s));
        PP_CUDA(upload(&M->d_prof_unit_shift, F.unit_shift, s));
        PP_CUDA(upload(&M->d_prof_unit_mask, F.unit_mask, s));
        PP_CUDA(upload(&M->d_prof_slot_begin, F.slot_begin, s));
        PP_CUDA(upload(&M->d_prof_slot_src, F.slot_src, s));
        std::vector<TraceRec> rec(static_cast<size_t>(M->m));
        for (int k = 0; k < M->m; ++k)
            rec[k] = TraceRec{static_cast<uint32_t>(F.slot_begin[k]), static_cast<uint32_t>(F.unit_off[k]), static_cast<uint32_t>(F.unit_bytes[k]),
                              static_cast<uint32_t>((F.unit_shift[k] << 4) | (F.unit_mask[k] & 15))};
        std::vector<uint16_t> src(F.slot_src.size());
        for (size_t i = 0; i < src.size(); ++i) src[i] = static_cast<uint16_t>(F.slot_src[i] < 0 ? 0 : F.slot_src[i]);
        PP_CUDA(upload(&M->d_trace_rec[1], rec, s));
        PP_CUDA(upload(&M->d_trace_src[1], src, s));
        M->n_trace_src[1] = static_cast<int32_t>(src.size());
    }
    if (M->plan_fb.ok && M->plan_fb.fb_ok) {
        const ProfilePlan& F = M->plan_fb;
        PP_CUDA(upload(&M->d_pfb_rec_f, F.rec_f, s));
        PP_CUDA(upload(&M->d_pfb_end_f, F.end_f, s));
        PP_CUDA(upload(&M->d_prof_rec_b, F.rec_b, s));
        PP_CUDA(upload(&M->d_prof_fb_map, F.fb_map, s));
        PP_CUDA(upload(&M->d_prof_fb_log2, F.fb_log2, s));
        PP_CUDA(upload(&M->d_prof_emit_state, F.fb_emit_state, s));
    }
    PP_CUDA(upload(&M->d_in_ptr, M->in_ptr, s));
    PP_CUDA(upload(&M->d_in_src, M->in_src, s));
    PP_CUDA(upload(&M->d_in_logp, M->in_logp, s));
    PP_CUDA(upload(&M->d_out_ptr, M->out_ptr, s));
    PP_CUDA(upload(&M->d_out_dst, M->out_dst, s));
    PP_CUDA(upload(&M->d_out_logp, M->out_logp, s));
    std::vector<double> prob(M->out_logp.size());
    for (size_t i = 0; i < prob.size(); ++i) prob[i] = std::exp(M->out_logp[i]);
    PP_CUDA(upload(&M->d_out_prob, prob, s));
    // alive[s]: an emission is still reachable from s without passing through the end state
    std::vector<int32_t> alive(M->m, 0);
    for (int k = 0; k < M->ne; ++k) alive[k] = 1;
    for (bool changed = true; changed;) {
        changed = false;
        for (int k = M->ne; k < M->m; ++k) {
            if (alive[k] || k == M->end) continue;
            for (int e = M->out_ptr[k]; e < M->out_ptr[k + 1]; ++e)
                if (M->out_dst[e] != M->end && alive[M->out_dst[e]]) { alive[k] = 1; changed = true; break; }
        }
    }
    PP_CUDA(upload(&M->d_alive, alive, s));
    int rc = upload_state_model(M);
    if (rc != PP_OK) return rc;
    PP_CUDA(cudaStreamSynchronize(s));
    return PP_OK;
}

static int copy_desc(pp_model* M, const pp_model_desc* d) {
    M->m = d->n_states; M->ne = d->silent_start; M->n_edges = d->n_edges;
    M->start = d->start_index; M->end = d->end_index; M->finite = d->finite;
    M->in_ptr.assign(d->in_ptr, d->in_ptr + d->n_states + 1);
    M->out_ptr.assign(d->out_ptr, d->out_ptr + d->n_states + 1);
    M->in_src.assign(d->in_src, d->in_src + d->n_edges);
    M->out_dst.assign(d->out_dst, d->out_dst + d->n_edges);
    M->in_logp.assign(d->in_logp, d->in_logp + d->n_edges);
    M->out_logp.assign(d->out_logp, d->out_logp + d->n_edges);
    M->emit_kind.assign(d->emit_kind, d->emit_kind + d->silent_start);
    M->emit_p0.assign(d->emit_p0, d->emit_p0 + d->silent_start);
    M->emit_p1.assign(d->emit_p1, d->emit_p1 + d->silent_start);
    M->emit_logw.assign(d->emit_logw, d->emit_logw + d->silent_start);
    M->has_table_states = false;
    for (int k = 0; k < d->silent_start; ++k) M->has_table_states |= d->emit_kind[k] == PP_EMIT_TABLE;
    return PP_OK;
}

LaneSched make_sched(const pp_model* M, bool logspace) {
    LaneSched S;
    const Lowered& L = M->low;
    for (int q = 0; q < 2; ++q) S.end_edges[q] = (logspace ? M->d_edges_logp[q] : M->d_edges_prob[q]).as<EdgeRec>();
    for (int i = 0; i < kMaxLaneWarps * 6; ++i) S.run_ranges[i] = i < static_cast<int>(L.warp_run_ranges.size()) ? L.warp_run_ranges[i] : 0;
    S.wtab = (logspace ? M->d_wtab_logp : M->d_wtab_prob).as<double>();
    S.etab = (logspace ? M->d_etab_vit : M->d_etab_fwd).as<double>();
    S.n_low = static_cast<int32_t>(L.low_src.size());
    S.etab_doubles = logspace ? 5 : 4;
    S.end_edge_begin = L.state_edge_begin[L.end];
    S.end_n_edges = L.state_n_edges[L.end];
    S.ptr_bits = L.ptr_bits;
    S.m = L.m; S.ne = L.ne; S.ns = L.ns; S.n_slots = L.n_slots;
    S.start_slot_off = static_cast<int32_t>((L.start - L.ne) * kSlotBytes);
    S.end_slot_off = static_cast<int32_t>((L.end - L.ne) * kSlotBytes);
    S.end_state = L.end; S.finite = L.finite; S.end_final_only = L.end_final_only ? 1 : 0;
    S.n_ptr_words = L.n_ptr_words;
    S.gcol = nullptr;
    S.gcol_stride = 0;
    return S;
}

// The register-resident kernels of pp_profile_kernels.cuh serve a model whose baked graph is a linear profile.
static bool use_profile(const pp_model* M) { return M->plan.ok && (M->kernel_path == 0 || M->kernel_path == 3); }
static size_t ptr_row_bytes(const pp_model* M) {
    return use_profile(M) ? static_cast<size_t>(M->plan.row_bytes) : static_cast<size_t>(M->low.n_ptr_words) * kLanes * sizeof(uint32_t);
}
ProfSched make_prof_sched(const pp_model* M, bool logspace) {
    const ProfilePlan& F = M->plan;
    ProfSched S;
    S.rec = (logspace ? M->d_prof_rec_v : M->d_prof_rec_f).as<double>();
    S.end_list = (logspace ? M->d_prof_end_v : M->d_prof_end_f).as<ProfEndRec>();
    S.end_n = static_cast<int32_t>(F.end_v.size());
    S.K = F.K; S.Kp = F.Kp; S.Kr = F.Kr; S.NW = F.NW;
    S.row_bytes = F.row_bytes;
    S.end_state = M->end;
    return S;
}

struct Timer {
    pp_model* M;
    int slot;
    Timer(pp_model* m, int s) : M(m), slot(s) { cudaEventRecord(M->ev[0], M->stream); }
    void stop() {
        cudaEventRecord(M->ev[1], M->stream);
        cudaEventSynchronize(M->ev[1]);
        float ms = 0;
        cudaEventElapsedTime(&ms, M->ev[0], M->ev[1]);
        M->prof[slot] += ms;
    }
};

struct ClampLen {
    __host__ __device__ int64_t operator()(const int64_t& x) const { return x > 0 ? x : 0; }
};

// Validates the CSR-style event description and returns host offsets.
static int host_offsets(pp_model* M, const int64_t* offsets, int64_t n_events, int32_t mem,
                        std::vector<int64_t>* store, const int64_t** h_off, const int64_t** d_off) {
    if (!offsets || n_events < 0) { set_error("null offsets / negative n_events"); return PP_ERR_ARG; }
    if (n_events > 0x7fffffff - 64) { set_error("more than 2^31 events in one call"); return PP_ERR_ARG; }
    if (mem == PP_MEM_DEVICE) {
        store->resize(static_cast<size_t>(n_events) + 1);
        PP_CUDA(cudaMemcpyAsync(store->data(), offsets, store->size() * sizeof(int64_t), cudaMemcpyDeviceToHost, M->stream));
        PP_CUDA(cudaStreamSynchronize(M->stream));
        *h_off = store->data();
        *d_off = offsets;
    } else {
        *h_off = offsets;
        PP_CUDA(M->d_offsets.reserve((static_cast<size_t>(n_events) + 1) * sizeof(int64_t)));
        PP_CUDA(cudaMemcpyAsync(M->d_offsets.p, offsets, (static_cast<size_t>(n_events) + 1) * sizeof(int64_t),
                                cudaMemcpyHostToDevice, M->stream));
        *d_off = M->d_offsets.as<int64_t>();
    }
    const int64_t* h = *h_off;
    for (int64_t e = 0; e < n_events; ++e) {
        const int64_t len = h[e + 1] - h[e];
        if (len < 1) {                                   // the reference raises on empty sequences (yahmm.pyx:2836)
            set_error("event %lld has %lld observations; every event needs at least one", (long long)e, (long long)len);
            return PP_ERR_ARG;
        }
        if (len > 0x7ffffff0) { set_error("event %lld too long", (long long)e); return PP_ERR_ARG; }
    }
    return PP_OK;
}

// Column in shared memory, or -- for a model whose column does not fit next to the tables -- in a per-CTA global
// scratch (k_*_lane<..., CG = true>): same kernels, same results.
static size_t lane_smem_col(const pp_model* M) {
    return lane_smem_bytes2(M->low.n_slots, static_cast<int>(M->low.low_src.size()), M->low.ne, kLaneWarps);
}
static size_t lane_smem_nocol(const pp_model* M) {
    return lane_smem_bytes2(0, static_cast<int>(M->low.low_src.size()), M->low.ne, kLaneWarps);
}
static bool lane_col_global(const pp_model* M) {
    return M->low.lane_ok && M->low.ptr_bits == 4 && lane_smem_col(M) > M->smem_optin && lane_smem_nocol(M) <= M->smem_optin &&
           M->arena.size() <= static_cast<size_t>(kArenaDoubles);
}
constexpr int kGcolGroups = 1024;         // groups (CTAs) per launch of the global-column kernels: bounds the scratch

template <typename OutT, typename Launch>
static int launch_gcol(pp_model* M, const LaneSched& S0, const LaneBatchAny& A, const OutT& O, int n_groups, Launch launch,
                       const char* what) {
    LaneSched S = S0;
    const size_t col_bytes = static_cast<size_t>(M->low.n_slots) * kSlotBytes;
    const int per = std::min(n_groups, kGcolGroups);
    PP_CUDA(M->d_gcol.reserve(static_cast<size_t>(per) * col_bytes));
    S.gcol = M->d_gcol.as<unsigned char>();
    S.gcol_stride = col_bytes;
    const size_t smem = lane_smem_nocol(M);
    for (int g0 = 0; g0 < n_groups; g0 += per) {
        const int ng = std::min(per, n_groups - g0);
        const LaneBatchAny Ag{A.obs, A.obs_f64, A.offsets, A.order + static_cast<size_t>(g0) * kLanes, A.group_len + g0, A.group_row + g0};
        const int e = launch(S, Ag, O, ng, smem, M->stream, M->arena.data(), M->arena.size(), M);
        if (e != 0) { set_error("%s kernel launch failed: %s", what, cudaGetErrorString(static_cast<cudaError_t>(e))); return PP_ERR_CUDA; }
        ++M->launches;
    }
    return PP_OK;
}

static int launch_viterbi(pp_model* M, const LaneBatchAny& A, const VitOut& O, int n_groups) {
    int rc = check_lane_model(M);
    if (rc != PP_OK) return rc;
    if (use_profile(M)) {
        const int e = profile_viterbi_launch(make_prof_sched(M, true), A, O, n_groups, M->stream);
        if (e != 0) { set_error("profile Viterbi kernel launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return PP_ERR_CUDA; }
        ++M->launches;
        return PP_OK;
    }
    if (lane_col_global(M)) return launch_gcol(M, make_sched(M, true), A, O, n_groups, lane_viterbi4g_launch, "Viterbi (global column)");
    const size_t smem = lane_smem_col(M);
    // every item of a lane-lowered model has <= 8 in-edges (kMaxRunEdges), so its ordinals always take 4 bits
    if (M->low.ptr_bits != 4) { set_error("internal: %d-bit traceback ordinals have no kernel", M->low.ptr_bits); return PP_ERR_UNSUPPORTED; }
    const int e = lane_viterbi4_launch(make_sched(M, true), A, O, n_groups, smem, M->stream, M->arena.data(), M->arena.size(), M);
    if (e != 0) { set_error("Viterbi kernel launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return PP_ERR_CUDA; }
    ++M->launches;
    return PP_OK;
}

static int launch_forward(pp_model* M, const LaneBatchAny& A, const FwdOut& O, int n_groups) {
    int rc = check_lane_model(M);
    if (rc != PP_OK) return rc;
    if (use_profile(M)) {
        const int e = profile_forward_launch(make_prof_sched(M, false), A, O, n_groups, M->stream);
        if (e != 0) { set_error("profile forward kernel launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return PP_ERR_CUDA; }
        ++M->launches;
        return PP_OK;
    }
    if (lane_col_global(M)) return launch_gcol(M, make_sched(M, false), A, O, n_groups, lane_forwardg_launch, "forward (global column)");
    const size_t smem = lane_smem_col(M);
    const int e = lane_forward_launch(make_sched(M, false), A, O, n_groups, smem, M->stream, M->arena.data(), M->arena.size(), M);
    if (e != 0) { set_error("forward kernel launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return PP_ERR_CUDA; }
    ++M->launches;
    return PP_OK;
}

// The lane-per-event kernels need the float64 column of 32 events (+ weights, emission records) in shared
// memory, every in-degree <= 8 and the run records in the constant arena; other models take the
// state-parallel kernels.
static bool lane_kernels_fit(const pp_model* M) {
    const size_t smem = lane_smem_col(M);
    if (lane_col_global(M)) return true;
    return M->low.lane_ok && smem <= M->smem_optin && M->arena.size() <= static_cast<size_t>(kArenaDoubles);
}

// Splits [0, n_events) into chunks of whole events whose pointer table / staged observations fit.
static std::vector<int64_t> make_chunks(const int64_t* h_off, int64_t n_events, size_t row_bytes, size_t ws_limit,
                                        size_t obs_bytes, size_t obs_limit) {
    std::vector<int64_t> cuts{0};
    int64_t e0 = 0;
    while (e0 < n_events) {
        int64_t e = e0;
        size_t rows = 0, maxlen = 0, ob = 0;
        while (e < n_events) {
            const size_t len = static_cast<size_t>(h_off[e + 1] - h_off[e]);
            const size_t nrows = rows + len + 1, nmax = std::max(maxlen, len);
            // sorted groups of 32: sum_g 32*(Lmax_g+1) <= sum_e (n_e+1) + 32*(Lmax+1)
            const size_t need = row_bytes ? ((nrows + kLanes * (nmax + 1)) / kLanes + 1) * row_bytes : 0;
            const size_t nob = ob + len * obs_bytes;
            if (e > e0 && (need > ws_limit || nob > obs_limit)) break;
            rows = nrows; maxlen = nmax; ob = nob;
            ++e;
        }
        cuts.push_back(e);
        e0 = e;
    }
    return cuts;
}

}  // namespace pp

using namespace pp;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int pp_abi_version(void) { return 1; }

const char* pp_last_error(void) { return g_err.c_str(); }

int pp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pp_model_create(const pp_model_desc* desc, int device, pp_model** out) {
    if (!desc || !out) { set_error("null argument"); return PP_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: protpore_b200 has no CPU fallback");
        return PP_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return PP_ERR_ARG; }
    pp_model* M = new (std::nothrow) pp_model();
    if (!M) { set_error("out of host memory"); return PP_ERR_NOMEM; }
    M->device = device;
    std::string err = lower_model(*desc, M->warps, &M->low);
    if (!err.empty()) { set_error("cannot lower model: %s", err.c_str()); delete M; return PP_ERR_UNSUPPORTED; }
    copy_desc(M, desc);
    build_profile_plan(*desc, M->low, kProfPositionsPerWarp, kProfMaxPositionWarps, &M->plan);
    build_profile_plan(*desc, M->low, kPfbPositionsPerWarp, kPfbMaxPositionWarps, &M->plan_fb);
    if (const char* env = std::getenv("PP_KERNEL_PATH")) M->kernel_path = std::atoi(env);
    int rc = PP_OK;
    do {
        if (cudaSetDevice(device) != cudaSuccess) { rc = PP_ERR_CUDA; break; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rc = PP_ERR_CUDA; break; }
        M->sm_count = prop.multiProcessorCount;
        M->smem_optin = prop.sharedMemPerBlockOptin;
        if (cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = PP_ERR_CUDA; break; }
        if (cudaEventCreate(&M->ev[0]) != cudaSuccess || cudaEventCreate(&M->ev[1]) != cudaSuccess) { rc = PP_ERR_CUDA; break; }
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { rc = PP_ERR_CUDA; break; }
        M->ws_limit = static_cast<size_t>(fr * 0.70);
        rc = upload_model(M);
    } while (0);
    if (rc != PP_OK) {
        if (g_err.empty() || rc == PP_ERR_CUDA) set_error("CUDA initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        pp_model_destroy(M);
        return rc;
    }
    *out = M;
    return PP_OK;
}

void pp_model_destroy(pp_model* M) {
    if (!M) return;
    cudaSetDevice(M->device);
    if (M->stream) cudaStreamSynchronize(M->stream);
    lane_arena_invalidate(M);
    DevBuf* bufs[] = {&M->d_edges_logp[0], &M->d_edges_logp[1], &M->d_edges_prob[0],
                      &M->d_edges_prob[1], &M->d_emis, &M->d_warp_run_ranges, &M->d_wtab_logp, &M->d_wtab_prob, &M->d_etab_vit, &M->d_etab_fwd, &M->d_state_edge_begin, &M->d_low_src, &M->d_state_pword, &M->d_state_pshift, &M->d_state_pmul,
                      &M->d_in_ptr, &M->d_in_src, &M->d_in_logp, &M->d_out_ptr, &M->d_out_dst, &M->d_out_logp,
                      &M->d_out_prob, &M->d_alive, &M->d_ws, &M->d_obs, &M->d_offsets, &M->d_order,
                      &M->d_group_len, &M->d_group_row, &M->d_score, &M->d_end_state, &M->d_end_ord, &M->d_path_len,
                      &M->d_path_off, &M->d_path, &M->d_scan_tmp, &M->d_logp, &M->d_bw_wtab, &M->d_acc_map,
                      &M->d_acc_log2, &M->d_fb_acc, &M->d_fb_scratch, &M->d_fb_ctl, &M->d_fb_minmax, &M->d_etab_own, &M->d_gcol, &M->d_prof_rec_v, &M->d_prof_rec_f, &M->d_prof_end_v, &M->d_prof_end_f,
                      &M->d_prof_unit_off, &M->d_prof_unit_bytes, &M->d_prof_unit_shift, &M->d_prof_unit_mask, &M->d_prof_slot_begin,
                      &M->d_prof_slot_src, &M->d_pfb_rec_f, &M->d_pfb_end_f, &M->d_prof_rec_b, &M->d_prof_fb_map, &M->d_prof_fb_log2, &M->d_prof_emit_state, &M->d_state_pmask, &M->d_trace_rec[0], &M->d_trace_rec[1], &M->d_trace_src[0], &M->d_trace_src[1]};
    for (DevBuf* b : bufs) b->release();
    release_state_model(M);
    for (int b = 0; b < kPipeSlots; ++b) {
        pp_model::PipeSlot& P = M->ps[b];
        DevBuf* pb[] = {&P.obs, &P.ws, &P.path, &P.order, &P.group_len, &P.group_row, &P.list, &P.counter, &P.sched, &P.path_tmp};
        for (DevBuf* x : pb) x->release();
        cudaEvent_t evs[] = {P.in_done, P.a_done, P.tb_done, P.out_done, P.sch_done};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : M->evpool) cudaEventDestroy(e);
    if (M->h_pin) cudaFreeHost(M->h_pin);
    if (M->s_in) cudaStreamDestroy(M->s_in);
    if (M->s_out) cudaStreamDestroy(M->s_out);
    if (M->ev[0]) cudaEventDestroy(M->ev[0]);
    if (M->ev[1]) cudaEventDestroy(M->ev[1]);
    if (M->stream) cudaStreamDestroy(M->stream);
    delete M;
}

int pp_model_update_params(pp_model* M, const pp_model_desc* desc) {
    if (!M || !desc) { set_error("null argument"); return PP_ERR_ARG; }
    std::string err = refresh_params(*desc, &M->low);
    if (!err.empty()) { set_error("cannot refresh parameters: %s", err.c_str()); return PP_ERR_ARG; }
    copy_desc(M, desc);
    build_profile_plan(*desc, M->low, kProfPositionsPerWarp, kProfMaxPositionWarps, &M->plan);
    build_profile_plan(*desc, M->low, kPfbPositionsPerWarp, kPfbMaxPositionWarps, &M->plan_fb);
    PP_CUDA(cudaSetDevice(M->device));
    PP_CUDA(cudaStreamSynchronize(M->stream));
    return upload_model(M);
}

int pp_model_set_emission_table(pp_model* M, const double* table, int64_t n_rows, int32_t mem) {
    if (!M || n_rows < 0) { set_error("null or negative argument"); return PP_ERR_ARG; }
    if (mem != PP_MEM_HOST && mem != PP_MEM_DEVICE) { set_error("bad mem"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(M->device));
    if (!table) { M->etab = nullptr; M->etab_rows = 0; return PP_OK; }
    if (mem == PP_MEM_DEVICE) {
        M->etab = table;
    } else {
        const size_t bytes = static_cast<size_t>(n_rows) * M->ne * sizeof(double);
        PP_CUDA(cudaStreamSynchronize(M->stream));            // an earlier sweep may still read the old copy
        PP_CUDA(M->d_etab_own.reserve(bytes + 16));
        PP_CUDA(cudaMemcpyAsync(M->d_etab_own.p, table, bytes, cudaMemcpyHostToDevice, M->stream));
        PP_CUDA(cudaStreamSynchronize(M->stream));
        M->etab = M->d_etab_own.as<double>();
    }
    M->etab_rows = n_rows;
    return PP_OK;
}

int64_t pp_schedule_describe(const pp_model_desc* desc, char* buf, int64_t cap) {
    if (!desc) { set_error("null argument"); return PP_ERR_ARG; }
    Lowered low;
    std::string err = lower_model(*desc, kLaneWarps, &low);
    if (!err.empty()) { set_error("cannot lower model: %s", err.c_str()); return PP_ERR_UNSUPPORTED; }
    std::string text = describe_schedule(low);
    {   // which kernel family takes forward / Viterbi sweeps of this model (pp_profile.h)
        ProfilePlan plan;
        build_profile_plan(*desc, low, kProfPositionsPerWarp, kProfMaxPositionWarps, &plan);
        char line[512];
        if (plan.ok)
            std::snprintf(line, sizeof line, "profile ok positions %d position_warps %d positions_per_warp %d ptr_row_bytes %d aliases %d end_edges %zu\n",
                          plan.K, plan.NW, plan.P, plan.row_bytes, plan.n_alias, plan.end_v.size());
        else
            std::snprintf(line, sizeof line, "profile no: %s\n", plan.why.c_str());
        text += line;
        if (plan.ok) {
            ProfilePlan pfb;
            build_profile_plan(*desc, low, kPfbPositionsPerWarp, kPfbMaxPositionWarps, &pfb);
            if (pfb.ok && pfb.fb_ok) std::snprintf(line, sizeof line, "profile_fb ok position_warps %d positions_per_warp %d slots %d\n", pfb.NW, pfb.P, pfb.fb_slots);
            else std::snprintf(line, sizeof line, "profile_fb no: %s\n", (pfb.ok ? pfb.fb_why : pfb.why).c_str());
            text += line;
        }
    }
    if (buf && cap > 0) std::snprintf(buf, static_cast<size_t>(cap), "%s", text.c_str());
    return static_cast<int64_t>(text.size());
}


int pp_model_set_kernel_path(pp_model* M, int32_t path) {
    if (!M || path < 0 || path > 3) { set_error("bad kernel path"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(M->device));
    PP_CUDA(cudaStreamSynchronize(M->stream));
    M->kernel_path = path;
    return PP_OK;
}

int pp_model_kernel_info(const pp_model* M, int32_t* out4) {
    if (!M || !out4) { set_error("null argument"); return PP_ERR_ARG; }
    out4[0] = use_profile(M) ? 2 : (lane_kernels_fit(M) ? 1 : 0);
    out4[1] = M->plan.ok ? M->plan.K : 0;
    out4[2] = use_profile(M) ? M->plan.NW + 1 : kLaneWarps;
    out4[3] = static_cast<int32_t>(ptr_row_bytes(M));
    return PP_OK;
}

int pp_model_set_workspace_limit(pp_model* M, int64_t bytes) {
    if (!M || bytes <= 0) { set_error("bad workspace limit"); return PP_ERR_ARG; }
    M->ws_limit = static_cast<size_t>(bytes);
    return PP_OK;
}

int64_t pp_viterbi_workspace_bytes(const pp_model* M, const int64_t* offsets_host, int64_t n_events) {
    if (!M || !offsets_host || n_events < 0) return -1;
    Schedule sc;
    build_schedule(offsets_host, 0, n_events, &sc);
    return sc.total_rows * static_cast<int64_t>(ptr_row_bytes(M));
}

#ifdef PP_PHASE_TIMING
// debug build only (tools/phase_timing.py): per-warp phase clocks of the Viterbi lane kernel
int pp_debug_phase_cycles(unsigned long long* out, int reset) { return lane_debug_phase_cycles(out, reset); }
#endif

int64_t pp_launch_count(const pp_model* M) { return M ? M->launches : 0; }

int pp_debug_check_guards(void) {
    if (!GuardRegistry::enabled()) { set_error("guard mode is off (set PP_GUARD=1 before the library allocates)"); return -1; }
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("device error before the guard check: %s", cudaGetErrorString(cudaGetLastError())); return -2; }
    GuardRegistry& G = GuardRegistry::get();
    std::lock_guard<std::mutex> lk(G.mu);
    std::vector<unsigned char> h(2 * kGuardBytes + 256);
    int bad = 0;
    for (const auto& kv : G.live) {
        const char* base = static_cast<const char*>(kv.first);
        const size_t bytes = kv.second, payload = (bytes + 255) / 256 * 256, back = payload - bytes + kGuardBytes;
        if (cudaMemcpy(h.data(), base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h.data() + kGuardBytes, base + kGuardBytes + bytes, back, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("cannot read the guards of a %zu-byte buffer", bytes);
            return -2;
        }
        for (size_t i = 0; i < kGuardBytes + back; ++i)
            if (h[i] != kGuardPattern) {
                if (!bad) set_error("guard of a %zu-byte device buffer overwritten: %s the buffer, byte %zu", bytes,
                                    i < kGuardBytes ? "in front of" : "behind", i < kGuardBytes ? kGuardBytes - i : i - kGuardBytes);
                ++bad;
                break;
            }
    }
    return bad;
}

int pp_debug_guard_selftest(void) {
    if (!GuardRegistry::enabled()) { set_error("guard mode is off"); return -1; }
    const int before = pp_debug_check_guards();
    if (before != 0) return -3;
    DevBuf b;
    PP_CUDA(b.reserve(1000));
    PP_CUDA(cudaMemset(b.p, 0, 1001));                     // one byte too many
    const int behind = pp_debug_check_guards();
    b.release();
    PP_CUDA(b.reserve(1000));
    PP_CUDA(cudaMemset(static_cast<char*>(b.p) - 8, 0, 8));  // eight bytes in front
    const int front = pp_debug_check_guards();
    b.release();
    const int after = pp_debug_check_guards();
    if (behind == 1 && front == 1 && after == 0) return PP_OK;
    set_error("guard self-test: behind %d, in front %d, after %d (expected 1, 1, 0)", behind, front, after);
    return -3;
}

int pp_debug_guard_count(void) {
    GuardRegistry& G = GuardRegistry::get();
    std::lock_guard<std::mutex> lk(G.mu);
    return static_cast<int>(G.live.size());
}
uint64_t pp_model_stream(const pp_model* M) { return M ? reinterpret_cast<uint64_t>(M->stream) : 0; }

int pp_profile_get(const pp_model* M, double* out, int32_t n) {
    if (!M || !out) { set_error("null argument"); return PP_ERR_ARG; }
    for (int i = 0; i < n && i < PP_PROF_SLOTS; ++i) out[i] = M->prof[i];
    return PP_OK;
}

void* pp_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, static_cast<size_t>(bytes), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pp_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------------
// Chunk pipeline shared by pp_viterbi_batch and pp_forward_batch.  A batch is cut into chunks of whole events; chunk c
// uses buffer slot c % kPipeSlots.  Per chunk, in enqueue order:
//   s_in    : upload of the chunk's bucketing arrays and (host buffers) its observations            -> in_done
//   stream  : phase A = sweep kernel(s) + whatever produces the scalar the host needs (path total,
//             number of events for the exact re-run), copied to pinned memory                        -> a_done
//   host    : waits for a_done of chunk c - 1 only AFTER phase A of chunk c is enqueued, so the GPU never idles
//   stream  : phase B of chunk c - 1 (traceback write pass / exact re-run)                           -> tb_done
//   s_out   : device -> host copy of the chunk's paths                                              -> out_done
// so copies in, sweeps and copies out of neighbouring chunks overlap.
static int pipe_init(pp_model* M) {
    if (M->s_in) return PP_OK;
    PP_CUDA(cudaStreamCreateWithFlags(&M->s_in, cudaStreamNonBlocking));
    PP_CUDA(cudaStreamCreateWithFlags(&M->s_out, cudaStreamNonBlocking));
    for (int b = 0; b < kPipeSlots; ++b) {
        PP_CUDA(cudaEventCreateWithFlags(&M->ps[b].in_done, cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&M->ps[b].a_done, cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&M->ps[b].tb_done, cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&M->ps[b].out_done, cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&M->ps[b].sch_done, cudaEventDisableTiming));
    }
    PP_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&M->h_pin), 64 * sizeof(int64_t), cudaHostAllocDefault));
    return PP_OK;
}

static cudaEvent_t pool_event(pp_model* M) {
    if (M->evnext == M->evpool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        M->evpool.push_back(e);
    }
    return M->evpool[M->evnext++];
}
// timed span on a stream, resolved by spans_collect after the call's final synchronisation
struct SpanGuard {
    pp_model* M;
    cudaStream_t s;
    cudaEvent_t b;
    SpanGuard(pp_model* m, int slot, cudaStream_t st) : M(m), s(st) {
        cudaEvent_t a = pool_event(M);
        b = pool_event(M);
        cudaEventRecord(a, s);
        M->spans.push_back({a, b, slot});
    }
    ~SpanGuard() { cudaEventRecord(b, s); }
};
static void spans_collect(pp_model* M) {
    if (!M->spans.empty() && std::getenv("PP_TRACE_SPANS")) {       // debugging aid: the call's timeline, ms from its first span
        static const char* names[] = {"viterbi", "traceback", "forward", "backward", "sample", "h2d", "d2h", "host"};
        for (const pp_model::Span& sp : M->spans) {
            float a = 0, b = 0;
            if (cudaEventElapsedTime(&a, M->spans.front().a, sp.a) == cudaSuccess && cudaEventElapsedTime(&b, M->spans.front().a, sp.b) == cudaSuccess)
                std::fprintf(stderr, "[pp span] %-9s %9.3f .. %9.3f ms\n", sp.slot >= 0 && sp.slot < 8 ? names[sp.slot] : "?", a, b);
        }
        cudaGetLastError();
    }
    for (const pp_model::Span& sp : M->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) M->prof[sp.slot] += ms;
    }
    cudaGetLastError();
    M->spans.clear();
    M->evnext = 0;
}

// After a failed batched call nothing may be left in flight on the handle's streams (async copies into the caller's
// buffers, reads of the caller's observations) and the timing spans of the aborted call must not leak into the next one.
static int settle(pp_model* M, int rc) {
    if (rc != PP_OK && rc != PP_ERR_CAPACITY) {
        cudaDeviceSynchronize();
        cudaGetLastError();
        M->spans.clear();
        M->evnext = 0;
    }
    return rc;
}

// One pipeline for the batched Viterbi decode (logp_out / path_* set), the batched forward score (fwd_out set) or both
// (pp_score_batch): one upload, one bucketing.
static int score_pipeline(pp_model* M, const void* obs, int32_t obs_dtype, const int64_t* offsets, int64_t n_events,
                          int32_t mem, double* logp_out, int64_t* path_len_out, uint16_t* path_out, int64_t path_capacity,
                          int64_t* path_total, double* fwd_out) {
    const bool want_vit = logp_out != nullptr, want_fwd = fwd_out != nullptr;
    const size_t ob = obs_dtype == PP_F32 ? 4 : 8;
    int rc = pipe_init(M);
    if (rc != PP_OK) return rc;
    const bool host = mem == PP_MEM_HOST;

    // ---- offsets: on the device for the kernels; on the host only where the host has to cut the batch into chunks
    std::vector<int64_t> store;
    const int64_t* h_off = nullptr;
    const int64_t* d_off = nullptr;
    if (host) {
        if (!offsets || n_events < 0) { set_error("null offsets / negative n_events"); return PP_ERR_ARG; }
        std::vector<int64_t> dummy;
        rc = host_offsets(M, offsets, n_events, mem, &dummy, &h_off, &d_off);
        if (rc != PP_OK) return rc;
    } else {
        if (!offsets || n_events < 0) { set_error("null offsets / negative n_events"); return PP_ERR_ARG; }
        if (n_events > 0x7fffffff - 64) { set_error("more than 2^31 events in one call"); return PP_ERR_ARG; }
        d_off = offsets;
    }

    double* d_score = nullptr;
    int64_t* d_plen = nullptr;
    double* d_fwd = fwd_out;
    if (want_vit) {
        if (!host) {
            d_score = logp_out;
            d_plen = path_len_out;
        } else {
            PP_CUDA(M->d_score.reserve(n_events * sizeof(double)));
            PP_CUDA(M->d_path_len.reserve(n_events * sizeof(int64_t)));
            d_score = M->d_score.as<double>();
            d_plen = M->d_path_len.as<int64_t>();
        }
        PP_CUDA(M->d_end_state.reserve(n_events * sizeof(int32_t)));
        PP_CUDA(M->d_end_ord.reserve(n_events * sizeof(int32_t)));
        PP_CUDA(M->d_path_off.reserve((n_events + 1) * sizeof(int64_t)));
    }
    if (want_fwd && host) {
        PP_CUDA(M->d_logp.reserve(n_events * sizeof(double)));
        d_fwd = M->d_logp.as<double>();
    }
    PP_CUDA(cudaStreamSynchronize(M->stream));            // offsets are on the device before s_in / kernels use them

    const size_t row_bytes = want_vit ? ptr_row_bytes(M) : 0;
    // host wall time this call spends on bucketing: enqueueing it and waiting for its three scalars
    auto clock_now = [] { return std::chrono::steady_clock::now(); };
    auto add_host_ms = [&](std::chrono::steady_clock::time_point t0) {
        M->prof[PP_PROF_SCHEDULE_HOST_MS] += std::chrono::duration<double, std::milli>(clock_now() - t0).count();
    };
    std::vector<int64_t> cuts;
    DevSchedule pre;                                       // device-resident batch: the whole batch is bucketed first
    bool pre_scheduled = false;
    int64_t total_obs_dev = 0;
    if (host) {
        // kPipeSlots chunks in flight share the workspace; host batches are cut fine enough for the copies to overlap
        // ... but a chunk cannot finish before its longest group of 32 events has been swept row by row: keep the
        // chunk's share of work per resident CTA (2 per SM) at >= 4x that tail, or mixed-length batches (config 5:
        // 20 .. 20,000 segments) pay the tail once per chunk
        int64_t longest = 1;
        for (int64_t e = 0; e < n_events; ++e) longest = std::max(longest, h_off[e + 1] - h_off[e]);
        const size_t tail_rule = static_cast<size_t>(longest) * kLanes * 2 * M->sm_count * 4 * ob;
        static const size_t chunk_bytes = [] { const char* e = std::getenv("PP_CHUNK_MB"); return (e && std::atoi(e) > 0 ? static_cast<size_t>(std::atoi(e)) : size_t(96)) << 20; }();
        cuts = make_chunks(h_off, n_events, row_bytes, M->ws_limit / kPipeSlots, ob, std::max(chunk_bytes, tail_rule));
    } else {
        const auto t0 = clock_now();
        rc = device_schedule_enqueue(M, 0, d_off, 0, n_events, M->stream, &pre);
        if (rc != PP_OK) return rc;
        PP_CUDA(cudaStreamSynchronize(M->stream));
        add_host_ms(t0);
        const int64_t minlen = M->h_pin[16], maxlen = M->h_pin[17], rows = M->h_pin[18];
        if (minlen < 1 || maxlen > 0x7ffffff0LL) return bad_event_error(M, d_off, n_events, M->stream);
        int64_t ends[2] = {0, 0};
        PP_CUDA(cudaMemcpy(&ends[0], d_off, sizeof(int64_t), cudaMemcpyDeviceToHost));
        PP_CUDA(cudaMemcpy(&ends[1], d_off + n_events, sizeof(int64_t), cudaMemcpyDeviceToHost));
        total_obs_dev = ends[1] - ends[0];
        if (static_cast<size_t>(rows) * row_bytes <= M->ws_limit) {
            cuts = {0, n_events};
            pre_scheduled = true;
        } else {                                           // too large for one pointer table: cut on the host
            store.resize(static_cast<size_t>(n_events) + 1);
            PP_CUDA(cudaMemcpy(store.data(), d_off, store.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
            h_off = store.data();
            cuts = make_chunks(h_off, n_events, row_bytes, M->ws_limit / kPipeSlots, 0, size_t(96) << 20);
        }
    }
    const int n_chunks = static_cast<int>(cuts.size()) - 1;
    int64_t base = 0;
    bool overflow = false;
    struct ChunkInfo { int64_t e0, e1; int n_groups; int64_t maxlen; const void* obs_base; };
    std::vector<ChunkInfo> info(n_chunks);

    auto trace_args = [&](int c) {
        pp_model::PipeSlot& P = M->ps[c % kPipeSlots];
        const int f = use_profile(M) ? 1 : 0;
        return TraceArgs{P.ws.as<uint32_t>(), P.order.as<int32_t>(), P.group_row.as<int64_t>(), d_off, d_score,
                         M->d_end_state.as<int32_t>(), M->d_end_ord.as<int32_t>(), M->d_trace_rec[f].as<TraceRec>(),
                         M->d_trace_src[f].as<uint16_t>(), M->m, M->n_trace_src[f], M->ne, M->start,
                         (f || M->low.end_final_only) ? 1 : 0, static_cast<int64_t>(row_bytes), cuts[c]};
    };
    // phase B of chunk c: the host learns the chunk's path total (and how many events the scaled forward sweep could
    // not resolve), then the traceback write pass, the exact forward re-run and the copy out
    auto finish = [&](int c) -> int {
        pp_model::PipeSlot& P = M->ps[c % kPipeSlots];
        PP_CUDA(cudaEventSynchronize(P.a_done));
        if (want_fwd) {
            int r = exact_forward_run(M, c % kPipeSlots, info[c].obs_base, obs_dtype, d_off + info[c].e0, info[c].maxlen, d_fwd + info[c].e0);
            if (r != PP_OK) return r;
        }
        if (!want_vit) {
            PP_CUDA(cudaEventRecord(P.tb_done, M->stream));
            return PP_OK;
        }
        const int64_t last_off = M->h_pin[(c % kPipeSlots) * 2], last_len = M->h_pin[(c % kPipeSlots) * 2 + 1];
        const int64_t chunk_total = last_off + (last_len > 0 ? last_len : 0);
        const int nslots = info[c].n_groups * kLanes;
        if (path_out && !overflow && base + chunk_total <= path_capacity) {
            uint16_t* dst;
            if (!host) dst = path_out + base;
            else {
                PP_CUDA(cudaStreamWaitEvent(M->stream, P.out_done, 0));      // the slot's previous copy out has left the buffer
                PP_CUDA(P.path.reserve(static_cast<size_t>(chunk_total) * sizeof(uint16_t) + 16));
                dst = P.path.as<uint16_t>();
            }
            {
                SpanGuard t(M, PP_PROF_TRACEBACK, M->stream);
                if (M->h_pin[32 + (c % kPipeSlots)] == 0)      // every path fitted its scratch region: pack them
                    k_traceback_compact<<<M->sm_count * 16, 256, 0, M->stream>>>(trace_args(c), d_plen, M->d_path_off.as<int64_t>(),
                                                                                  P.path_tmp.as<uint16_t>(), dst, info[c].e1);
                else                                  // an unusually long path: walk again, writing in place
                    k_traceback<<<(nslots + 127) / 128, 128, trace_smem_bytes(M->m, M->n_trace_src[use_profile(M) ? 1 : 0]), M->stream>>>(trace_args(c), d_plen, M->d_path_off.as<int64_t>(), dst, nslots);
            }
            PP_CUDA(cudaGetLastError());
            ++M->launches;
            PP_CUDA(cudaEventRecord(P.tb_done, M->stream));
            if (host && chunk_total > 0) {
                PP_CUDA(cudaStreamWaitEvent(M->s_out, P.tb_done, 0));
                SpanGuard t(M, PP_PROF_D2H, M->s_out);
                PP_CUDA(cudaMemcpyAsync(path_out + base, dst, static_cast<size_t>(chunk_total) * sizeof(uint16_t),
                                        cudaMemcpyDeviceToHost, M->s_out));
            }
            if (host) PP_CUDA(cudaEventRecord(P.out_done, M->s_out));
        } else {
            PP_CUDA(cudaEventRecord(P.tb_done, M->stream));
            if (path_out) overflow = true;
        }
        base += chunk_total;
        return PP_OK;
    };

    for (int c = 0; c < n_chunks; ++c) {
        pp_model::PipeSlot& P = M->ps[c % kPipeSlots];
        const int64_t e0 = cuts[c], e1 = cuts[c + 1];
        // ---- copy in and bucketing (the slot's buffers were last used by chunk c - 2: its traceback must be through)
        DevSchedule sc = pre;
        if (!pre_scheduled) {
            const auto t0 = clock_now();
            PP_CUDA(cudaStreamWaitEvent(M->s_in, P.tb_done, 0));
            rc = device_schedule_enqueue(M, c % kPipeSlots, d_off, e0, e1, M->s_in, &sc);
            if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
            PP_CUDA(cudaEventRecord(P.sch_done, M->s_in));
            add_host_ms(t0);
        }
        const void* d_obs_base = obs;                        // indexable with the GLOBAL offsets
        if (host) {
            const size_t nob = static_cast<size_t>(h_off[e1] - h_off[e0]) * ob;
            PP_CUDA(P.obs.reserve(nob));
            SpanGuard t(M, PP_PROF_H2D, M->s_in);
            PP_CUDA(cudaMemcpyAsync(P.obs.p, static_cast<const char*>(obs) + h_off[e0] * ob, nob, cudaMemcpyHostToDevice, M->s_in));
            d_obs_base = static_cast<const char*>(P.obs.p) - h_off[e0] * ob;
        }
        PP_CUDA(cudaEventRecord(P.in_done, M->s_in));
        if (!pre_scheduled) PP_CUDA(cudaEventSynchronize(P.sch_done));
        const int64_t minlen = M->h_pin[16 + 4 * (c % kPipeSlots)], maxlen = M->h_pin[17 + 4 * (c % kPipeSlots)], total_rows = M->h_pin[18 + 4 * (c % kPipeSlots)];
        if (minlen < 1 || maxlen > 0x7ffffff0LL) { cudaDeviceSynchronize(); return bad_event_error(M, d_off, n_events, M->stream); }
        info[c] = {e0, e1, sc.n_groups, maxlen, d_obs_base};
        const size_t ws = static_cast<size_t>(total_rows) * row_bytes;
        if (ws > M->ws_limit) {
            set_error("one event needs %zu bytes of traceback workspace (limit %zu)", ws, M->ws_limit);
            cudaDeviceSynchronize();
            return PP_ERR_NOMEM;
        }
        if (want_vit && path_out) {
            const size_t tmp_elems = 4 * static_cast<size_t>((host || h_off) ? (h_off[e1] - h_off[e0]) : total_obs_dev) + 64 * static_cast<size_t>(e1 - e0);
            if (P.path_tmp.cap < tmp_elems * sizeof(uint16_t) || P.counter.cap < 16) {
                PP_CUDA(cudaStreamSynchronize(M->stream));
                PP_CUDA(P.counter.reserve(16));
                if (P.path_tmp.reserve(tmp_elems * sizeof(uint16_t)) != cudaSuccess) { cudaGetLastError(); set_error("cannot allocate %zu bytes of traceback scratch", tmp_elems * 2); return PP_ERR_NOMEM; }
            }
            M->h_pin[32 + (c % kPipeSlots)] = 0;
        }
        if (P.ws.cap < ws) {                                  // growing the table frees memory: nothing may be in flight on it
            PP_CUDA(cudaStreamSynchronize(M->stream));
            if (P.ws.reserve(ws) != cudaSuccess) { cudaGetLastError(); set_error("cannot allocate %zu bytes of workspace", ws); return PP_ERR_NOMEM; }
        }
        // ---- phase A: sweeps, path lengths, exclusive scan, the chunk's totals to pinned memory
        PP_CUDA(cudaStreamWaitEvent(M->stream, P.in_done, 0));
        const LaneBatchAny A{d_obs_base, obs_dtype == PP_F64 ? 1 : 0, d_off, P.order.as<int32_t>(), P.group_len.as<int32_t>(), P.group_row.as<int64_t>()};
        if (want_vit) {
            VitOut O{d_score, M->d_end_state.as<int32_t>(), M->d_end_ord.as<int32_t>(), P.ws.as<uint32_t>()};
            SpanGuard t(M, PP_PROF_VITERBI, M->stream);
            rc = launch_viterbi(M, A, O, sc.n_groups);
        }
        if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
        if (want_fwd) {
            {
                SpanGuard t(M, PP_PROF_FORWARD, M->stream);
                rc = launch_forward(M, A, FwdOut{d_fwd}, sc.n_groups);
            }
            if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
            rc = exact_forward_collect(M, c % kPipeSlots, d_fwd + e0, e1 - e0);
            if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
        }
        if (want_vit) {
            SpanGuard t(M, PP_PROF_TRACEBACK, M->stream);
            const int nslots = sc.n_groups * kLanes;
            if (path_out) {
                // one walk: states into per-event scratch regions (4 n + 64 entries each), lengths into d_plen
                PP_CUDA(cudaMemsetAsync(P.counter.p, 0, 4, M->stream));
                k_traceback_walk<<<(nslots + 127) / 128, 128, trace_smem_bytes(M->m, M->n_trace_src[use_profile(M) ? 1 : 0]), M->stream>>>(trace_args(c), d_plen, P.path_tmp.as<uint16_t>(),
                                                                              P.counter.as<int32_t>(), nslots);
                PP_CUDA(cudaMemcpyAsync(M->h_pin + 32 + (c % kPipeSlots), P.counter.p, 4, cudaMemcpyDeviceToHost, M->stream));
            } else {
                k_traceback<<<(nslots + 127) / 128, 128, trace_smem_bytes(M->m, M->n_trace_src[use_profile(M) ? 1 : 0]), M->stream>>>(trace_args(c), d_plen, nullptr, nullptr, nslots);
            }
            PP_CUDA(cudaGetLastError());
            ++M->launches;
            const int64_t ne = e1 - e0;                       // exclusive scan of max(len, 0) over the chunk's events
            auto it = thrust::make_transform_iterator(static_cast<const int64_t*>(d_plen + e0), ClampLen());
            size_t tmp = 0;
            int64_t* d_poff = M->d_path_off.as<int64_t>() + e0;
            PP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, d_poff, static_cast<int>(ne), M->stream));
            if (M->d_scan_tmp.cap < tmp + 16) {
                PP_CUDA(cudaStreamSynchronize(M->stream));
                PP_CUDA(M->d_scan_tmp.reserve(tmp + 16));
            }
            PP_CUDA(cub::DeviceScan::ExclusiveSum(M->d_scan_tmp.p, tmp, it, d_poff, static_cast<int>(ne), M->stream));
            ++M->launches;
            PP_CUDA(cudaMemcpyAsync(M->h_pin + (c % kPipeSlots) * 2, d_poff + ne - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, M->stream));
            PP_CUDA(cudaMemcpyAsync(M->h_pin + (c % kPipeSlots) * 2 + 1, d_plen + e1 - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, M->stream));
        }
        PP_CUDA(cudaEventRecord(P.a_done, M->stream));
        if (c >= 1) { rc = finish(c - 1); if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; } }
    }
    rc = finish(n_chunks - 1);
    if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
    if (host) {
        SpanGuard t(M, PP_PROF_D2H, M->stream);
        if (want_vit) {
            PP_CUDA(cudaMemcpyAsync(logp_out, d_score, n_events * sizeof(double), cudaMemcpyDeviceToHost, M->stream));
            PP_CUDA(cudaMemcpyAsync(path_len_out, d_plen, n_events * sizeof(int64_t), cudaMemcpyDeviceToHost, M->stream));
        }
        if (want_fwd) PP_CUDA(cudaMemcpyAsync(fwd_out, d_fwd, n_events * sizeof(double), cudaMemcpyDeviceToHost, M->stream));
    }
    PP_CUDA(cudaStreamSynchronize(M->stream));
    PP_CUDA(cudaStreamSynchronize(M->s_out));
    PP_CUDA(cudaStreamSynchronize(M->s_in));
    spans_collect(M);
    if (path_total) *path_total = base;
    if (overflow) {
        set_error("path buffer too small: %lld elements needed, capacity %lld", (long long)base, (long long)path_capacity);
        return PP_ERR_CAPACITY;
    }
    return PP_OK;
}

int pp_viterbi_batch(pp_model* M, const void* obs, int32_t obs_dtype, const int64_t* offsets, int64_t n_events,
                     int32_t mem, double* logp_out, int64_t* path_len_out, uint16_t* path_out, int64_t path_capacity,
                     int64_t* path_total) {
    if (!M || !obs || !logp_out || !path_len_out) { set_error("null argument"); return PP_ERR_ARG; }
    if (obs_dtype != PP_F32 && obs_dtype != PP_F64) { set_error("bad obs_dtype"); return PP_ERR_ARG; }
    if (mem != PP_MEM_HOST && mem != PP_MEM_DEVICE) { set_error("bad mem"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(M->device));
    std::memset(M->prof, 0, sizeof M->prof);
    if (path_total) *path_total = 0;
    if (n_events == 0) return PP_OK;
    if (!lane_kernels_fit(M))                         // large model: exact state-parallel kernels
        return viterbi_batch_state(M, obs, obs_dtype, offsets, n_events, mem, logp_out, path_len_out, path_out,
                                   path_capacity, path_total);
    return settle(M, score_pipeline(M, obs, obs_dtype, offsets, n_events, mem, logp_out, path_len_out, path_out, path_capacity, path_total, nullptr));
}

int pp_score_batch(pp_model* M, const void* obs, int32_t obs_dtype, const int64_t* offsets, int64_t n_events, int32_t mem,
                   double* viterbi_logp_out, int64_t* path_len_out, uint16_t* path_out, int64_t path_capacity,
                   int64_t* path_total, double* forward_logp_out) {
    if (!M || !obs || !viterbi_logp_out || !path_len_out || !forward_logp_out) { set_error("null argument"); return PP_ERR_ARG; }
    if (obs_dtype != PP_F32 && obs_dtype != PP_F64) { set_error("bad obs_dtype"); return PP_ERR_ARG; }
    if (mem != PP_MEM_HOST && mem != PP_MEM_DEVICE) { set_error("bad mem"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(M->device));
    if (path_total) *path_total = 0;
    if (n_events == 0) { std::memset(M->prof, 0, sizeof M->prof); return PP_OK; }
    if (!lane_kernels_fit(M)) {                       // models the lane kernels do not serve: the two single calls
        int rc = pp_viterbi_batch(M, obs, obs_dtype, offsets, n_events, mem, viterbi_logp_out, path_len_out, path_out, path_capacity, path_total);
        if (rc != PP_OK) return rc;
        double keep[PP_PROF_SLOTS];
        std::memcpy(keep, M->prof, sizeof keep);
        rc = pp_forward_batch(M, obs, obs_dtype, offsets, n_events, mem, forward_logp_out);
        for (int i = 0; i < PP_PROF_SLOTS; ++i) M->prof[i] += keep[i];
        return rc;
    }
    std::memset(M->prof, 0, sizeof M->prof);
    return settle(M, score_pipeline(M, obs, obs_dtype, offsets, n_events, mem, viterbi_logp_out, path_len_out, path_out, path_capacity,
                                    path_total, forward_logp_out));
}

// ---------------------------------------------------------------------------------------------
int pp_forward_batch(pp_model* M, const void* obs, int32_t obs_dtype, const int64_t* offsets, int64_t n_events,
                     int32_t mem, double* logp_out) {
    if (!M || !obs || !logp_out) { set_error("null argument"); return PP_ERR_ARG; }
    if (obs_dtype != PP_F32 && obs_dtype != PP_F64) { set_error("bad obs_dtype"); return PP_ERR_ARG; }
    if (mem != PP_MEM_HOST && mem != PP_MEM_DEVICE) { set_error("bad mem"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(M->device));
    std::memset(M->prof, 0, sizeof M->prof);
    if (n_events == 0) return PP_OK;
    if (lane_kernels_fit(M))
        return settle(M, score_pipeline(M, obs, obs_dtype, offsets, n_events, mem, nullptr, nullptr, nullptr, 0, nullptr, logp_out));
    // large model, or emissions from a table: every event takes the exact log-space kernel
    const size_t ob = obs_dtype == PP_F32 ? 4 : 8;
    const bool host = mem == PP_MEM_HOST;
    int rc = pipe_init(M);
    if (rc != PP_OK) return rc;
    std::vector<int64_t> store;
    const int64_t *h_off, *d_off;
    rc = host_offsets(M, offsets, n_events, mem, &store, &h_off, &d_off);
    if (rc != PP_OK) return rc;
    rc = pp_check_emission_table(M, h_off[n_events]);
    if (rc != PP_OK) return rc;
    double* d_logp = logp_out;
    if (host) {
        PP_CUDA(M->d_logp.reserve(n_events * sizeof(double)));
        d_logp = M->d_logp.as<double>();
    }
    PP_CUDA(cudaStreamSynchronize(M->stream));
    const std::vector<int64_t> cuts = make_chunks(h_off, n_events, 0, 0, host ? ob : 0, size_t(96) << 20);
    const int n_chunks = static_cast<int>(cuts.size()) - 1;
    std::vector<const void*> obs_base(n_chunks);
    auto finish = [&](int c) -> int {
        pp_model::PipeSlot& P = M->ps[c % kPipeSlots];
        PP_CUDA(cudaEventSynchronize(P.a_done));
        const int64_t e0 = cuts[c], e1 = cuts[c + 1];
        int64_t maxlen = 0;
        for (int64_t e = e0; e < e1; ++e) maxlen = std::max(maxlen, h_off[e + 1] - h_off[e]);
        int r = exact_forward_run(M, c % kPipeSlots, obs_base[c], obs_dtype, d_off + e0, maxlen, d_logp + e0);
        PP_CUDA(cudaEventRecord(P.tb_done, M->stream));
        return r;
    };
    for (int c = 0; c < n_chunks; ++c) {
        pp_model::PipeSlot& P = M->ps[c % kPipeSlots];
        const int64_t e0 = cuts[c], e1 = cuts[c + 1];
        PP_CUDA(cudaStreamWaitEvent(M->s_in, P.tb_done, 0));
        obs_base[c] = obs;
        if (host) {
            const size_t nob = static_cast<size_t>(h_off[e1] - h_off[e0]) * ob;
            PP_CUDA(P.obs.reserve(nob));
            SpanGuard t(M, PP_PROF_H2D, M->s_in);
            PP_CUDA(cudaMemcpyAsync(P.obs.p, static_cast<const char*>(obs) + h_off[e0] * ob, nob, cudaMemcpyHostToDevice, M->s_in));
            obs_base[c] = static_cast<const char*>(P.obs.p) - h_off[e0] * ob;
        }
        PP_CUDA(cudaEventRecord(P.in_done, M->s_in));
        PP_CUDA(cudaStreamWaitEvent(M->stream, P.in_done, 0));
        {
            SpanGuard t(M, PP_PROF_FORWARD, M->stream);
            PP_CUDA(cudaMemsetAsync(d_logp + e0, 0xFF, (e1 - e0) * sizeof(double), M->stream));     // all-ones = NaN: "not resolved"
        }
        rc = exact_forward_collect(M, c % kPipeSlots, d_logp + e0, e1 - e0);
        if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
        PP_CUDA(cudaEventRecord(P.a_done, M->stream));
        if (c >= 1) { rc = finish(c - 1); if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; } }
    }
    rc = finish(n_chunks - 1);
    if (rc != PP_OK) { cudaDeviceSynchronize(); return rc; }
    if (host) {
        SpanGuard t(M, PP_PROF_D2H, M->stream);
        PP_CUDA(cudaMemcpyAsync(logp_out, d_logp, n_events * sizeof(double), cudaMemcpyDeviceToHost, M->stream));
    }
    PP_CUDA(cudaStreamSynchronize(M->stream));
    PP_CUDA(cudaStreamSynchronize(M->s_in));
    spans_collect(M);
    return PP_OK;
}

}  // extern "C"
