// pp_state_kernels.cuh -- "state-parallel" float64 LOG-SPACE kernels: one CTA per event, one thread
// per state (strided), silent states swept level by level of the silent sub-DAG.
//
// These follow the reference literally -- same in/out-edge CSR order, same pairwise log-sum-exp
// (yahmm.pyx:52-70), same two silent passes -- and return the reference's own DP tables:
//   k_forward_table  : Model._forward  (yahmm.pyx:2810-2930)
//   k_backward_table : Model._backward (yahmm.pyx:2961-3156)
//   k_fb_accumulate  : Model._forward_backward / _train_once_baum_welch E-step
//                      (yahmm.pyx:3244-3321, 4141-4236) in sufficient-statistic form
// They are the exact path behind Model.forward()/backward()/forward_backward(), the re-run path for
// events the scaled linear-space forward sweep reports as -inf/NaN, and the path for models too
// large for the lane-per-event kernels' shared-memory column.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pp_device_common.cuh"

namespace pp {

struct StateModel {
    const int32_t *in_ptr, *in_src, *out_ptr, *out_dst;
    const double *in_logp, *out_logp;
    const EmisRec* emis;
    const double* etab;          // caller's emission table [row][ne] for EMIS_TABLE states (row = index into obs), or null
    const int32_t* lev_ptr;      // [n_levels + 1] silent states grouped by forward level (ascending)
    const int32_t* lev_states;
    const int32_t* rlev_ptr;     // [n_rlevels + 1] silent states grouped by reverse level
    const int32_t* rlev_states;
    int32_t m, ne, start, end, finite, n_levels, n_rlevels;
};

// e[i, l] (yahmm.pyx:2836-2840): evaluated here for Normal / point-mass / Uniform states, read from the caller's table
// for every other family; `row` is the observation's index in the batch's obs array
__device__ __forceinline__ double emit_state(const StateModel& S, int l, double x, int64_t row) {
    const EmisRec& R = S.emis[l];
    if (R.kind == EMIS_TABLE) return __dadd_rn(S.etab[row * S.ne + l], R.logw);
    return emission_log(R, x);
}

__device__ __forceinline__ double pair_lse(double x, double y) {       // yahmm.pyx:52-70
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    if (x == inf || y == inf) return inf;
    if (x == -inf) return y;
    if (y == -inf) return x;
    if (x > y) return x + log(exp(y - x) + 1);
    return y + log(exp(x - y) + 1);
}

// Silent states of one row, level by level of the silent sub-DAG.  A level wider than a warp is spread over the CTA
// and closed by a CTA barrier; a narrow level (the links of a skip / back-slip chain: one to three states) is done by
// warp 0 alone behind a warp barrier, so a chain of several hundred links costs several hundred __syncwarp instead of
// several hundred __syncthreads.  `body(l)` computes and stores state l.  Control flow is CTA-uniform.
template <typename Body>
__device__ __forceinline__ void sweep_levels(const int32_t* __restrict__ lev_ptr, const int32_t* __restrict__ lev_states,
                                             int n_levels, Body body) {
    const int t = threadIdx.x, nt = blockDim.x;
    bool pending = false;                         // warp 0 wrote values the rest of the CTA has not synchronised with
    for (int lev = 0; lev < n_levels; ++lev) {
        const int b = lev_ptr[lev], e = lev_ptr[lev + 1];
        if (e - b > 32) {
            if (pending) { __syncthreads(); pending = false; }
            for (int j = b + t; j < e; j += nt) body(lev_states[j]);
            __syncthreads();
        } else {
            if (t < 32) {
                if (b + t < e) body(lev_states[b + t]);
                __syncwarp();
            }
            pending = true;
        }
    }
    if (pending) __syncthreads();
}

// One event per CTA.  ROLL = false: table (n+1) x m row-major float64 (Model.forward's return value).  ROLL = true: only
// two rows are kept (row i at f + (i & 1) * m, shared memory) -- all a score needs.  e_row: scratch [ne] per CTA.
template <bool ROLL>
__device__ inline void forward_table_event(const StateModel& S, const double* __restrict__ x, int n, double* f,
                                           double* e_row, int64_t row0) {
    const int m = S.m, ne = S.ne;
    const int t = threadIdx.x, nt = blockDim.x;
    for (int l = t; l < m; l += nt) f[l] = (l == S.start) ? 0.0 : neg_inf();
    __syncthreads();
    sweep_levels(S.lev_ptr, S.lev_states, S.n_levels, [&](int l) {      // 2847-2871
        if (l == S.start) return;
        double lp = neg_inf();
        for (int k = S.in_ptr[l]; k < S.in_ptr[l + 1]; ++k) {
            const int ki = S.in_src[k];
            if (ki < ne || ki >= l) continue;
            lp = pair_lse(lp, f[ki] + S.in_logp[k]);
        }
        f[l] = lp;
    });
    for (int i = 0; i < n; ++i) {
        const double* fi = ROLL ? f + static_cast<size_t>(i & 1) * m : f + static_cast<size_t>(i) * m;
        double* fn = ROLL ? f + static_cast<size_t>((i + 1) & 1) * m : f + static_cast<size_t>(i + 1) * m;
        const double xi = x[i];
        for (int l = t; l < ne; l += nt) {                               // 2874-2889
            double lp = neg_inf();
            for (int k = S.in_ptr[l]; k < S.in_ptr[l + 1]; ++k) lp = pair_lse(lp, fi[S.in_src[k]] + S.in_logp[k]);
            fn[l] = lp + emit_state(S, l, xi, row0 + i);
        }
        __syncthreads();
        sweep_levels(S.lev_ptr, S.lev_states, S.n_levels, [&](int l) {
            double lp1 = neg_inf();                                      // 2891-2906
            for (int k = S.in_ptr[l]; k < S.in_ptr[l + 1]; ++k) {
                const int ki = S.in_src[k];
                if (ki >= ne) continue;
                lp1 = pair_lse(lp1, fn[ki] + S.in_logp[k]);
            }
            double lp2 = neg_inf();                                      // 2908-2926
            for (int k = S.in_ptr[l]; k < S.in_ptr[l + 1]; ++k) {
                const int ki = S.in_src[k];
                if (ki < ne || ki >= l) continue;
                lp2 = pair_lse(lp2, fn[ki] + S.in_logp[k]);
            }
            fn[l] = pair_lse(lp1, lp2);
        });
    }
    (void)e_row;
}

__device__ inline double logp_from_forward(const StateModel& S, const double* f, int n, bool roll = false) {   // 3378-3396
    const double* fn = f + static_cast<size_t>(roll ? (n & 1) : n) * S.m;
    if (S.finite == 1) return fn[S.end];
    double s = neg_inf();
    for (int i = 0; i < S.ne; ++i) s = pair_lse(s, fn[i]);
    return s;
}

__device__ inline void backward_table_event(const StateModel& S, const double* __restrict__ x, int n, double* b,
                                            double* e_row, int64_t row0) {
    const int m = S.m, ne = S.ne;
    const int t = threadIdx.x, nt = blockDim.x;
    double* bn = b + static_cast<size_t>(n) * m;
    if (S.finite == 1) {                                                 // 2996-3004
        for (int l = t; l < m; l += nt) bn[l] = (l == S.end) ? 0.0 : neg_inf();
    } else {
        for (int l = t; l < m; l += nt) bn[l] = (l < ne) ? 0.0 : neg_inf();
    }
    __syncthreads();
    if (S.finite == 1) {
        for (int lev = 0; lev < S.n_rlevels; ++lev) {                    // 3006-3037
            for (int j = S.rlev_ptr[lev] + t; j < S.rlev_ptr[lev + 1]; j += nt) {
                const int k = S.rlev_states[j];
                if (k == S.end) continue;
                double lp = neg_inf();
                for (int l = S.out_ptr[k]; l < S.out_ptr[k + 1]; ++l) {
                    const int li = S.out_dst[l];
                    if (li < k + 1) continue;
                    lp = pair_lse(lp, bn[li] + S.out_logp[l]);
                }
                bn[k] = lp;
            }
            __syncthreads();
        }
        for (int k = t; k < ne; k += nt) {                               // 3039-3061
            double lp = neg_inf();
            for (int l = S.out_ptr[k]; l < S.out_ptr[k + 1]; ++l) {
                const int li = S.out_dst[l];
                if (li < ne) continue;
                lp = pair_lse(lp, bn[li] + S.out_logp[l]);
            }
            bn[k] = lp;
        }
        __syncthreads();
    }
    for (int i = n - 1; i >= 0; --i) {                                   // 3064-3152
        double* bi = b + static_cast<size_t>(i) * m;
        const double* bp = b + static_cast<size_t>(i + 1) * m;
        const double xi = x[i];
        for (int l = t; l < ne; l += nt) e_row[l] = emit_state(S, l, xi, row0 + i);
        __syncthreads();
        for (int lev = 0; lev < S.n_rlevels; ++lev) {
            for (int j = S.rlev_ptr[lev] + t; j < S.rlev_ptr[lev + 1]; j += nt) {
                const int k = S.rlev_states[j];
                double lp1 = neg_inf();                                  // 3070-3094
                for (int l = S.out_ptr[k]; l < S.out_ptr[k + 1]; ++l) {
                    const int li = S.out_dst[l];
                    if (li >= ne) continue;
                    lp1 = pair_lse(lp1, bp[li] + S.out_logp[l] + e_row[li]);
                }
                double lp2 = neg_inf();                                  // 3096-3119
                for (int l = S.out_ptr[k]; l < S.out_ptr[k + 1]; ++l) {
                    const int li = S.out_dst[l];
                    if (li < k + 1) continue;
                    lp2 = pair_lse(lp2, bi[li] + S.out_logp[l]);
                }
                bi[k] = pair_lse(lp2, lp1);
            }
            __syncthreads();
        }
        for (int k = t; k < ne; k += nt) {                               // 3121-3152
            double lp = neg_inf();
            for (int l = S.out_ptr[k]; l < S.out_ptr[k + 1]; ++l) {
                const int li = S.out_dst[l];
                if (li >= ne) continue;
                lp = pair_lse(lp, bp[li] + S.out_logp[l] + e_row[li]);
            }
            for (int l = S.out_ptr[k]; l < S.out_ptr[k + 1]; ++l) {
                const int li = S.out_dst[l];
                if (li < ne) continue;
                lp = pair_lse(lp, bi[li] + S.out_logp[l]);
            }
            bi[k] = lp;
        }
        __syncthreads();
    }
}

// mode 0: forward table, mode 1: backward table.  Single event (grid = 1) or one event per CTA.
template <typename ObsT>
__global__ void k_table(const StateModel S, const ObsT* __restrict__ obs, const int64_t* __restrict__ offsets,
                        int n_events, double* __restrict__ tables, const int64_t* __restrict__ table_off,
                        double* __restrict__ xbuf, int64_t xbuf_stride, int mode) {
    extern __shared__ double e_row[];
    for (int ev = blockIdx.x; ev < n_events; ev += gridDim.x) {
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        double* x = xbuf + static_cast<size_t>(blockIdx.x) * xbuf_stride;
        for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = static_cast<double>(obs[off + i]);
        __syncthreads();
        double* tab = tables + table_off[ev];
        if (mode == 0) forward_table_event<false>(S, x, n, tab, e_row, off);
        else backward_table_event(S, x, n, tab, e_row, off);
        __syncthreads();
    }
}

// Exact log-space score of selected events (re-run path of the linear-space forward sweep): one
// CTA per listed event, forward table in per-CTA scratch, only log P is kept.
// `scratch` == nullptr: two rolling rows in dynamic shared memory (2 * m doubles) -- the normal case; otherwise the full
// table in per-CTA global scratch (models with more than ~14,000 states).
template <typename ObsT>
__global__ void k_forward_exact(const StateModel S, const ObsT* __restrict__ obs, const int64_t* __restrict__ offsets,
                                const int32_t* __restrict__ list, int n_list, double* __restrict__ scratch,
                                int64_t scratch_stride, double* __restrict__ xbuf, int64_t xbuf_stride,
                                double* __restrict__ logp_out) {
    extern __shared__ double e_row[];
    for (int j = blockIdx.x; j < n_list; j += gridDim.x) {
        const int ev = list[j];
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        double* x = xbuf + static_cast<size_t>(blockIdx.x) * xbuf_stride;
        for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = static_cast<double>(obs[off + i]);
        __syncthreads();
        if (scratch == nullptr) {
            double* f = e_row;                          // [2][m]
            forward_table_event<true>(S, x, n, f, nullptr, off);
            if (threadIdx.x == 0) logp_out[ev] = logp_from_forward(S, f, n, true);
        } else {
            double* f = scratch + static_cast<size_t>(blockIdx.x) * scratch_stride;
            forward_table_event<false>(S, x, n, f, e_row, off);
            if (threadIdx.x == 0) logp_out[ev] = logp_from_forward(S, f, n);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// State-parallel Viterbi (Model._viterbi, yahmm.pyx:3471-3652) for models whose column does not fit the
// lane-per-event kernels: one CTA per event, one thread per state, two float64 rows in shared memory,
// silent states level by level.  Every state scans its lowered in-edge list (reference visiting order) with
// the reference's operations -- (v + t) + e / v + t, strict '>' -- so scores and paths are bit-identical
// too; the winning ordinal per cell goes to HBM as uint16, [event][row][state].
// ------------------------------------------------------------------------------------------
struct StateVit {
    const int32_t* edge_begin;   // [m] first lowered edge of a state
    const int32_t* n_edges;      // [m]
    const int32_t* low_src;      // [lowered edges]
    const double* low_logp;      // [lowered edges]
    int32_t end_final_only;
};

template <typename ObsT>
__global__ void k_viterbi_state(const StateModel S, const StateVit V, const ObsT* __restrict__ obs,
                                const int64_t* __restrict__ offsets, int ev0, int n_events,
                                const int64_t* __restrict__ ptr_off, uint16_t* __restrict__ ptr,
                                double* __restrict__ score_out, int32_t* __restrict__ end_state_out) {
    extern __shared__ double rows[];                   // [2][m]
    const int m = S.m, ne = S.ne, t = threadIdx.x, nt = blockDim.x;
    for (int j = blockIdx.x; j < n_events; j += gridDim.x) {
        const int ev = ev0 + j;
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        uint16_t* tb = ptr + ptr_off[j];
        double* cur = rows;
        double* prv = rows + m;
        for (int r = 0; r <= n; ++r) {
            uint16_t* tr = tb + static_cast<size_t>(r) * m;
            if (r == 0) {
                for (int l = t; l < m; l += nt) cur[l] = (l == S.start) ? 0.0 : neg_inf();
            } else {
                const double x = static_cast<double>(obs[off + r - 1]);
                for (int l = t; l < ne; l += nt) {                                   // 3545-3562
                    const double e = emit_state(S, l, x, off + r - 1);
                    double best = neg_inf();
                    int arg = 0;
                    const int b = V.edge_begin[l];
                    for (int k = 0; k < V.n_edges[l]; ++k) {
                        const double s = __dadd_rn(__dadd_rn(prv[V.low_src[b + k]], V.low_logp[b + k]), e);
                        if (s > best) { best = s; arg = k; }
                    }
                    cur[l] = best;
                    tr[l] = static_cast<uint16_t>(arg);
                }
            }
            __syncthreads();
            sweep_levels(S.lev_ptr, S.lev_states, S.n_levels, [&](int l) {            // 3515-3542, 3564-3605
                if (r == 0 && l == S.start) return;
                if (V.end_final_only && l == S.end && r != n) return;
                double best = neg_inf();
                int arg = 0;
                const int b = V.edge_begin[l];
                for (int k = 0; k < V.n_edges[l]; ++k) {
                    const double s = __dadd_rn(cur[V.low_src[b + k]], V.low_logp[b + k]);
                    if (s > best) { best = s; arg = k; }
                }
                cur[l] = best;
                tr[l] = static_cast<uint16_t>(arg);
            });
            double* sw = cur; cur = prv; prv = sw;                                    // prv = the row just finished
        }
        if (t == 0) {                                                                // 3614-3619
            int est = S.end;
            double score;
            if (S.finite == 1) {
                score = prv[S.end];
            } else {
                est = 0;
                score = prv[0];
                for (int k = 1; k < m; ++k) if (prv[k] > score) { score = prv[k]; est = k; }
            }
            score_out[ev] = score;
            end_state_out[ev] = est;
        }
        __syncthreads();
    }
}

// traceback over the [event][row][state] uint16 table (3632-3646); pass 1 counts, pass 2 writes
static __global__ void k_traceback_state(const uint16_t* __restrict__ ptr, const int64_t* __restrict__ ptr_off,
                                         const int64_t* __restrict__ offsets, int ev0, int n_events,
                                         const double* __restrict__ score, const int32_t* __restrict__ end_state,
                                         const int32_t* __restrict__ edge_begin, const int32_t* __restrict__ low_src,
                                         int m, int ne, int start, int64_t* __restrict__ path_len,
                                         const int64_t* __restrict__ path_off, uint16_t* __restrict__ path) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_events) return;
    const int ev = ev0 + j;
    if (score[ev] == neg_inf()) {
        if (!path) path_len[ev] = -1;
        return;
    }
    const int n = static_cast<int>(offsets[ev + 1] - offsets[ev]);
    const uint16_t* tb = ptr + ptr_off[j];
    int px = n, py = end_state[ev];
    int64_t len = 1, pos = 0;
    uint16_t* out = nullptr;
    if (path) { pos = path_len[ev] - 1; out = path + path_off[ev]; }
    while (px != 0 || py != start) {
        if (out) out[pos--] = static_cast<uint16_t>(py);
        const int ord = tb[static_cast<size_t>(px) * m + py];
        const int src = low_src[edge_begin[py] + ord];
        if (py < ne) --px;
        py = src;
        ++len;
    }
    if (out) out[pos] = static_cast<uint16_t>(py);
    else path_len[ev] = len;
}

// atomic min / max on doubles through the order-preserving integer view
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(static_cast<long long>(old)) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(static_cast<long long>(old)) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}

// Forward-backward statistics of a batch: one event per CTA (grid-stride), forward and backward
// tables in per-CTA scratch, then
//   edge_counts[out-edge] += sum_i exp(f[i,k] + t + e[i,l] + b[i+1,l] - logZ)   (emitting l, 3246-3271)
//                            sum_i exp(f[i,k] + t + b[i,l] - logZ)              (silent l,   3273-3302)
//   emit_moments[k]       += (sum w, sum w x, sum w x^2), w_i = exp(f[i+1,k] + b[i+1,k] - logZ) (4215-4229)
//   emit_minmax[k]         = min / max over whole events with sum w != 0        (304-326)
//   posteriors[(off+i)*ne + k] = f[i+1,k] + b[i+1,k] - logZ                     (3304-3321)
// Events with logZ == -inf contribute nothing (3240-3242, 4131-4133).
template <typename ObsT>
__global__ void k_fb_accumulate(const StateModel S, const ObsT* __restrict__ obs, const int64_t* __restrict__ offsets,
                                int n_events, double* __restrict__ scratch, int64_t scratch_stride,
                                double* __restrict__ xbuf, int64_t xbuf_stride, double* __restrict__ logp_out,
                                double* __restrict__ edge_counts, double* __restrict__ emit_moments,
                                double* __restrict__ emit_minmax, double* __restrict__ posteriors,
                                const int32_t* __restrict__ list) {
    extern __shared__ double e_row[];
    __shared__ double s_lo, s_hi;
    const int m = S.m, ne = S.ne, t = threadIdx.x, nt = blockDim.x;
    for (int idx = blockIdx.x; idx < n_events; idx += gridDim.x) {
        const int ev = list ? list[idx] : idx;                 // optional event list (re-run of the lane kernel's rejects)
        const int64_t off = offsets[ev];
        const int n = static_cast<int>(offsets[ev + 1] - off);
        double* x = xbuf + static_cast<size_t>(blockIdx.x) * xbuf_stride;
        for (int i = t; i < n; i += nt) x[i] = static_cast<double>(obs[off + i]);
        __syncthreads();
        double* f = scratch + static_cast<size_t>(blockIdx.x) * scratch_stride;
        double* b = f + static_cast<size_t>(n + 1) * m;
        forward_table_event<false>(S, x, n, f, e_row, off);
        const double lsp = logp_from_forward(S, f, n);
        if (t == 0) logp_out[ev] = lsp;
        if (lsp == neg_inf() || lsp != lsp) { __syncthreads(); continue; }
        backward_table_event(S, x, n, b, e_row, off);
        if (t == 0) {
            double lo = x[0], hi = x[0];
            for (int i = 1; i < n; ++i) { lo = fmin(lo, x[i]); hi = fmax(hi, x[i]); }
            s_lo = lo; s_hi = hi;
        }
        __syncthreads();
        // expected transition counts, one out-edge per thread
        const int E = S.out_ptr[m];
        for (int eidx = t; eidx < E; eidx += nt) {
            int lo_k = 0, hi_k = m;                                      // source state of this out-edge
            while (hi_k - lo_k > 1) { const int mid = (lo_k + hi_k) >> 1; if (S.out_ptr[mid] <= eidx) lo_k = mid; else hi_k = mid; }
            const int k = lo_k, li = S.out_dst[eidx];
            const double tl = S.out_logp[eidx];
            double acc = 0.0;
            if (li < ne) {
                for (int i = 0; i < n; ++i) {
                    const double term = f[static_cast<size_t>(i) * m + k] + tl + emit_state(S, li, x[i], off + i) +
                                        b[static_cast<size_t>(i + 1) * m + li];
                    if (term != neg_inf()) acc += exp(term - lsp);
                }
            } else {
                for (int i = 0; i <= n; ++i) {
                    const double term = f[static_cast<size_t>(i) * m + k] + tl + b[static_cast<size_t>(i) * m + li];
                    if (term != neg_inf()) acc += exp(term - lsp);
                }
            }
            if (acc != 0.0) atomicAdd(edge_counts + eidx, acc);
        }
        for (int k = t; k < ne; k += nt) {
            double sw = 0, swx = 0, swxx = 0;
            for (int i = 0; i < n; ++i) {
                const double lw = f[static_cast<size_t>(i + 1) * m + k] + b[static_cast<size_t>(i + 1) * m + k] - lsp;
                if (posteriors) posteriors[(off + i) * ne + k] = lw;
                const double w = exp(lw);
                sw += w; swx += w * x[i]; swxx += w * x[i] * x[i];
            }
            if (sw != 0.0) {
                atomicAdd(emit_moments + 3 * k, sw);
                atomicAdd(emit_moments + 3 * k + 1, swx);
                atomicAdd(emit_moments + 3 * k + 2, swxx);
                atomic_min_f64(emit_minmax + 2 * k, s_lo);
                atomic_max_f64(emit_minmax + 2 * k + 1, s_hi);
            }
        }
        __syncthreads();
    }
}

// Length-conditioned ancestral sampler: walk the out-edge CSR from start; while fewer
// than L symbols are out, edges into `end` (and into states from which no emission is reachable
// any more) are masked and the rest renormalised; stop after the L-th emission.
__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t& s) { return (splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static __global__ void k_sample(const StateModel S, const double* __restrict__ out_prob, const int32_t* __restrict__ alive,
                         const int64_t* __restrict__ offsets, int64_t n_events, uint64_t seed, float* __restrict__ obs_out) {
    const int64_t ev = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (ev >= n_events) return;
    const int64_t off = offsets[ev];
    const int64_t L = offsets[ev + 1] - off;
    uint64_t rng = seed * 0xD1342543DE82EF95ULL + static_cast<uint64_t>(ev) * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;
    splitmix64(rng);
    int s = S.start;
    int64_t emitted = 0;
    int restarts = 0;
    while (emitted < L) {
        if (s < S.ne) {
            const EmisRec R = S.emis[s];
            double v;
            if (R.kind == EMIS_NORMAL) {
                const double sigma = sqrt(R.a1 * 0.5);
                const double u1 = fmax(u01(rng), 1e-300), u2 = u01(rng);
                v = R.a0 + sigma * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
            } else if (R.kind == EMIS_POINT) {
                v = R.a0;
            } else {
                v = R.a0 + (R.a1 - R.a0) * u01(rng);
            }
            obs_out[off + emitted] = static_cast<float>(v);
            if (++emitted == L) break;
        }
        double total = 0.0;
        for (int e = S.out_ptr[s]; e < S.out_ptr[s + 1]; ++e) {
            const int d = S.out_dst[e];
            if (d != S.end && alive[d]) total += out_prob[e];
        }
        if (!(total > 0.0)) {                       // dead end: start the event over
            if (++restarts > 1000) { for (int64_t i = emitted; i < L; ++i) obs_out[off + i] = 0.f; return; }
            s = S.start;
            emitted = 0;
            continue;
        }
        const double u = u01(rng) * total;
        double c = 0.0;
        int nxt = -1;
        for (int e = S.out_ptr[s]; e < S.out_ptr[s + 1]; ++e) {
            const int d = S.out_dst[e];
            if (d == S.end || !alive[d]) continue;
            nxt = d;
            c += out_prob[e];
            if (c > u) break;
        }
        s = nxt;
    }
}

}  // namespace pp
