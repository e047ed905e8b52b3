// pp_lane_kernels.cuh -- "lane-per-event" column kernels for sm_100a (runtime-schedule variant).
//
// One CTA = W warps x 32 lanes sweeps 32 events.  Lane j of every warp works on event j; the warps
// split the states of a column (pp_lowering.h).  Consequences:
//   * the per-warp program (items, edges, emission records) is warp-uniform: record loads are
//     broadcast loads, there is no divergence and no per-lane CSR walk;
//   * column accesses are [slot][lane]: stride-1 across lanes, bank-conflict free;
//   * a lane only ever touches ITS event's column, so the serial silent chain of a column is plain
//     program order in one thread -- exact reference arithmetic, no scan;
//   * two __syncthreads per row: after the emitting phase (silent states read the new emitting
//     values) and after the silent phase (next row's emitting states read the silent values).
//
// Viterbi (k_viterbi_lane): float64 max-plus sweep that follows Model._viterbi
// (yahmm.pyx:3471-3652) operation for operation -- candidate (v + t) + e for emitting states,
// v + t for silent ones, strict '>' in the reference's visiting order -- and writes the ordinal of
// the winning in-edge (uint8) per cell to HBM for the traceback kernel.
//
// Forward (k_forward_lane): Model._forward / _log_probability (yahmm.pyx:2810-2930, 3378-3396) in
// scaled LINEAR space, float64: one DFMA per edge instead of the reference's exp + log per edge.
// Each event carries a power-of-two scale that is re-derived every row from the previous row's
// maximum, so the mantissas are untouched by rescaling (DESIGN.md "forward").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pp_lowering.h"

namespace pp {

struct LaneSched {
    const ItemRec* items[2];
    const EdgeRec* edges[2];
    const EmisRec* emis;
    const int32_t* warp_ranges;      // [W][4]
    int32_t m, ne, ns, n_slots;
    int32_t start_slot_off;          // byte offset of the start state's slot
    int32_t end_slot_off;            // byte offset of the end state's slot
    int32_t end_state, end_item, finite, end_final_only;
    int32_t n_ptr_words;             // 32-bit traceback pointer words per lane per row
};

template <typename ObsT>
struct LaneBatch {
    const ObsT* obs;                 // all observations back to back
    const int64_t* offsets;          // [n_events + 1]
    const int32_t* order;            // [n_groups * 32] event id per (group, lane), -1 = padding
    const int32_t* group_len;        // [n_groups] longest event of the group
    const int64_t* group_row;        // [n_groups] first pointer-table row of the group
};

struct VitOut {
    double* score;                   // [n_events]
    int32_t* end_state;              // [n_events] state the traceback starts from
    int32_t* end_ord;                // [n_events] winning in-edge ordinal of a final-only end state
    uint32_t* ptr;                   // [rows][n_ptr_words][32] packed winning in-edge ordinals
};

struct FwdOut {
    double* logp;                    // [n_events]
};

constexpr int kTileCols = 32;
constexpr int kTileStride = 33;      // +1 pad: conflict-free transposed fill

__host__ __device__ inline size_t lane_smem_bytes(int n_slots, int obs_bytes, int warps) {
    return static_cast<size_t>(n_slots) * kSlotBytes + static_cast<size_t>(kTileCols) * kTileStride * obs_bytes +
           static_cast<size_t>(warps) * kLanes * sizeof(int32_t);
}

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xFFF0000000000000LL); }

template <typename T>
__device__ __forceinline__ T ldg_rec(const T* p) {
    // 16-byte vector loads of a warp-uniform record through the read-only path
    static_assert(sizeof(T) % 16 == 0, "record size");
    T out;
    const int4* s = reinterpret_cast<const int4*>(p);
    int4* dst = reinterpret_cast<int4*>(&out);
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(T) / 16); ++i) dst[i] = __ldg(s + i);
    return out;
}

// NormalDistribution._log_probability (yahmm.pyx:377-390), UniformDistribution._log_probability
// (257-262), plus log(state.weight) (2839-2840).  Every operation is an explicitly rounded IEEE
// float64 op (no FMA contraction).  The division (x-mu)^2 / (2 sigma^2) is evaluated as a
// Markstein sequence q0 = a*y, r = fma(-q0, b, a), q = fma(r, y, q0) with y = RN(1/b): it returns
// the correctly rounded quotient (checked against IEEE division on 2e8 random operands and on
// every model constant in tests/test_emission.py) at 3 flops instead of a ~30-instruction divide.
__device__ __forceinline__ double emission_log(const EmisRec& R, double x) {
    double r;
    if (R.kind == EMIS_NORMAL) {
        const double d = __dsub_rn(x, R.a0);
        const double d2 = __dmul_rn(d, d);
        const double q0 = __dmul_rn(d2, R.a2);
        const double rem = __fma_rn(-q0, R.a1, d2);
        const double q = __fma_rn(rem, R.a2, q0);
        r = __dsub_rn(R.a3, q);
    } else if (R.kind == EMIS_POINT) {
        r = (fabs(__dsub_rn(x, R.a0)) < 1E-4) ? 0.0 : neg_inf();
    } else {
        r = (x == R.a0 && x == R.a1) ? 0.0 : ((x >= R.a0 && x <= R.a1) ? R.a3 : neg_inf());
    }
    return __dadd_rn(r, R.logw);
}

// 2^(t + shift) as a double, t in the log2 domain.  The fraction goes through MUFU.EX2 (fp32,
// rel. error ~2^-22); the integer part is written straight into the exponent field.  Below 2^-1000
// the factor is flushed to zero, above 2^1000 it becomes +inf (caught by the caller's status).
__device__ __forceinline__ double exp2_scaled(double t) {
    if (!(t > -1000.0)) return 0.0;                  // also catches -inf and NaN
    if (t > 1000.0) return __longlong_as_double(0x7FF0000000000000LL);
    const double n = rint(t);
    const float f = static_cast<float>(t - n);
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f));
    const int hi = (static_cast<int>(n) + 1023) << 20;
    return static_cast<double>(p) * __hiloint2double(hi, 0);
}

__device__ __forceinline__ double emission_log2(const EmisRec& R, double x) {
    if (R.kind == EMIS_NORMAL) {
        const double d = x - R.a0;
        return fma(-(d * d), R.k2, R.c2);
    }
    if (R.kind == EMIS_POINT) return (fabs(x - R.a0) < 1E-4) ? R.c2 : neg_inf();
    if (x == R.a0 && x == R.a1) return R.logw * 1.4426950408889634074;
    return (x >= R.a0 && x <= R.a1) ? R.c2 : neg_inf();
}

// Stage the next 32 observations of the CTA's 32 events: warp w loads events w, w+W, ... with one
// fully coalesced 32-wide load each and stores them transposed ([column][event]).
template <typename ObsT, int W>
__device__ __forceinline__ void fill_tile(ObsT* xt, const ObsT* __restrict__ obs, int64_t off, int n, int i0,
                                          int warp, int lane) {
#pragma unroll
    for (int j = warp; j < kLanes; j += W) {
        const int64_t oj = __shfl_sync(0xffffffffu, off, j);
        const int nj = __shfl_sync(0xffffffffu, n, j);
        const int c = i0 + lane;
        xt[lane * kTileStride + j] = (c < nj) ? __ldg(obs + oj + c) : ObsT(0);
    }
}

// ------------------------------------------------------------------------------------------
// Viterbi
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void vit_item(const ItemRec& I, const EdgeRec* __restrict__ edges,
                                         const unsigned char* colb, double e, bool emitting, double& best,
                                         int& arg) {
    best = neg_inf();
    arg = 0;
    const EdgeRec* ep = edges + I.edge_begin;
#pragma unroll 4
    for (int k = 0; k < I.n_edges; ++k) {
        const EdgeRec E = ldg_rec(ep + k);
        const double v = *reinterpret_cast<const double*>(colb + E.off);
        double s = __dadd_rn(v, E.val);
        if (emitting) s = __dadd_rn(s, e);
        if (s > best) { best = s; arg = k; }
    }
}

template <typename ObsT, int W>
__global__ void __launch_bounds__(W * 32, 2)
k_viterbi_lane(const LaneSched S, const LaneBatch<ObsT> B, const VitOut O) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    ObsT* xt = reinterpret_cast<ObsT*>(smem + static_cast<size_t>(S.n_slots) * kSlotBytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    unsigned char* colb = smem + lane * 8;                       // this lane's column slice
    uint32_t* prow = O.ptr + (static_cast<size_t>(B.group_row[g]) * S.n_ptr_words) * kLanes + lane;
    const size_t row_stride = static_cast<size_t>(S.n_ptr_words) * kLanes;
    uint32_t pw = 0;                                             // pointer word being assembled

    const int eb = S.warp_ranges[warp * 4 + 0], ee = S.warp_ranges[warp * 4 + 1];
    const int sb = S.warp_ranges[warp * 4 + 2], se = S.warp_ranges[warp * 4 + 3];

    // row 0: v[0, :] = -inf, v[0, start] = 0 (3513-3514), then the silent closure (3515-3542)
    for (int s = threadIdx.x; s < S.n_slots * kLanes; s += W * 32) col[s] = neg_inf();
    __syncthreads();
    if (warp == 0) *reinterpret_cast<double*>(colb + S.start_slot_off) = 0.0;
    __syncthreads();
    {
        const ItemRec* items = S.items[0];
        for (int it = sb; it < se; ++it) {
            const ItemRec I = ldg_rec(items + it);
            if (I.flags & ITEM_START) {
                if (I.flags & ITEM_PFLUSH) {
                    if (ev >= 0) prow[static_cast<size_t>(I.pword) * kLanes] = pw;
                    pw = 0;
                }
                continue;
            }
            double best; int arg;
            vit_item(I, S.edges[0], colb, 0.0, false, best, arg);
            *reinterpret_cast<double*>(colb + I.dst_off) = best;
            pw |= static_cast<uint32_t>(arg) << I.pshift;
            if (I.flags & ITEM_PFLUSH) {
                if (ev >= 0) prow[static_cast<size_t>(I.pword) * kLanes] = pw;
                pw = 0;
            }
        }
    }
    __syncthreads();

    for (int r = 1; r <= nmax; ++r) {
        const int q = r & 1;
        if (((r - 1) & (kTileCols - 1)) == 0) {
            fill_tile<ObsT, W>(xt, B.obs, off, n, r - 1, warp, lane);
            __syncthreads();
        }
        const double x = static_cast<double>(xt[((r - 1) & (kTileCols - 1)) * kTileStride + lane]);
        const bool active = r <= n;
        const ItemRec* items = q ? S.items[1] : S.items[0];
        const EdgeRec* edges = q ? S.edges[1] : S.edges[0];
        uint32_t* pr = prow + static_cast<size_t>(r) * row_stride;

        for (int it = eb; it < ee; ++it) {                           // emitting states (3545-3562)
            const ItemRec I = ldg_rec(items + it);
            const EmisRec R = ldg_rec(S.emis + I.emis);
            const double e = emission_log(R, x);
            double best; int arg;
            vit_item(I, edges, colb, e, true, best, arg);
            *reinterpret_cast<double*>(colb + I.dst_off) = best;
            pw |= static_cast<uint32_t>(arg) << I.pshift;
            if (I.flags & ITEM_PFLUSH) {
                if (active) pr[static_cast<size_t>(I.pword) * kLanes] = pw;
                pw = 0;
            }
        }
        __syncthreads();
        for (int it = sb; it < se; ++it) {                           // silent states (3564-3605)
            const ItemRec I = ldg_rec(items + it);
            double best; int arg;
            vit_item(I, edges, colb, 0.0, false, best, arg);
            *reinterpret_cast<double*>(colb + I.dst_off) = best;
            pw |= static_cast<uint32_t>(arg) << I.pshift;
            if (I.flags & ITEM_PFLUSH) {
                if (active) pr[static_cast<size_t>(I.pword) * kLanes] = pw;
                pw = 0;
            }
        }
        __syncthreads();

        // events whose last row this is: final score (3614-3619).  Warp 0 does it while the others
        // already run the next row's emitting phase (they write E[q^1]; this reads E[q] and SIL).
        if (warp == 0 && __any_sync(0xffffffffu, r == n)) {
            double score; int est = S.end_state, eord = 0;
            if (S.finite && S.end_final_only) {
                const ItemRec I = ldg_rec(items + S.end_item);
                vit_item(I, edges, colb, 0.0, false, score, eord);
            } else if (S.finite) {
                score = *reinterpret_cast<const double*>(colb + S.end_slot_off);
            } else {                                                 // numpy.argmax(v[n]) : first maximum
                score = *reinterpret_cast<const double*>(colb + static_cast<size_t>(S.ns + q * S.ne) * kSlotBytes);
                est = 0;
                for (int k = 1; k < S.m; ++k) {
                    const size_t slot = k < S.ne ? static_cast<size_t>(S.ns + q * S.ne + k) : static_cast<size_t>(k - S.ne);
                    const double v = *reinterpret_cast<const double*>(colb + slot * kSlotBytes);
                    if (v > score) { score = v; est = k; }
                }
            }
            if (r == n) {
                O.score[ev] = score;
                O.end_state[ev] = est;
                O.end_ord[ev] = eord;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Traceback (yahmm.pyx:3632-3646): one thread per event walks the ordinals back from (n, end).
// Pass 1 (path == nullptr) counts, pass 2 writes the path front-to-back at path_off[ev].
// ------------------------------------------------------------------------------------------
struct TraceArgs {
    const uint32_t* ptr;
    const int32_t* order;
    const int64_t* group_row;
    const int64_t* offsets;
    const double* score;
    const int32_t* end_state;
    const int32_t* end_ord;
    const int32_t* state_edge_begin;   // [m]
    const int32_t* low_src;            // [lowered edges]
    const int32_t* state_pword;        // [m] pointer word of a state's ordinal
    const int32_t* state_pshift;       // [m] bit position inside the word
    int32_t m, ne, start, end_final_only, n_ptr_words, ptr_mask;
};

static __global__ void k_traceback(const TraceArgs T, int64_t* __restrict__ path_len, const int64_t* __restrict__ path_off,
                            uint16_t* __restrict__ path, int n_slots) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int ev = T.order[slot];
    if (ev < 0) return;
    const int g = slot >> 5, lane = slot & 31;
    if (T.score[ev] == neg_inf()) {                      // impossible: (-inf, None) (3621-3623)
        if (!path) path_len[ev] = -1;
        return;
    }
    const int n = static_cast<int>(T.offsets[ev + 1] - T.offsets[ev]);
    const uint32_t* base = T.ptr + (static_cast<size_t>(T.group_row[g]) * T.n_ptr_words) * kLanes + lane;
    const size_t row_stride = static_cast<size_t>(T.n_ptr_words) * kLanes;
    int64_t len = 1;
    int px = n, py = T.end_state[ev];
    int64_t pos = 0;
    uint16_t* out = nullptr;
    if (path) {
        pos = path_len[ev] - 1;
        out = path + path_off[ev];
    }
    bool first = true;
    while (px != 0 || py != T.start) {
        if (out) out[pos--] = static_cast<uint16_t>(py);
        int ord;
        if (first && T.end_final_only) ord = T.end_ord[ev];
        else ord = (base[static_cast<size_t>(px) * row_stride + static_cast<size_t>(T.state_pword[py]) * kLanes] >>
                    T.state_pshift[py]) & T.ptr_mask;
        first = false;
        const int src = T.low_src[T.state_edge_begin[py] + ord];
        if (py < T.ne) --px;                             // emitting states point into the previous row
        py = src;
        ++len;
    }
    if (out) out[pos] = static_cast<uint16_t>(py);
    else path_len[ev] = len;
}

// ------------------------------------------------------------------------------------------
// Forward (scaled linear space)
// ------------------------------------------------------------------------------------------
constexpr int kFwdTargetExp = 900;   // a row's largest value is steered to ~2^900

__device__ __forceinline__ double fwd_item(const ItemRec& I, const EdgeRec* __restrict__ edges,
                                           const unsigned char* colb) {
    double acc = 0.0;
    const EdgeRec* ep = edges + I.edge_begin;
#pragma unroll 4
    for (int k = 0; k < I.n_edges; ++k) {
        const EdgeRec E = ldg_rec(ep + k);
        acc = fma(*reinterpret_cast<const double*>(colb + E.off), E.val, acc);
    }
    return acc;
}

template <typename ObsT, int W>
__global__ void __launch_bounds__(W * 32, 2)
k_forward_lane(const LaneSched S, const LaneBatch<ObsT> B, const FwdOut O) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    ObsT* xt = reinterpret_cast<ObsT*>(smem + static_cast<size_t>(S.n_slots) * kSlotBytes);
    int32_t* wmax = reinterpret_cast<int32_t*>(smem + static_cast<size_t>(S.n_slots) * kSlotBytes +
                                               static_cast<size_t>(kTileCols) * kTileStride * sizeof(ObsT));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    unsigned char* colb = smem + lane * 8;

    const int eb = S.warp_ranges[warp * 4 + 0], ee = S.warp_ranges[warp * 4 + 1];
    const int sb = S.warp_ranges[warp * 4 + 2], se = S.warp_ranges[warp * 4 + 3];

    for (int s = threadIdx.x; s < S.n_slots * kLanes; s += W * 32) col[s] = 0.0;
    __syncthreads();
    if (warp == 0) *reinterpret_cast<double*>(colb + S.start_slot_off) = 1.0;
    __syncthreads();
    int mx = 0;                                   // max of the high words of this warp's new values
    {
        const ItemRec* items = S.items[0];
        for (int it = sb; it < se; ++it) {
            const ItemRec I = ldg_rec(items + it);
            if (I.flags & ITEM_START) { mx = max(mx, __double2hiint(1.0)); continue; }
            const double v = fwd_item(I, S.edges[0], colb);
            *reinterpret_cast<double*>(colb + I.dst_off) = v;
            mx = max(mx, __double2hiint(v));
        }
    }
    wmax[warp * kLanes + lane] = mx;
    __syncthreads();

    long long scale_total = 0;                    // sum of the power-of-two shifts applied so far
    for (int r = 1; r <= nmax; ++r) {
        const int q = r & 1;
        // shift for this row from the previous row's maximum (exact power of two)
        int M = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) M = max(M, wmax[w * kLanes + lane]);
        const int ex = (M >> 20) & 0x7ff;
        const int shift = (ex == 0 || ex == 0x7ff) ? 0 : (kFwdTargetExp + 1023 - ex);
        scale_total += shift;
        if (((r - 1) & (kTileCols - 1)) == 0) {
            fill_tile<ObsT, W>(xt, B.obs, off, n, r - 1, warp, lane);
        }
        __syncthreads();                          // tile visible; everyone has read wmax
        const double x = static_cast<double>(xt[((r - 1) & (kTileCols - 1)) * kTileStride + lane]);
        const ItemRec* items = q ? S.items[1] : S.items[0];
        const EdgeRec* edges = q ? S.edges[1] : S.edges[0];
        mx = 0;
        for (int it = eb; it < ee; ++it) {                           // emitting states (2874-2889)
            const ItemRec I = ldg_rec(items + it);
            const EmisRec R = ldg_rec(S.emis + I.emis);
            const double p = exp2_scaled(emission_log2(R, x) + static_cast<double>(shift));
            const double v = fwd_item(I, edges, colb) * p;
            *reinterpret_cast<double*>(colb + I.dst_off) = v;
            mx = max(mx, __double2hiint(v));
        }
        __syncthreads();
        for (int it = sb; it < se; ++it) {                           // silent states (2891-2926)
            const ItemRec I = ldg_rec(items + it);
            const double v = fwd_item(I, edges, colb);
            *reinterpret_cast<double*>(colb + I.dst_off) = v;
            mx = max(mx, __double2hiint(v));
        }
        wmax[warp * kLanes + lane] = mx;
        __syncthreads();

        if (warp == 0 && __any_sync(0xffffffffu, r == n)) {          // log_probability (3378-3396)
            double P;
            if (S.finite && S.end_final_only) {
                const ItemRec I = ldg_rec(items + S.end_item);
                P = fwd_item(I, edges, colb);
            } else if (S.finite) {
                P = *reinterpret_cast<const double*>(colb + S.end_slot_off);
            } else {
                P = 0.0;
                for (int k = 0; k < S.ne; ++k)
                    P += *reinterpret_cast<const double*>(colb + static_cast<size_t>(S.ns + q * S.ne + k) * kSlotBytes);
            }
            if (r == n) {
                // P == 0 -> -inf (impossible, or every path underflowed: the host re-runs such events
                // with the exact log-space kernel); inf / NaN -> NaN, same treatment.
                double lp;
                if (P == 0.0) lp = neg_inf();
                else if (!(P < __longlong_as_double(0x7FF0000000000000LL))) lp = __longlong_as_double(0x7FF8000000000000LL);
                else lp = log(P) - static_cast<double>(scale_total) * 0.693147180559945309417;
                O.logp[ev] = lp;
            }
        }
    }
}

}  // namespace pp
