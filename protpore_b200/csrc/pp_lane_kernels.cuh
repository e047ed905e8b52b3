// pp_lane_kernels.cuh -- "lane-per-event" column kernels for sm_100a, run-based.
//
// One CTA = W warps x 32 lanes sweeps 32 events.  Lane j of every warp works on event j; the warps
// split the states of a column (pp_lowering.h).  Consequences:
//   * a warp's program is warp-uniform: no divergence, no per-lane CSR walk;
//   * column accesses are [slot][lane]: stride-1 across lanes, bank-conflict free;
//   * a lane only ever touches ITS event's column, so the serial silent chain of a column is plain
//     program order in one thread -- exact reference arithmetic, no scan;
//   * two __syncthreads per row: after the emitting phase (silent states read the new emitting
//     values) and after the silent phase (next row's emitting states read the silent values).
//
// The program of a warp is a short list of RUNS (pp_lowering.h): `count` items of one shape whose
// slots, weights and emission records advance by constant strides.  A run is executed by a body
// templated on the in-degree and on the number of items it handles at once (G = 2 or 4): all loads of
// the group are issued first, then G independent compare chains, then the stores -- with only 4 warps
// per scheduler (the float64 column of 32 events fills half an SM's shared memory) instruction-level
// parallelism inside the warp is what hides LDS and FP64 latency.  A silent chain (skip / back-slip)
// is a run whose last edge is forwarded in a register: per link the dependent sequence is
// DADD -> DSETP -> FSEL.  Because the bodies are shared by all runs of a shape, the code a CTA touches
// per row is a few KB however many states the model has (the first, fully unrolled specialisation of
// these kernels spent 41 % of its time waiting for instructions; profiles/).
//
// Shared memory per CTA: [column: n_slots x 32 doubles][transition weights][emission records]
// [observation tile][per-warp row maxima (forward)].  Run records live in a __constant__ arena.
//
// Viterbi (k_viterbi_lane): float64 max-plus sweep that follows Model._viterbi
// (yahmm.pyx:3471-3652) operation for operation -- candidate (v + t) + e for emitting states,
// v + t for silent ones, strict '>' in the reference's visiting order -- and writes the ordinal of
// the winning in-edge, packed 8 (or 4) to a 32-bit word, to HBM for the traceback kernel.
//
// Forward (k_forward_lane): Model._forward / _log_probability (yahmm.pyx:2810-2930, 3378-3396) in
// scaled LINEAR space, float64: one DFMA per edge instead of the reference's exp + log per edge.
// Each event carries a power-of-two scale that is re-derived every row from the previous row's
// maximum, so the mantissas are untouched by rescaling (DESIGN.md "forward").
#pragma once
#include "pp_device_common.cuh"
#include "pp_lane_types.h"

namespace pp {

static __constant__ double c_arena[kArenaDoubles];        // run records of the model whose image is resident

// Optional per-warp phase clocks (tools/phase_timing.py builds the library with -DPP_PHASE_TIMING): cycles a
// warp spends working in phase B (silent / early), waiting at B's barrier, working in phase A, waiting at A's.
#ifdef PP_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16 * 8];
#define PP_CLK(var) const long long var = clock64()
#define PP_ACC(slot, a, b) phase_acc[slot] += static_cast<unsigned long long>((b) - (a))
#else
#define PP_CLK(var)
#define PP_ACC(slot, a, b)
#endif

__device__ __forceinline__ const RunRec& run_rec(int i) { return reinterpret_cast<const RunRec*>(c_arena)[i]; }

// 32-bit shared-memory addressing: one register per pointer, LDS.64 / STS.64 with register + immediate
__device__ __forceinline__ double lds64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// Where the DP column of the CTA's 32 events lives.  CG = false: shared memory (the normal case; `a` is a 32-bit
// shared address).  CG = true: the column of a model too large for shared memory (n_slots x 256 B > ~220 KB) sits in a
// per-CTA global scratch that stays L2-resident; `a` is then a byte offset from the warp-uniform base `gcol`, so the
// load is LDG [R.U32 + UR.64] without address arithmetic.  Weights and emission records are in shared memory either way.
template <bool CG>
__device__ __forceinline__ double ldcol(uint32_t a, const unsigned char* gcol) {
    if (CG) return *reinterpret_cast<const double*>(gcol + a);
    return lds64(a);
}
template <bool CG>
__device__ __forceinline__ void stcol(uint32_t a, double v, unsigned char* gcol) {
    if (CG) *reinterpret_cast<double*>(gcol + a) = v;
    else sts64(a, v);
}

// induction state of one run: shared-memory addresses and byte strides
template <int NE>
struct RunPtrs {
    uint32_t sp[NE > 0 ? NE : 1];
    int ss[NE > 0 ? NE : 1];
    uint32_t dp, wp, ep;
    int ds, ws, es;
    int pbase;                       // byte offset of the next traceback unit inside the pointer row
    __device__ __forceinline__ void init(const RunRec& R, int par, uint32_t colb, uint32_t wtab, uint32_t etab,
                                         int erec_bytes) {
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            ss[j] = R.src_stride[j];
            sp[j] = colb + R.src_off[par][j];
        }
        ds = R.dst_stride;
        dp = colb + R.dst_off[par];
        ws = R.w_stride * 8;
        wp = wtab + R.w_base * 8;
        es = R.e_stride * erec_bytes;
        ep = etab + R.e_base * erec_bytes;
        pbase = R.pword;
    }
    __device__ __forceinline__ void advance(int g) {
#pragma unroll
        for (int j = 0; j < NE; ++j) sp[j] += g * ss[j];
        dp += g * ds;
        wp += g * ws;
        ep += g * es;
    }
};

constexpr int kVitEmisBytes = 40;    // Viterbi emission record {a0, a1, a2, a3, log(state.weight)}
constexpr int kFwdEmisBytes = 32;    // forward emission record {a0, a1 | k2, c2, cu}

// exact float64 emission log-density from the shared-memory record (see emission_log), + log(state.weight)
template <int KIND>
__device__ __forceinline__ double smem_emission_log(uint32_t ep, double x) {
    double r;
    if (KIND == EMIS_NORMAL) {
        const double mu = lds64(ep), b = lds64(ep + 8), rinv = lds64(ep + 16), c = lds64(ep + 24);
        const double d = __dsub_rn(x, mu);
        const double d2 = __dmul_rn(d, d);
        const double q0 = __dmul_rn(d2, rinv);
        const double rem = __fma_rn(-q0, b, d2);
        const double q = __fma_rn(rem, rinv, q0);
        r = __dsub_rn(c, q);
    } else if (KIND == EMIS_POINT) {
        r = (fabs(__dsub_rn(x, lds64(ep))) < 1E-4) ? 0.0 : neg_inf();
    } else {
        // a3 = log(1/(b-a)), or 0 when a == b (then "in range" means x == a == b, the reference's first branch)
        const double a = lds64(ep), b = lds64(ep + 8), c = lds64(ep + 24);
        r = (x >= a && x <= b) ? c : neg_inf();
    }
    return __dadd_rn(r, lds64(ep + 32));
}

// forward record {a0, a1 | k2, c2, cu}: log2-domain emission of the linear-space sweep
template <int KIND>
__device__ __forceinline__ double smem_emission_log2(uint32_t ep, double x) {
    if (KIND == EMIS_NORMAL) {
        const double d = x - lds64(ep);
        return fma(-(d * d), lds64(ep + 8), lds64(ep + 16));
    }
    if (KIND == EMIS_POINT) return (fabs(x - lds64(ep)) < 1E-4) ? lds64(ep + 16) : neg_inf();
    const double a = lds64(ep), b = lds64(ep + 8), c = lds64(ep + 16);     // c2 covers a == b as well (host)
    return (x >= a && x <= b) ? c : neg_inf();
}

enum : int { MODE_EMIT = 0, MODE_SILENT = 1, MODE_CHAIN = 2 };

// one coalesced store of the ordinals of a group: G items x BITS bits = 1, 2 or 4 bytes per lane
template <int G, int BITS>
__device__ __forceinline__ void store_ordinals(uint8_t* row, int& pbase, int lane, const uint32_t* arg, bool active) {
    constexpr int UNIT = (G * BITS / 8) > 0 ? (G * BITS / 8) : 1;
    uint32_t packed = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) packed |= arg[g] << (g * BITS);
    if (active) {
        uint8_t* p = row + pbase + lane * UNIT;
        if (UNIT == 1) *p = static_cast<uint8_t>(packed);
        else if (UNIT == 2) *reinterpret_cast<uint16_t*>(p) = static_cast<uint16_t>(packed);
        else *reinterpret_cast<uint32_t*>(p) = packed;
    }
    pbase += UNIT * kLanes;
}

struct VitCtx {
    uint8_t* row;        // this row's traceback pointer bytes (group base; lane offset applied per unit)
    int lane;
    bool active;
    double x;
    unsigned char* gcol; // CG kernels: base of the CTA's column in global memory
};

// ------------------------------------------------------------------------------------------
// Viterbi
// ------------------------------------------------------------------------------------------
// G consecutive items of a run.  MODE_CHAIN: the last edge of every item reads `prev`, the value of the
// item before it.  Candidate j of an item is (v_j + t_j) + e for emitting items, v_j + t_j for silent ones,
// exactly the reference's operations; the winner is the lowest index attaining the maximum, as the
// reference's strict '>' scan leaves it (a candidate of -inf never displaces index 0 there either).
template <int NE, int MODE, int G, int KIND, int BITS, bool CG>
__device__ __forceinline__ void vit_group(RunPtrs<NE>& P, const VitCtx& C, bool start_row0, double& prev) {
    constexpr int NL = (MODE == MODE_CHAIN) ? NE - 1 : NE;          // edges read from the column
    double s[G][NL > 0 ? NL : 1], wl[G], e[G], best[G];
    uint32_t arg[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < NL; ++j) s[g][j] = ldcol<CG>(P.sp[j] + g * P.ss[j], C.gcol);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < NL; ++j) s[g][j] = __dadd_rn(s[g][j], lds64(P.wp + g * P.ws + j * 8));
        if (MODE == MODE_CHAIN) wl[g] = lds64(P.wp + g * P.ws + (NE - 1) * 8);
    }
    if (MODE == MODE_EMIT) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            e[g] = smem_emission_log<KIND>(P.ep + g * P.es, C.x);
#pragma unroll
            for (int j = 0; j < NL; ++j) s[g][j] = __dadd_rn(s[g][j], e[g]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (NL > 0) {
            FirstMax<0, (NL > 0 ? NL : 1)>::run(s[g], best[g], arg[g]);
        } else {
            best[g] = neg_inf();
            arg[g] = 0;
        }
    }
    if (MODE == MODE_CHAIN) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const double t = __dadd_rn(prev, wl[g]);
            if (NE == 1 || t > best[g]) {
                best[g] = t;
                arg[g] = NE - 1;
            }
            prev = best[g];
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (start_row0) best[g] = 0.0;                // v[0, start] = 0 (3513-3514)
        stcol<CG>(P.dp + g * P.ds, best[g], C.gcol);
    }
    if (MODE != MODE_CHAIN) prev = best[G - 1];
    store_ordinals<G, BITS>(C.row, P.pbase, C.lane, arg, C.active);
    P.advance(G);
}

template <int NE, int MODE, int KIND, int BITS, bool CG>
__device__ __forceinline__ void vit_items(RunPtrs<NE>& P, int count, const VitCtx& C, bool start_row0, double& prev) {
    constexpr int G = run_group_size(NE);
    int left = count;
    for (; left >= G; left -= G) vit_group<NE, MODE, G, KIND, BITS, CG>(P, C, start_row0, prev);
    if (G > 2 && left >= 2) { vit_group<NE, MODE, 2, KIND, BITS, CG>(P, C, start_row0, prev); left -= 2; }
    if (left >= 1) vit_group<NE, MODE, 1, KIND, BITS, CG>(P, C, start_row0, prev);
}

template <int NE, int BITS, bool CG>
__device__ __forceinline__ void vit_run_ne(const RunRec& R, int par, bool row0, uint32_t colb, uint32_t wtab,
                                           uint32_t etab, const VitCtx& C) {
    RunPtrs<NE> P;
    P.init(R, par, colb, wtab, etab, kVitEmisBytes);
    double prev = 0.0;
    const int count = R.count;
    const int kind = R.kind;
    if (kind == EMIS_NORMAL) {
        vit_items<NE, MODE_EMIT, EMIS_NORMAL, BITS, CG>(P, count, C, false, prev);
    } else if (kind == EMIS_POINT) {
        vit_items<NE, MODE_EMIT, EMIS_POINT, BITS, CG>(P, count, C, false, prev);
    } else if (kind == EMIS_UNIFORM) {
        vit_items<NE, MODE_EMIT, EMIS_UNIFORM, BITS, CG>(P, count, C, false, prev);
    } else if (NE > 0 && R.fwd) {
        vit_group<NE, MODE_SILENT, 1, RUN_SILENT, BITS, CG>(P, C, false, prev);       // chain head reads the column
        vit_items<NE, MODE_CHAIN, RUN_SILENT, BITS, CG>(P, count - 1, C, false, prev);
    } else {
        vit_items<NE, MODE_SILENT, RUN_SILENT, BITS, CG>(P, count, C, row0 && kind == RUN_START, prev);
    }
}

template <int BITS, bool CG>
__device__ __forceinline__ void vit_run(int r, int par, bool row0, uint32_t colb, uint32_t wtab, uint32_t etab,
                                        const VitCtx& C) {
    const RunRec& R = run_rec(r);
    switch (R.n_edges) {
        case 0: vit_run_ne<0, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 1: vit_run_ne<1, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 2: vit_run_ne<2, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 3: vit_run_ne<3, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 4: vit_run_ne<4, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 5: vit_run_ne<5, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 6: vit_run_ne<6, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        case 7: vit_run_ne<7, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
        default: vit_run_ne<8, BITS, CG>(R, par, row0, colb, wtab, etab, C); break;
    }
}

// in-edges of the end state (any in-degree), evaluated once per event at its last row
__device__ __forceinline__ void vit_end_item(const EdgeRec* __restrict__ ep, int n_edges, const unsigned char* colp,
                                             double& best, int& arg) {
    best = neg_inf();
    arg = 0;
#pragma unroll 4
    for (int k = 0; k < n_edges; ++k) {
        const EdgeRec E = ldg_rec(ep + k);
        const double s = __dadd_rn(*reinterpret_cast<const double*>(colp + E.off), E.val);
        if (s > best) { best = s; arg = k; }
    }
}

// shared-memory carve-up shared by both kernels
struct LaneSmem {
    unsigned char* col;
    double* wtab;
    double* etab;
    unsigned char* xt;
    int32_t* wmax;
    __device__ __forceinline__ LaneSmem(unsigned char* smem, const LaneSched& S, bool col_in_smem = true) {
        col = smem;
        wtab = reinterpret_cast<double*>(smem + (col_in_smem ? static_cast<size_t>(S.n_slots) * kSlotBytes : 0));
        etab = wtab + S.n_low;
        xt = reinterpret_cast<unsigned char*>(etab + static_cast<size_t>(S.ne) * 5);
        wmax = reinterpret_cast<int32_t*>(xt + 32 * 33 * 4);
    }
};

template <int W>
__device__ __forceinline__ void fill_tile2(double* xt, const void* __restrict__ obs, bool f64, int64_t off, int n, int i0,
                                           int warp, int lane) {
    constexpr int TC = kLaneTileCols;                         // columns per tile; row stride 32 + 1 doubles
    for (int j = warp * (32 / TC) + lane / TC; j < kLanes; j += W * (32 / TC)) {
        const int64_t oj = __shfl_sync(0xffffffffu, off, j);
        const int nj = __shfl_sync(0xffffffffu, n, j);
        const int c = i0 + (lane % TC);
        double v = 0.0;
        if (c < nj) v = f64 ? __ldg(static_cast<const double*>(obs) + oj + c)
                            : static_cast<double>(__ldg(static_cast<const float*>(obs) + oj + c));
        xt[(lane % TC) * (kLanes + 1) + j] = v;
    }
}

template <int W, int BITS, bool CG = false>
__global__ void __launch_bounds__(W * 32, 2)
k_viterbi_lane(const LaneSched S, const LaneBatchAny B, const VitOut O) {
    extern __shared__ __align__(16) unsigned char smem[];
    const LaneSmem M(smem, S, !CG);
    constexpr int TC = kLaneTileCols;
    double* xt = reinterpret_cast<double*>(M.xt);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // tells the compiler it is warp-uniform
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    unsigned char* const gcol = CG ? S.gcol + static_cast<size_t>(blockIdx.x) * S.gcol_stride : nullptr;
    unsigned char* colp = (CG ? gcol : smem) + lane * 8;          // this lane's column slice
    const uint32_t colb = CG ? static_cast<uint32_t>(lane * 8) : static_cast<uint32_t>(__cvta_generic_to_shared(colp));
    const uint32_t wtab = static_cast<uint32_t>(__cvta_generic_to_shared(M.wtab));
    const uint32_t etab = static_cast<uint32_t>(__cvta_generic_to_shared(M.etab));
    const size_t row_stride = static_cast<size_t>(S.n_ptr_words) * kLanes * 4;                  // bytes per pointer row
    uint8_t* prow = reinterpret_cast<uint8_t*>(O.ptr) + static_cast<size_t>(B.group_row[g]) * row_stride;

    const int ab = S.run_ranges[warp * 6 + 0], ae = S.run_ranges[warp * 6 + 1];     // A: late emitting
    const int sb = S.run_ranges[warp * 6 + 2], se = S.run_ranges[warp * 6 + 3];     // B: silent
    const int xb = S.run_ranges[warp * 6 + 4], xe = S.run_ranges[warp * 6 + 5];     // B: early emitting of the next row

    // stage the model: weights and emission records; row 0: v[0, :] = -inf, v[0, start] = 0 (3513-3514)
    for (int i = threadIdx.x; i < S.n_low; i += W * 32) M.wtab[i] = __ldg(S.wtab + i);
    for (int i = threadIdx.x; i < S.ne * S.etab_doubles; i += W * 32) M.etab[i] = __ldg(S.etab + i);
    double* col = reinterpret_cast<double*>(CG ? gcol : smem);
    for (int s = threadIdx.x; s < S.n_slots * kLanes; s += W * 32) col[s] = neg_inf();
    __syncthreads();

#ifdef PP_PHASE_TIMING
    unsigned long long phase_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    double x = 0.0;                                          // observation of the row phase A works on
    for (int r = 0;; ++r) {
        // ---- phase B of row r: silent states of row r (3515-3542 for r = 0, 3564-3605 after) and the
        //      early emitting states of row r + 1 (3545-3562), which only read emitting values of row r
        const int q = r & 1;
        if (r < nmax && (r & (TC - 1)) == 0) {
            fill_tile2<W>(xt, B.obs, B.obs_f64 != 0, off, n, r, warp, lane);
            __syncthreads();
        }
        PP_CLK(c0);
        uint8_t* pr = prow + static_cast<size_t>(r) * row_stride;
        {
            const VitCtx C{pr, lane, r == 0 ? ev >= 0 : r <= n, 0.0, gcol};
            for (int u = sb; u < se; ++u) vit_run<BITS, CG>(u, q, r == 0, colb, wtab, etab, C);
        }
        PP_CLK(c1);
        if (r < nmax) {
            x = xt[(r & (TC - 1)) * (kLanes + 1) + lane];       // obs[r] belongs to row r + 1
            const VitCtx C{pr + row_stride, lane, r + 1 <= n, x, gcol};
            for (int u = xb; u < xe; ++u) vit_run<BITS, CG>(u, q ^ 1, false, colb, wtab, etab, C);
        }
        PP_CLK(c2);
        __syncthreads();
        PP_CLK(c3);
        PP_ACC(0, c0, c1); PP_ACC(1, c1, c2); PP_ACC(2, c2, c3);

        // events whose last row this is: final score (3614-3619).  Warp 0 does it while the others
        // already run phase A of the next row (they write E[q^1]; this reads E[q] and SIL).
        if (r >= 1 && warp == 0 && __any_sync(0xffffffffu, r == n)) {
            double score; int est = S.end_state, eord = 0;
            if (S.finite && S.end_final_only) {
                vit_end_item((q ? S.end_edges[1] : S.end_edges[0]) + S.end_edge_begin, S.end_n_edges, colp, score, eord);
            } else if (S.finite) {
                score = *reinterpret_cast<const double*>(colp + S.end_slot_off);
            } else {                                                 // numpy.argmax(v[n]) : first maximum
                score = *reinterpret_cast<const double*>(colp + static_cast<size_t>(S.ns + q * S.ne) * kSlotBytes);
                est = 0;
                for (int k = 1; k < S.m; ++k) {
                    const size_t slot = k < S.ne ? static_cast<size_t>(S.ns + q * S.ne + k) : static_cast<size_t>(k - S.ne);
                    const double v = *reinterpret_cast<const double*>(colp + slot * kSlotBytes);
                    if (v > score) { score = v; est = k; }
                }
            }
            if (r == n) {
                O.score[ev] = score;
                O.end_state[ev] = est;
                O.end_ord[ev] = eord;
            }
        }
        if (r == nmax) break;

        // ---- phase A of row r + 1: late emitting states (they have a silent source of row r)
        PP_CLK(c4);
        {
            const VitCtx C{pr + row_stride, lane, r + 1 <= n, x, gcol};
            for (int u = ab; u < ae; ++u) vit_run<BITS, CG>(u, q ^ 1, false, colb, wtab, etab, C);
        }
        PP_CLK(c5);
        __syncthreads();
        PP_CLK(c6);
        PP_ACC(3, c4, c5); PP_ACC(4, c5, c6); PP_ACC(5, c3, c4);
#ifdef PP_PHASE_TIMING
        phase_acc[6] += 1;
#endif
    }
#ifdef PP_PHASE_TIMING
    if (lane == 0)
        for (int k = 0; k < 8; ++k) atomicAdd(&g_phase_cycles[warp * 8 + k], phase_acc[k]);
#endif
}

// ------------------------------------------------------------------------------------------
// Forward (scaled linear space)
// ------------------------------------------------------------------------------------------
template <int NE, int MODE, int G, int KIND, bool CG>
__device__ __forceinline__ void fwd_group(RunPtrs<NE>& P, double x, double shiftd, bool start_row0, double& prev,
                                          int& mx, unsigned char* gcol) {
    constexpr int NL = (MODE == MODE_CHAIN) ? NE - 1 : NE;
    double v[G][NL > 0 ? NL : 1], w[G][NE > 0 ? NE : 1], acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < NL; ++j) v[g][j] = ldcol<CG>(P.sp[j] + g * P.ss[j], gcol);
#pragma unroll
        for (int j = 0; j < NE; ++j) w[g][j] = lds64(P.wp + g * P.ws + j * 8);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {                 // two interleaved accumulators halve the dependent DFMA chain
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < NL; j += 2) a0 = fma(v[g][j], w[g][j], a0);
#pragma unroll
        for (int j = 1; j < NL; j += 2) a1 = fma(v[g][j], w[g][j], a1);
        acc[g] = a0 + a1;
    }
    if (MODE == MODE_EMIT) {
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] *= exp2_scaled(smem_emission_log2<KIND>(P.ep + g * P.es, x) + shiftd);
    }
    if (MODE == MODE_CHAIN) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            acc[g] = fma(prev, w[g][NE - 1], acc[g]);
            prev = acc[g];
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (start_row0) acc[g] = 1.0;
        stcol<CG>(P.dp + g * P.ds, acc[g], gcol);
        mx = max(mx, __double2hiint(acc[g]));
    }
    if (MODE != MODE_CHAIN) prev = acc[G - 1];
    P.advance(G);
}

template <int NE, int MODE, int KIND, bool CG>
__device__ __forceinline__ void fwd_items(RunPtrs<NE>& P, int count, double x, double shiftd, bool start_row0,
                                          double& prev, int& mx, unsigned char* gcol) {
    constexpr int G = run_group_size(NE);
    int left = count;
    for (; left >= G; left -= G) fwd_group<NE, MODE, G, KIND, CG>(P, x, shiftd, start_row0, prev, mx, gcol);
    if (G > 2 && left >= 2) { fwd_group<NE, MODE, 2, KIND, CG>(P, x, shiftd, start_row0, prev, mx, gcol); left -= 2; }
    if (left >= 1) fwd_group<NE, MODE, 1, KIND, CG>(P, x, shiftd, start_row0, prev, mx, gcol);
}

template <int NE, bool CG>
__device__ __forceinline__ void fwd_run_ne(const RunRec& R, int par, bool row0, uint32_t colb, uint32_t wtab,
                                           uint32_t etab, double x, double shiftd, int& mx, unsigned char* gcol) {
    RunPtrs<NE> P;
    P.init(R, par, colb, wtab, etab, kFwdEmisBytes);
    double prev = 0.0;
    const int count = R.count;
    const int kind = R.kind;
    if (kind == EMIS_NORMAL) {
        fwd_items<NE, MODE_EMIT, EMIS_NORMAL, CG>(P, count, x, shiftd, false, prev, mx, gcol);
    } else if (kind == EMIS_POINT) {
        fwd_items<NE, MODE_EMIT, EMIS_POINT, CG>(P, count, x, shiftd, false, prev, mx, gcol);
    } else if (kind == EMIS_UNIFORM) {
        fwd_items<NE, MODE_EMIT, EMIS_UNIFORM, CG>(P, count, x, shiftd, false, prev, mx, gcol);
    } else if (NE > 0 && R.fwd) {
        fwd_group<NE, MODE_SILENT, 1, RUN_SILENT, CG>(P, x, shiftd, false, prev, mx, gcol);
        fwd_items<NE, MODE_CHAIN, RUN_SILENT, CG>(P, count - 1, x, shiftd, false, prev, mx, gcol);
    } else {
        fwd_items<NE, MODE_SILENT, RUN_SILENT, CG>(P, count, x, shiftd, row0 && kind == RUN_START, prev, mx, gcol);
    }
}

template <bool CG = false>
__device__ __forceinline__ void fwd_run(int r, int par, bool row0, uint32_t colb, uint32_t wtab, uint32_t etab,
                                        double x, double shiftd, int& mx, unsigned char* gcol = nullptr) {
    const RunRec& R = run_rec(r);
    switch (R.n_edges) {
        case 0: fwd_run_ne<0, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 1: fwd_run_ne<1, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 2: fwd_run_ne<2, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 3: fwd_run_ne<3, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 4: fwd_run_ne<4, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 5: fwd_run_ne<5, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 6: fwd_run_ne<6, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        case 7: fwd_run_ne<7, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
        default: fwd_run_ne<8, CG>(R, par, row0, colb, wtab, etab, x, shiftd, mx, gcol); break;
    }
}

template <int W, bool CG = false>
__global__ void __launch_bounds__(W * 32, 2)
k_forward_lane(const LaneSched S, const LaneBatchAny B, const FwdOut O) {
    extern __shared__ __align__(16) unsigned char smem[];
    const LaneSmem M(smem, S, !CG);
    constexpr int TC = kLaneTileCols;
    double* xt = reinterpret_cast<double*>(M.xt);
    int32_t* wmax = M.wmax;

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // tells the compiler it is warp-uniform
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    unsigned char* const gcol = CG ? S.gcol + static_cast<size_t>(blockIdx.x) * S.gcol_stride : nullptr;
    unsigned char* colp = (CG ? gcol : smem) + lane * 8;
    const uint32_t colb = CG ? static_cast<uint32_t>(lane * 8) : static_cast<uint32_t>(__cvta_generic_to_shared(colp));
    const uint32_t wtab = static_cast<uint32_t>(__cvta_generic_to_shared(M.wtab));
    const uint32_t etab = static_cast<uint32_t>(__cvta_generic_to_shared(M.etab));

    const int ab = S.run_ranges[warp * 6 + 0], ae = S.run_ranges[warp * 6 + 1];     // A: late emitting
    const int sb = S.run_ranges[warp * 6 + 2], se = S.run_ranges[warp * 6 + 3];     // B: silent
    const int xb = S.run_ranges[warp * 6 + 4], xe = S.run_ranges[warp * 6 + 5];     // B: early emitting of the next row

    for (int i = threadIdx.x; i < S.n_low; i += W * 32) M.wtab[i] = __ldg(S.wtab + i);
    for (int i = threadIdx.x; i < S.ne * S.etab_doubles; i += W * 32) M.etab[i] = __ldg(S.etab + i);
    double* col = reinterpret_cast<double*>(CG ? gcol : smem);
    for (int s = threadIdx.x; s < S.n_slots * kLanes; s += W * 32) col[s] = 0.0;
    wmax[warp * kLanes + lane] = 0;
    __syncthreads();

    // Scaling: the whole of row r + 1 is multiplied by 2^shift, shift chosen from the largest EMITTING value
    // of row r (silent values of a row derive from its emitting values by probabilities <= 1).  Each warp
    // tracks the maximum over the row-r values it produced (early ones in phase B of row r - 1, late ones in
    // phase A of row r) and publishes it before the barrier that ends phase A.
    long long scale_total = 0;                    // sum of the shifts of rows 1 .. r + 1
    int mx = 0;                                   // max high word over this warp's emitting values of the next row
    double x = 0.0, shiftd = 0.0;
    for (int r = 0;; ++r) {
        const int q = r & 1;
        const long long scale_row = scale_total;  // shifts applied up to row r (what the final score needs)
        if (r < nmax) {
            int Mx = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) Mx = max(Mx, wmax[w * kLanes + lane]);
            const int ex = (Mx >> 20) & 0x7ff;
            int shift = (ex == 0 || ex == 0x7ff) ? 0 : (kFwdTargetExp + 1023 - ex);
            shift = max(-900, min(900, shift));
            scale_total += shift;
            shiftd = static_cast<double>(shift);
            if ((r & (TC - 1)) == 0) {
                fill_tile2<W>(xt, B.obs, B.obs_f64 != 0, off, n, r, warp, lane);
                __syncthreads();
            }
        }
        int mxs = 0;                              // silent values do not steer the scale
        for (int u = sb; u < se; ++u) fwd_run<CG>(u, q, r == 0, colb, wtab, etab, 0.0, 0.0, mxs, gcol);    // silent (2847-2871, 2891-2926)
        mx = 0;
        if (r < nmax) {
            x = xt[(r & (TC - 1)) * (kLanes + 1) + lane];
            for (int u = xb; u < xe; ++u) fwd_run<CG>(u, q ^ 1, false, colb, wtab, etab, x, shiftd, mx, gcol);   // early emitting of row r + 1
        }
        __syncthreads();

        if (r >= 1 && warp == 0 && __any_sync(0xffffffffu, r == n)) {          // log_probability (3378-3396)
            double P;
            if (S.finite && S.end_final_only) {
                const EdgeRec* ep = (q ? S.end_edges[1] : S.end_edges[0]) + S.end_edge_begin;
                P = 0.0;
#pragma unroll 4
                for (int k = 0; k < S.end_n_edges; ++k) {
                    const EdgeRec E = ldg_rec(ep + k);
                    P = fma(*reinterpret_cast<const double*>(colp + E.off), E.val, P);
                }
            } else if (S.finite) {
                P = *reinterpret_cast<const double*>(colp + S.end_slot_off);
            } else {
                P = 0.0;
                for (int k = 0; k < S.ne; ++k)
                    P += *reinterpret_cast<const double*>(colp + static_cast<size_t>(S.ns + q * S.ne + k) * kSlotBytes);
            }
            if (r == n) {
                // P == 0 -> -inf (impossible, or every path underflowed: the host re-runs such events
                // with the exact log-space kernel); inf / NaN -> NaN, same treatment.
                double lp;
                if (P == 0.0) lp = neg_inf();
                else if (!(P < __longlong_as_double(0x7FF0000000000000LL))) lp = __longlong_as_double(0x7FF8000000000000LL);
                else lp = log(P) - static_cast<double>(scale_row) * 0.693147180559945309417;
                O.logp[ev] = lp;
            }
        }
        if (r == nmax) break;

        for (int u = ab; u < ae; ++u) fwd_run<CG>(u, q ^ 1, false, colb, wtab, etab, x, shiftd, mx, gcol);       // late emitting (2874-2889)
        wmax[warp * kLanes + lane] = mx;
        __syncthreads();
    }
}

}  // namespace pp
