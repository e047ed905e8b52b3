// pp_lane_fb.cuh -- lane-per-event forward-backward (Baum-Welch E-step) kernel for sm_100a.
//
// Model._forward_backward / _train_once_baum_welch (yahmm.pyx:3193-3364, 4107-4236) in sufficient-statistic
// form, for models the lane kernels serve (pp_lowering.h) with a final-only end state:
//
//   phase F  the scaled linear-space forward sweep of pp_lane_kernels.cuh (same runs, same power-of-two row
//            scaling) that also writes every column f^_r[k] = f_r[k] 2^{S_r} and S_r to a per-CTA HBM scratch
//            ([row][silent | emitting | scale][lane], 256-byte coalesced stores), and P^ = P 2^{S_n};
//   phase B  the backward sweep over the out-edge runs of pp_lowering_bw.cpp, rows n .. 0, with its own row
//            scaling b^_r = b_r 2^{T_r} (T_n = 0, b^_n[end] = 1).  For an item k with out-edges j the products
//            p_j = w_j * value_j that make up b^_r[k] are also the expected-count terms: with
//            g = f^_r[k] * c_r,  c_r = 2^{S_n - S_r - T_r} / P^   (so that g * b^_r[k] is the posterior of k at r)
//              count(k -> l_j)            += g * p_j                      (3246-3302)
//              (sum w, sum wx, sum wxx)_k += g * b^_r[k] * (1, x, x^2)    (4215-4236; emitting k, x = obs[r - 1])
//            The per-lane (= per-event) terms are summed over the 32 lanes with a shuffle TRANSPOSE-reduce: a
//            chunk of 2^s terms takes 2^s - 1 exchange steps and leaves lane L with the partial sum of term
//            L >> (5 - s); that one double is added to the chunk's private accumulation slot (per CTA, in L2).
//            A finalize kernel folds the slots of all CTAs into the caller's statistics.
//
// Emissions use a float64-accurate 2^t (table of 2^(j/32) + degree-6 polynomial) instead of MUFU.EX2, so the
// statistics agree with the reference to ~1e-12 relative.  An event whose scaled sweep cannot be trusted
// (P^ == 0 / inf / NaN) contributes exactly nothing here and is handed to the exact log-space kernel by the
// host; a scale that leaves the double range, or a final check  f^_0[start] b^_0[start] c_0 != 1, raises the
// batch's fail counter and the host redoes the whole batch with the exact kernel.
#pragma once
#include "pp_lane_kernels.cuh"

namespace pp {

constexpr int kFbTargetExp = 480;        // row maxima of both sweeps are steered to ~2^480 (c_r ~ 2^-960 stays normal)
constexpr int kFbTileCols = 8;
constexpr int kT2Entries = 17;           // 2^(j/16), j = -8 .. 8 (136 bytes: bank-conflict free but for -8 / 8)


__host__ __device__ inline size_t fb_smem_bytes(int n_slots, int n_w, int ne, int warps) {
    return static_cast<size_t>(n_slots) * kSlotBytes + static_cast<size_t>(n_w) * 8 + static_cast<size_t>(ne) * 32 +
           static_cast<size_t>(kFbTileCols) * 33 * 8 + static_cast<size_t>(warps) * kLanes * 4 + kLanes * 16 +
           static_cast<size_t>(ne) * 4 + kT2Entries * 8 + 64;
}

__device__ __forceinline__ double ldcg64(const void* p) { return __ldcg(static_cast<const double*>(p)); }
__device__ __forceinline__ void stcg64(void* p, double v) { __stcg(static_cast<double*>(p), v); }

// 2^t for t in the log2 domain, float64-accurate (rel. error ~6e-16): t = n + j/16 + g, |g| <= 1/32;
// 2^g by a degree-6 Taylor polynomial of exp(g ln 2), 2^(j/32) from a shared-memory table, 2^n into the exponent.
// Below 2^-1000 the factor is an exact 0.
__device__ __forceinline__ double exp2_acc(double t, uint32_t t2tab) {
    // no clamp: t <= log2(max density) + 900 < 1000 by construction; for t <= -1000 (or -inf / NaN) the intermediate
    // values are meaningless but never used -- the final select returns an exact 0
    const double tc = t;
    const double magic = 6755399441055744.0;
    const double tm = tc + magic;
    const int n = __double2loint(tm);
    const double fr = tc - (tm - magic);                       // [-0.5, 0.5]
    const double um = fma(fr, 16.0, magic);
    const int j = __double2loint(um);                           // [-8, 8]
    const double g = fma(-(um - magic), 0.0625, fr);            // [-1/32, 1/32]
    const double z = g * 0.693147180559945309417;
    double p = fma(z, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(p, z, 1.0 / 24.0);
    p = fma(p, z, 1.0 / 6.0);
    p = fma(p, z, 0.5);
    p = fma(p, z, 1.0);
    p = fma(p, z, 1.0);
    const double tv = lds64(t2tab + static_cast<uint32_t>(j + 8) * 8u);
    const double scale = __hiloint2double((n + 1023) << 20, 0);
    const double r = (p * tv) * scale;
    return (t > -1000.0) ? r : 0.0;               // false for NaN as well
}

// membership test of the point-mass / uniform densities (the value is the run's shared 2^(c2 + shift))
template <int KIND>
__device__ __forceinline__ bool emis_in_range(uint32_t ep, double x) {
    if (KIND == EMIS_POINT) return fabs(x - lds64(ep)) < 1E-4;
    return x >= lds64(ep) && x <= lds64(ep + 8);
}

template <int W>
__device__ __forceinline__ void fill_tile_fb(double* xt, const void* __restrict__ obs, bool f64, int64_t off, int n,
                                             int i0, int warp, int lane) {
    constexpr int TC = kFbTileCols;
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

// ------------------------------------------------------------------------------------------
// phase F: forward groups that also store the column to HBM
// ------------------------------------------------------------------------------------------
struct FwdCtx {
    double x, shiftd;
    uint32_t t2tab;
};

template <int NE, int MODE, int G, int KIND, bool SH = false>
__device__ __forceinline__ void fbf_group(RunPtrs<NE>& P, unsigned char*& gp, int gs, const FwdCtx& C, bool start_row0,
                                          double& prev, int& mx, double E = 0.0) {
    constexpr int NL = (MODE == MODE_CHAIN) ? NE - 1 : NE;
    double v[G][NL > 0 ? NL : 1], w[G][NE > 0 ? NE : 1], acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < NL; ++j) v[g][j] = lds64(P.sp[j] + g * P.ss[j]);
#pragma unroll
        for (int j = 0; j < NE; ++j) w[g][j] = lds64(P.wp + g * P.ws + j * 8);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < NL; j += 2) a0 = fma(v[g][j], w[g][j], a0);
#pragma unroll
        for (int j = 1; j < NL; j += 2) a1 = fma(v[g][j], w[g][j], a1);
        acc[g] = a0 + a1;
    }
    if (MODE == MODE_EMIT) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (SH) acc[g] = emis_in_range<KIND>(P.ep + g * P.es, C.x) ? acc[g] * E : 0.0;
            else acc[g] *= exp2_acc(smem_emission_log2<KIND>(P.ep + g * P.es, C.x) + C.shiftd, C.t2tab);
        }
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
        sts64(P.dp + g * P.ds, acc[g]);
        stcg64(gp + g * gs, acc[g]);
        mx = max(mx, __double2hiint(acc[g]));
    }
    if (MODE != MODE_CHAIN) prev = acc[G - 1];
    gp += G * gs;
    P.advance(G);
}

template <int NE, int MODE, int KIND, bool SH = false>
__device__ __forceinline__ void fbf_items(RunPtrs<NE>& P, unsigned char*& gp, int gs, int count, const FwdCtx& C,
                                          bool start_row0, double& prev, int& mx, double E = 0.0) {
    constexpr int G = run_group_size(NE);
    int left = count;
    for (; left >= G; left -= G) fbf_group<NE, MODE, G, KIND, SH>(P, gp, gs, C, start_row0, prev, mx, E);
    if (G > 2 && left >= 2) { fbf_group<NE, MODE, 2, KIND, SH>(P, gp, gs, C, start_row0, prev, mx, E); left -= 2; }
    if (left >= 1) fbf_group<NE, MODE, 1, KIND, SH>(P, gp, gs, C, start_row0, prev, mx, E);
}

template <int NE>
__device__ __forceinline__ void fbf_run_ne(const RunRec& R, int par, bool row0, uint32_t colb, uint32_t wtab,
                                           uint32_t etab, unsigned char* grow, const FwdCtx& C, int& mx) {
    RunPtrs<NE> P;
    P.init(R, par, colb, wtab, etab, kFwdEmisBytes);
    unsigned char* gp = grow + R.dst_off[0];                  // scratch rows use the parity-0 column layout
    const int gs = R.dst_stride;
    double prev = 0.0;
    const int count = R.count;
    const int kind = R.kind;
    if (kind == EMIS_NORMAL) {
        fbf_items<NE, MODE_EMIT, EMIS_NORMAL>(P, gp, gs, count, C, false, prev, mx);
    } else if (kind == EMIS_POINT || kind == EMIS_UNIFORM) {
        if (R.shared_c2) {                                    // one exponentiation per run and row
            const double E = exp2_acc(lds64(P.ep + 16) + C.shiftd, C.t2tab);
            if (kind == EMIS_POINT) fbf_items<NE, MODE_EMIT, EMIS_POINT, true>(P, gp, gs, count, C, false, prev, mx, E);
            else fbf_items<NE, MODE_EMIT, EMIS_UNIFORM, true>(P, gp, gs, count, C, false, prev, mx, E);
        } else if (kind == EMIS_POINT) {
            fbf_items<NE, MODE_EMIT, EMIS_POINT>(P, gp, gs, count, C, false, prev, mx);
        } else {
            fbf_items<NE, MODE_EMIT, EMIS_UNIFORM>(P, gp, gs, count, C, false, prev, mx);
        }
    } else if (NE > 0 && R.fwd) {
        fbf_group<NE, MODE_SILENT, 1, RUN_SILENT>(P, gp, gs, C, false, prev, mx);
        fbf_items<NE, MODE_CHAIN, RUN_SILENT>(P, gp, gs, count - 1, C, false, prev, mx);
    } else {
        fbf_items<NE, MODE_SILENT, RUN_SILENT>(P, gp, gs, count, C, row0 && kind == RUN_START, prev, mx);
    }
}

__device__ __forceinline__ void fbf_run(int r, int par, bool row0, uint32_t colb, uint32_t wtab, uint32_t etab,
                                        unsigned char* grow, const FwdCtx& C, int& mx) {
    const RunRec& R = run_rec(r);
    switch (R.n_edges) {
        case 0: fbf_run_ne<0>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 1: fbf_run_ne<1>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 2: fbf_run_ne<2>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 3: fbf_run_ne<3>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 4: fbf_run_ne<4>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 5: fbf_run_ne<5>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 6: fbf_run_ne<6>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        case 7: fbf_run_ne<7>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
        default: fbf_run_ne<8>(R, par, row0, colb, wtab, etab, grow, C, mx); break;
    }
}

// ------------------------------------------------------------------------------------------
// phase B: backward groups with statistic accumulation
// ------------------------------------------------------------------------------------------
// Sum over the 32 lanes of 2^S per-lane terms: step j pairs lane bit 4-j with term-index bit S-1-j; afterwards lane
// L holds the partial sum (over the 2^S lanes that differ from L in their top S bits) of term L >> (5 - S).
template <int S>
__device__ __forceinline__ double lane_transpose_sum(double* v, int lane) {
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int half = (1 << S) >> (j + 1);
        const bool hi_side = (lane >> (4 - j)) & 1;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double lo = v[i], hi = v[i + half];
            const double keep = hi_side ? hi : lo;
            const double send = hi_side ? lo : hi;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16 >> j);
        }
    }
    return v[0];
}

// number of accumulation slots N terms take: chunks of kAccChunk, then 2^(kAccLog2-1) .. 1 (pp_lowering_bw.cpp replays this)
constexpr int acc_slots_for(int n) {
    int slots = 0;
    for (int s = kAccLog2; s >= 0; --s)
        while (n >= (1 << s)) { ++slots; n -= 1 << s; }
    return slots;
}

struct AccCtx {
    double* p;           // this lane's double in the CTA's next accumulation slot
    int lane;
};

template <int N, int S, int CI>
struct AccChunks {
    // consumes t[0 .. N) in chunks of 2^S, then recurses with S - 1; CI = index of the next prefetched slot value
    template <int NS>
    __device__ static __forceinline__ void run(double* t, const double (&a)[NS], AccCtx& A) {
        if constexpr (N >= (1 << S)) {
            const double v = lane_transpose_sum<S>(t, A.lane);
            stcg64(A.p + CI * kLanes, a[CI] + v);
            AccChunks<N - (1 << S), S, CI + 1>::run(t + (1 << S), a, A);
        } else if constexpr (S > 0) {
            AccChunks<N, S - 1, CI>::run(t, a, A);
        }
    }
};

struct BwdCtx {
    double c;            // c_r of this lane (0 when the lane is inactive at this row)
    double xb;           // obs[r - 1]: the observation the posteriors of row r weight
    uint32_t flags;      // shared-memory address of the per-state "some lane gave it weight" ballots
};

// One group of G items of a backward run.
//   STORE: the items' values b^_r[k] = sum_j w_j * value_j are computed and written to the column
//          (MODE_CHAIN: the last edge reads the previous item's value from a register);
//   ACC:   the per-lane terms g * w_j * value_j (and g * b^_r[k] * (1, x, x^2) for EMIT items) go through the
//          transpose-reduce into the warp's accumulation slots.
// List 1 (silent values) is <STORE, !ACC>; list 2 is <STORE, ACC, EMIT> for emitting states and <!STORE, ACC> for the
// accumulate-only items of the silent states, which re-read the finished column.
template <int NE, int MODE, int G, bool EMIT, bool ACC, bool STORE>
__device__ __forceinline__ void fbb_group(RunPtrs<NE>& P, const unsigned char*& fp, int gs, int& st, int sts,
                                          AccCtx& A, const BwdCtx& C, double& prev, int& mx) {
    constexpr int NL = (MODE == MODE_CHAIN) ? NE - 1 : NE;
    constexpr int NT1 = NE + (EMIT ? 3 : 0);
    constexpr int NT = ACC ? G * NT1 : 0;
    constexpr int NS = acc_slots_for(NT);
    double a[NS > 0 ? NS : 1], fh[G];
    if (ACC) {
#pragma unroll
        for (int i = 0; i < NS; ++i) a[i] = ldcg64(A.p + i * kLanes);       // issued first: L2 latency hides behind the group
#pragma unroll
        for (int g = 0; g < G; ++g) fh[g] = ldcg64(fp + g * gs);
    }
    double v[G][NL > 0 ? NL : 1], w[G][NE > 0 ? NE : 1], acc[G];
    double t[NT > 0 ? NT : 1];
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < NL; ++j) v[g][j] = lds64(P.sp[j] + g * P.ss[j]);
#pragma unroll
        for (int j = 0; j < NE; ++j) w[g][j] = lds64(P.wp + g * P.ws + j * 8);
    }
    if (ACC) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                const double p = v[g][j] * w[g][j];
                t[g * NT1 + j] = p;
                if (STORE) { if (j & 1) a1 += p; else a0 += p; }
            }
            acc[g] = a0 + a1;
        }
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < NL; j += 2) a0 = fma(v[g][j], w[g][j], a0);
#pragma unroll
            for (int j = 1; j < NL; j += 2) a1 = fma(v[g][j], w[g][j], a1);
            acc[g] = a0 + a1;
        }
    }
    if (MODE == MODE_CHAIN) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            acc[g] = fma(prev, w[g][NE - 1], acc[g]);
            prev = acc[g];
        }
    }
    if (STORE) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            sts64(P.dp + g * P.ds, acc[g]);
            if (EMIT) mx = max(mx, __double2hiint(acc[g]));
        }
        if (MODE != MODE_CHAIN) prev = acc[G - 1];
    }
    if (ACC) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            // gamma-like terms are products of three factors with independent power-of-two scales: f^ (around 2^480 times
            // the state's relative forward weight), the backward-scaled term (likewise) and c_r (around 2^-960).  f^ c_r
            // can leave the double range although the full product does not (a state with a forward weight of 1e-200
            // and an ordinary backward weight): where that happens for any lane the two large factors are multiplied
            // first -- they cannot overflow (<= 2^962) -- so a term is lost only if it is not a double at all.  That is the
            // reference's own limit: exp(f + b - logZ) != 0 (yahmm.pyx:304-326, 4215-4236).
            const double gk = fh[g] * C.c;
            const bool careful = __any_sync(0xffffffffu, fabs(gk) < 2.3e-308 && fh[g] != 0.0 && C.c != 0.0);
            if (!careful) {
#pragma unroll
                for (int j = 0; j < NE; ++j) t[g * NT1 + j] *= gk;
            } else {
#pragma unroll
                for (int j = 0; j < NE; ++j) t[g * NT1 + j] = (t[g * NT1 + j] * fh[g]) * C.c;
            }
            if (EMIT) {
                const double gam = careful ? (fh[g] * acc[g]) * C.c : gk * acc[g];
                const double gx = gam * C.xb;
                t[g * NT1 + NE] = gam;
                t[g * NT1 + NE + 1] = gx;
                t[g * NT1 + NE + 2] = gx * C.xb;
                const bool nz = gam > 0.0;                 // "this event gave the state weight"
                const uint32_t b = __ballot_sync(0xffffffffu, nz);
                if (A.lane == 0 && b != 0u) {
                    const uint32_t fa = C.flags + static_cast<uint32_t>(st + g * sts) * 4u;
                    uint32_t old;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old) : "r"(fa) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(fa), "r"(old | b) : "memory");
                }
            }
        }
        if (NT > 0) AccChunks<NT, kAccLog2, 0>::run(t, a, A);
        A.p += NS * kLanes;
        fp += G * gs;
        st += G * sts;
    }
    P.advance(G);
}

template <int NE, int MODE, bool EMIT, bool ACC, bool STORE>
__device__ __forceinline__ void fbb_items(RunPtrs<NE>& P, const unsigned char*& fp, int gs, int& st, int sts, int count,
                                          AccCtx& A, const BwdCtx& C, double& prev, int& mx) {
    constexpr int G = run_group_size(NE);
    int left = count;
    for (; left >= G; left -= G) fbb_group<NE, MODE, G, EMIT, ACC, STORE>(P, fp, gs, st, sts, A, C, prev, mx);
    if (G > 2 && left >= 2) { fbb_group<NE, MODE, 2, EMIT, ACC, STORE>(P, fp, gs, st, sts, A, C, prev, mx); left -= 2; }
    if (left >= 1) fbb_group<NE, MODE, 1, EMIT, ACC, STORE>(P, fp, gs, st, sts, A, C, prev, mx);
}

// LIST 1: silent values; LIST 2: accumulating items (emitting: value + statistics; RUN_ACC: statistics only)
template <int NE, int LIST>
__device__ __forceinline__ void fbb_run_ne(const RunRec& R, uint32_t colb, uint32_t wtab, const unsigned char* grow,
                                           AccCtx& A, const BwdCtx& C, int& mx) {
    RunPtrs<NE> P;
    P.init(R, 0, colb, wtab, 0u, 0);
    const unsigned char* fp = grow + R.dst_off[0];            // f^_r of the item's own state: same slot layout
    const int gs = R.dst_stride;
    int st = R.e_base;
    const int sts = R.e_stride;
    double prev = 0.0;
    if (LIST == 1) {
        if (NE > 0 && R.fwd) {
            fbb_group<NE, MODE_SILENT, 1, false, false, true>(P, fp, gs, st, sts, A, C, prev, mx);   // chain head reads the column
            fbb_items<NE, MODE_CHAIN, false, false, true>(P, fp, gs, st, sts, R.count - 1, A, C, prev, mx);
        } else {
            fbb_items<NE, MODE_SILENT, false, false, true>(P, fp, gs, st, sts, R.count, A, C, prev, mx);
        }
    } else if (R.kind == RUN_ACC) {
        fbb_items<NE, MODE_SILENT, false, true, false>(P, fp, gs, st, sts, R.count, A, C, prev, mx);
    } else {
        fbb_items<NE, MODE_SILENT, true, true, true>(P, fp, gs, st, sts, R.count, A, C, prev, mx);
    }
}

template <int LIST>
__device__ __forceinline__ void fbb_run(int r, uint32_t colb, uint32_t wtab, const unsigned char* grow, AccCtx& A,
                                        const BwdCtx& C, int& mx) {
    const RunRec& R = run_rec(r);
    switch (R.n_edges) {
        case 0: fbb_run_ne<0, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 1: fbb_run_ne<1, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 2: fbb_run_ne<2, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 3: fbb_run_ne<3, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 4: fbb_run_ne<4, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 5: fbb_run_ne<5, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 6: fbb_run_ne<6, LIST>(R, colb, wtab, grow, A, C, mx); break;
        case 7: fbb_run_ne<7, LIST>(R, colb, wtab, grow, A, C, mx); break;
        default: fbb_run_ne<8, LIST>(R, colb, wtab, grow, A, C, mx); break;
    }
}

// H[l] = p_l(x_r) 2^shift * b^_{r+1}[l] for the emitting states of a run (list 0 of pp_lowering_bw.cpp)
template <int KIND, bool SH = false>
__device__ __forceinline__ void fbh_items(const RunRec& R, uint32_t colb, uint32_t etab, double x, double shiftd,
                                          uint32_t t2tab) {
    uint32_t sp = colb + R.src_off[0][0], dp = colb + R.dst_off[0];
    uint32_t ep = etab + R.e_base * kFwdEmisBytes;
    const int ss = R.src_stride[0], ds = R.dst_stride, es = R.e_stride * kFwdEmisBytes;
    int left = R.count;
    double E = 0.0;
    if (SH) E = exp2_acc(lds64(ep + 16) + shiftd, t2tab);
    for (; left >= 4; left -= 4) {
        double b[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) b[g] = lds64(sp + g * ss);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (SH) b[g] = emis_in_range<KIND>(ep + g * es, x) ? b[g] * E : 0.0;
            else b[g] *= exp2_acc(smem_emission_log2<KIND>(ep + g * es, x) + shiftd, t2tab);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) sts64(dp + g * ds, b[g]);
        sp += 4 * ss; dp += 4 * ds; ep += 4 * es;
    }
    for (; left >= 1; --left) {
        double b = lds64(sp);
        if (SH) b = emis_in_range<KIND>(ep, x) ? b * E : 0.0;
        else b *= exp2_acc(smem_emission_log2<KIND>(ep, x) + shiftd, t2tab);
        sts64(dp, b);
        sp += ss; dp += ds; ep += es;
    }
}

__device__ __forceinline__ void fbh_run(int r, uint32_t colb, uint32_t etab, double x, double shiftd, uint32_t t2tab) {
    const RunRec& R = run_rec(r);
    if (R.kind == EMIS_NORMAL) fbh_items<EMIS_NORMAL>(R, colb, etab, x, shiftd, t2tab);
    else if (R.shared_c2 && R.kind == EMIS_POINT) fbh_items<EMIS_POINT, true>(R, colb, etab, x, shiftd, t2tab);
    else if (R.shared_c2) fbh_items<EMIS_UNIFORM, true>(R, colb, etab, x, shiftd, t2tab);
    else if (R.kind == EMIS_POINT) fbh_items<EMIS_POINT>(R, colb, etab, x, shiftd, t2tab);
    else fbh_items<EMIS_UNIFORM>(R, colb, etab, x, shiftd, t2tab);
}

__device__ __forceinline__ double load_obs(const LaneBatchAny& B, int64_t i) {
    return B.obs_f64 ? __ldg(static_cast<const double*>(B.obs) + i)
                     : static_cast<double>(__ldg(static_cast<const float*>(B.obs) + i));
}

__device__ __forceinline__ int shift_from_hiword(int Mx, int target) {
    const int ex = (Mx >> 20) & 0x7ff;
    const int shift = (ex == 0 || ex == 0x7ff) ? 0 : (target + 1023 - ex);
    return max(-900, min(900, shift));
}

__device__ __forceinline__ void atomic_min_f64_g(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(static_cast<long long>(old)) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f64_g(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(static_cast<long long>(old)) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}

template <int W>
__global__ void __launch_bounds__(W * 32, 2)
k_fb_lane(const LaneSched S, const FbSched F, const LaneBatchAny B) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TC = kFbTileCols;
    // carve-up: column | weights | emission records | observation tile | row maxima | (1/P^, S_n) | ballots | 2^(j/32)
    const int n_w = max(S.n_low, F.n_w);
    double* s_wtab = reinterpret_cast<double*>(smem + static_cast<size_t>(S.n_slots) * kSlotBytes);
    double* s_etab = s_wtab + n_w;
    double* xt = s_etab + static_cast<size_t>(S.ne) * 4;
    int32_t* wmax = reinterpret_cast<int32_t*>(xt + TC * 33);
    double* s_invp = reinterpret_cast<double*>(wmax + W * kLanes);
    double* s_sn = s_invp + kLanes;
    uint32_t* s_flags = reinterpret_cast<uint32_t*>(s_sn + kLanes);
    double* s_t2 = reinterpret_cast<double*>(s_flags + S.ne + (S.ne & 1));
    __shared__ int s_group;

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    unsigned char* colp = smem + lane * 8;
    const uint32_t colb = static_cast<uint32_t>(__cvta_generic_to_shared(colp));
    const uint32_t wtab = static_cast<uint32_t>(__cvta_generic_to_shared(s_wtab));
    const uint32_t etab = static_cast<uint32_t>(__cvta_generic_to_shared(s_etab));
    const uint32_t t2tab = static_cast<uint32_t>(__cvta_generic_to_shared(s_t2));
    const uint32_t flagsa = static_cast<uint32_t>(__cvta_generic_to_shared(s_flags));
    double* col = reinterpret_cast<double*>(smem);
    const size_t row_bytes = static_cast<size_t>(F.row_slots) * kSlotBytes;
    unsigned char* scratch = F.scratch + static_cast<size_t>(blockIdx.x) * F.scratch_stride + lane * 8;
    double* acc_cta = F.acc + static_cast<size_t>(blockIdx.x) * F.n_acc_slots * kLanes + lane;
    const uint32_t scale_off = static_cast<uint32_t>(S.ns + S.ne) * kSlotBytes;

    const int ab = S.run_ranges[warp * 6 + 0], ae = S.run_ranges[warp * 6 + 1];     // forward: late emitting
    const int sb = S.run_ranges[warp * 6 + 2], se = S.run_ranges[warp * 6 + 3];     //          silent
    const int xb_ = S.run_ranges[warp * 6 + 4], xe = S.run_ranges[warp * 6 + 5];    //          early emitting
    const int hb = F.run_ranges[warp * 6 + 0], he = F.run_ranges[warp * 6 + 1];     // backward: H items
    const int tb = F.run_ranges[warp * 6 + 2], te = F.run_ranges[warp * 6 + 3];     //           silent
    const int eb = F.run_ranges[warp * 6 + 4], ee = F.run_ranges[warp * 6 + 5];     //           emitting

    for (int i = threadIdx.x; i < S.ne * 4; i += W * 32) s_etab[i] = __ldg(S.etab + i);
    for (int i = threadIdx.x; i < kT2Entries; i += W * 32) s_t2[i] = exp2(static_cast<double>(i - 8) * 0.0625);

    for (;;) {
        __syncthreads();                                         // previous group is done with shared memory
        if (threadIdx.x == 0) s_group = atomicAdd(F.counter, 1);
        __syncthreads();
        if (s_group >= F.n_groups) break;
        // ticket -> group through a stride coprime to the number of groups (1 = sorted by length, the default; see
        // fb_batch_lane for the measurement behind it)
        const int g = static_cast<int>((static_cast<long long>(s_group) * F.group_stride) % F.n_groups);
        const int ev = B.order[g * kLanes + lane];
        const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
        const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
        const int nmax = B.group_len[g];

        // ================= phase F: forward sweep, columns to scratch =================
        for (int i = threadIdx.x; i < S.n_low; i += W * 32) s_wtab[i] = __ldg(S.wtab + i);
        for (int s = threadIdx.x; s < S.n_slots * kLanes; s += W * 32) col[s] = 0.0;
        for (int i = threadIdx.x; i < S.ne; i += W * 32) s_flags[i] = 0u;
        wmax[warp * kLanes + lane] = 0;
        if (warp == 0) { s_invp[lane] = 0.0; s_sn[lane] = 0.0; }
        for (int k = warp; k < S.ne; k += W)                     // f^_0[emitting] = 0: the sweep never writes row 0's emitting slots
            stcg64(scratch + static_cast<size_t>(S.ns + k) * kSlotBytes, 0.0);
        __syncthreads();

        double xlo = __longlong_as_double(0x7FF0000000000000LL), xhi = neg_inf();
        long long scale_total = 0;
        int mx = 0;
        FwdCtx FC{0.0, 0.0, t2tab};
        for (int r = 0;; ++r) {
            const int q = r & 1;
            const long long scale_row = scale_total;             // S_r
            if (r < nmax) {
                int Mx = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) Mx = max(Mx, wmax[w * kLanes + lane]);
                const int shift = shift_from_hiword(Mx, kFbTargetExp);
                scale_total += shift;
                FC.shiftd = static_cast<double>(shift);
                if ((r & (TC - 1)) == 0) {
                    fill_tile_fb<W>(xt, B.obs, B.obs_f64 != 0, off, n, r, warp, lane);
                    __syncthreads();
                }
            }
            unsigned char* grow = scratch + static_cast<size_t>(r) * row_bytes;
            int mxs = 0;
            {
                const FwdCtx C0{0.0, 0.0, t2tab};
                for (int u = sb; u < se; ++u) fbf_run(u, q, r == 0, colb, wtab, etab, grow, C0, mxs);
            }
            if (warp == 0) stcg64(grow + scale_off, static_cast<double>(scale_row));
            mx = 0;
            if (r < nmax) {
                FC.x = xt[(r & (TC - 1)) * (kLanes + 1) + lane];
                if (r < n) { xlo = fmin(xlo, FC.x); xhi = fmax(xhi, FC.x); }
                for (int u = xb_; u < xe; ++u) fbf_run(u, q ^ 1, false, colb, wtab, etab, grow + row_bytes, FC, mx);
            }
            __syncthreads();

            if (r >= 1 && warp == 0 && __any_sync(0xffffffffu, r == n)) {          // P^ = sum over the end state's in-edges
                const EdgeRec* ep = (q ? S.end_edges[1] : S.end_edges[0]) + S.end_edge_begin;
                double P = 0.0;
#pragma unroll 4
                for (int k = 0; k < S.end_n_edges; ++k) {
                    const EdgeRec E = ldg_rec(ep + k);
                    P = fma(*reinterpret_cast<const double*>(colp + E.off), E.val, P);
                }
                if (r == n) {
                    double lp;
                    const bool finite = P < __longlong_as_double(0x7FF0000000000000LL);
                    if (P == 0.0) lp = neg_inf();
                    else if (!finite) lp = __longlong_as_double(0x7FF8000000000000LL);
                    else lp = log(P) - static_cast<double>(scale_row) * 0.693147180559945309417;
                    F.logp[ev] = lp;
                    s_invp[lane] = (P > 0.0 && finite) ? 1.0 / P : 0.0;
                    s_sn[lane] = static_cast<double>(scale_row);
                }
            }
            if (r == nmax) break;
            for (int u = ab; u < ae; ++u) fbf_run(u, q ^ 1, false, colb, wtab, etab, grow + row_bytes, FC, mx);
            wmax[warp * kLanes + lane] = mx;
            __syncthreads();
        }
        __syncthreads();

        // ================= phase B: backward sweep with accumulation =================
        for (int i = threadIdx.x; i < F.n_w; i += W * 32) s_wtab[i] = __ldg(F.wtab + i);
        for (int s = threadIdx.x; s < S.n_slots * kLanes; s += W * 32) col[s] = 0.0;
        wmax[warp * kLanes + lane] = 0;
        __syncthreads();
        const double invp = s_invp[lane], sn = s_sn[lane];
        const bool valid = ev >= 0 && invp > 0.0;
        bool failed = false;
        double T = 0.0;                                          // T_r: sum of the backward shifts of rows n-1 .. r
        double c = 0.0;
        // xa = obs[r] (H of row r), xb = obs[r - 1] (posteriors of row r); one load per row, a row ahead
        double xa = 0.0;
        double xbv = (nmax - 1 < n && nmax >= 1) ? load_obs(B, off + nmax - 1) : 0.0;
        for (int r = nmax; r >= 0; --r) {
            double xpre = 0.0;
            if (r - 2 >= 0 && r - 2 < n) xpre = load_obs(B, off + r - 2);
            int Mx = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) Mx = max(Mx, wmax[w * kLanes + lane]);
            const int shift = shift_from_hiword(Mx, kFbTargetExp);
            T += static_cast<double>(shift);
            const unsigned char* grow = scratch + static_cast<size_t>(r) * row_bytes;
            const double sr = ldcg64(grow + scale_off);
            if (r > 0) {                                         // pull row r - 1 of the scratch into L2 while row r is processed
                const unsigned char* nrow = F.scratch + static_cast<size_t>(blockIdx.x) * F.scratch_stride +
                                            static_cast<size_t>(r - 1) * row_bytes;
                for (uint32_t o = threadIdx.x * 128u; o < static_cast<uint32_t>(row_bytes); o += W * 32 * 128u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + o));
            }
            for (int u = hb; u < he; ++u) fbh_run(u, colb, etab, xa, static_cast<double>(shift), t2tab);
            if (warp == 0) *reinterpret_cast<double*>(colp + S.end_slot_off) = (r == n && valid) ? 1.0 : 0.0;
            {
                // c_r = 2^(S_n - S_r - T_r) / P^ : the exponent alone may leave the double range although the product
                // does not (a tiny P^ after an outlier in the last rows), so it is applied in two halves
                const double ed = sn - sr - T;
                const bool active = valid && r <= n;
                const int ei = static_cast<int>(fmin(fmax(ed, -2000.0), 2000.0));
                const int e1 = ei / 2, e2 = ei - e1;
                const double cc = (invp * __hiloint2double((e1 + 1023) << 20, 0)) * __hiloint2double((e2 + 1023) << 20, 0);
                const bool in_range = cc > 1e-300 && cc < 1e300;
                if (active && !in_range) {
                    if (!failed && warp == 0 && atomicCAS(F.fail + 1, 0, 1) == 0) {
                        F.fail[2] = ev; F.fail[3] = r;
                        reinterpret_cast<double*>(F.fail + 7)[0] = ed;
                        reinterpret_cast<double*>(F.fail + 7)[1] = invp;
                    }
                    failed = true;
                }
                c = (active && in_range) ? cc : 0.0;
            }
            __syncthreads();
            const BwdCtx C{c, xbv, flagsa};
            AccCtx A{acc_cta + static_cast<size_t>(F.acc_base[warp]) * kLanes, lane};
            int mxs = 0;
            for (int u = tb; u < te; ++u) fbb_run<1>(u, colb, wtab, grow, A, C, mxs);
            __syncthreads();
            mx = 0;
            for (int u = eb; u < ee; ++u) fbb_run<2>(u, colb, wtab, grow, A, C, mx);
            wmax[warp * kLanes + lane] = mx;
            __syncthreads();
            xa = xbv;
            xbv = xpre;
        }
        // final check: the posterior of the start state at row 0 is 1
        if (warp == 0) {
            const double chk = *reinterpret_cast<const double*>(colp + S.start_slot_off) *
                               ldcg64(scratch + S.start_slot_off) * c;
            if (valid && !(fabs(chk - 1.0) < 1e-6)) {
                if (!failed && atomicCAS(F.fail + 1, 0, 2) == 0) {
                    F.fail[2] = ev; F.fail[3] = n;
                    reinterpret_cast<double*>(F.fail + 7)[0] = chk;
                    reinterpret_cast<double*>(F.fail + 7)[1] = c;
                    reinterpret_cast<double*>(F.fail + 7)[2] = *reinterpret_cast<const double*>(colp + S.start_slot_off);
                    reinterpret_cast<double*>(F.fail + 7)[3] = T;
                }
                failed = true;
            }
        }
        if (__any_sync(0xffffffffu, failed)) {
            const uint32_t b = __ballot_sync(0xffffffffu, failed);
            if (lane == 0 && warp == 0) atomicAdd(F.fail, __popc(b));   // every warp sees the same range failures
        }
        // observed range of every event that gave a state non-zero weight (304-326)
        for (int k = warp; k < S.ne; k += W) {
            const uint32_t mask = s_flags[k];
            if (mask == 0u) continue;
            double lo = ((mask >> lane) & 1u) ? xlo : __longlong_as_double(0x7FF0000000000000LL);
            double hi = ((mask >> lane) & 1u) ? xhi : neg_inf();
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0) {
                atomic_min_f64_g(F.minmax + 2 * k, lo);
                atomic_max_f64_g(F.minmax + 2 * k + 1, hi);
            }
        }
    }
}

// Folds the accumulation slots of all CTAs into the statistics: lanes that hold partial sums of the same term are
// consecutive; a segmented shuffle adds them in a fixed order and one lane does a plain read-modify-write.
static __global__ void k_fb_finalize(const double* __restrict__ acc, int n_ctas, int n_slots,
                                     const int32_t* __restrict__ acc_map, const int32_t* __restrict__ acc_log2,
                                     int n_edges, double* __restrict__ edge_counts, double* __restrict__ emit_moments) {
    const int slot = blockIdx.x, lane = threadIdx.x;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    const double* p = acc + static_cast<size_t>(slot) * kLanes + lane;
    const size_t cs = static_cast<size_t>(n_slots) * kLanes;
    int c = 0;
    for (; c + 4 <= n_ctas; c += 4) {
        v0 += p[(c + 0) * cs];
        v1 += p[(c + 1) * cs];
        v2 += p[(c + 2) * cs];
        v3 += p[(c + 3) * cs];
    }
    for (; c < n_ctas; ++c) v0 += p[c * cs];
    double v = (v0 + v1) + (v2 + v3);
    const int s = acc_log2[slot];
    const int width = 32 >> s;                                   // lanes per term
    for (int o = width >> 1; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, width);
    const int stat = acc_map[slot * kLanes + lane];
    if ((lane & (width - 1)) == 0 && stat >= 0) {
        if (stat < n_edges) edge_counts[stat] += v;
        else emit_moments[stat - n_edges] += v;
    }
}

static __global__ void k_fb_minmax_merge(const double* __restrict__ mm, int ne, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ne) return;
    out[2 * k] = fmin(out[2 * k], mm[2 * k]);
    out[2 * k + 1] = fmax(out[2 * k + 1], mm[2 * k + 1]);
}

}  // namespace pp
