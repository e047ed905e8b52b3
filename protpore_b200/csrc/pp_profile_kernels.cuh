// pp_profile_kernels.cuh -- Viterbi and forward sweeps of a LINEAR PROFILE HMM (pp_profile.h) for sm_100a, register-resident.
//
// One CTA = NW position warps + 1 chain warp sweeps 32 events; lane j of every warp works on event j (as in
// pp_lane_kernels.cuh).  Position warp w owns positions [k0, k0 + P), k0 = w P: the values of THEIR emitting states --
// M_k, I_BLIP_k, I_INS_k -- stay in registers from row to row.
//
// What serialises a row is the pair of silent chains: S_SKIP_k = max(A_k, S_SKIP_{k-1} + c_k) ascending and
// B_BACK_k = max(A'_k, B_BACK_{k+1} + c'_k) descending, with pass-1 candidates A_k = M_{k-1} + a_k, A'_k = M_k + a'_k
// (yahmm.pyx:3564-3605) -- 43 dependent links of DADD -> DSETP -> FSEL (22.7 cycles each on B200, tools/lat_bench.cu)
// that a scan cannot re-associate without changing bits.  They are taken off the critical path EXACTLY:
//   * rounding is monotone, so max(x, y) + c = max(x + c, y + c) bit for bit.  Unrolling the recurrence over the P links
//     of a block gives  out = max( (((in + c_1) + c_2) ... + c_P),  E ), where E -- the best way INTO the block through
//     one of its own pass-1 candidates, summed left to right like the reference does -- does not depend on `in`;
//   * every position warp computes E for its block from its own M values (plus the one entry that comes from its left
//     neighbour's last M, which that neighbour carries through for it: T);
//   * the chain warp only propagates the block CARRIES: P dependent DADDs and one maximum per block (~46 cycles instead
//     of P x 22.7), both chains interleaved, and publishes the value entering every block;
//   * each position warp then walks ITS P links from that value with the reference's own operations (strict '>', which
//     also gives the ordinals) -- all warps at once.
// Per row r -> r + 1, position warps:
//   early  beside the chain warp's carry sweep of row r: everything of row r + 1 that does not read the chains -- the
//          emissions, five (six) of M's seven (eight) candidates and their first-maximum tree, the insert and blip states;
//   fill   after the barrier: the warp's links of row r's chains from the published carries (and their ordinals);
//   late   the two candidates of M_k(r + 1) that read S_SKIP_{k-1}(r) and B_BACK_{k+1}(r), merged in the reference's order
//          (slot 1 beats every later slot on equality and loses to slot 0; slot 7 replaces only if strictly greater);
//          publishes M_k(r + 1), the pass-1 candidates it feeds and the block summaries E, T of row r + 1.
// S_DROPOFF_k = M_k + t has a single in-edge and is recomputed where it is read (same DADD as the reference's).  The end
// state is evaluated once per event at its last row (a CTA-uniform extra step).
//
// Shared memory is laid out per WARP: block w + 1 holds, for the warp's P positions, M | the two pass-1 candidate arrays
// | S_SKIP | B_BACK (kept only for the end state) and the block's E, T, E', carry-in values, 32 doubles (one per lane)
// each; blocks 0 and NW + 1 are phantoms holding "nothing beyond the ends".  A warp reaches everything it needs -- its
// own block and the edge slots of its two neighbours -- at COMPILE-TIME offsets from one per-thread base address, so a
// candidate costs no address arithmetic and no load for its value (registers), and half a load for its weight (16-byte
// broadcast loads of the position record).
//
// Arithmetic is the reference's, candidate by candidate: (v + t) + e for emitting states (3545-3562), v + t for silent
// ones, pass 1 before pass 2, strict '>' in the reference's visiting order -- slot order of pp_profile.h; the balanced
// first-maximum tree picks the same winner as the left-to-right scan.
//
// Forward: scaled linear space as in k_forward_lane (exact power-of-two row scaling steered by the largest emitting
// value of the previous row, collected with one shared-memory atomicMax per warp and row).  Its chains are linear
// recurrences, so a block is an affine map out = in * C + E with C the (constant) product of the block's chain weights;
// events whose result is 0 / inf / NaN are re-run by the exact log-space kernel (pp_api.cu).
#pragma once
#include "pp_device_common.cuh"
#include "pp_lane_types.h"
#include "pp_profile_sched.h"

namespace pp {

#ifndef PP_PROF_MIN_CTAS
#define PP_PROF_MIN_CTAS 2            // CTAs per SM the register budget is set for
#endif
#define PP_PROF_THREADS ((kProfMaxPositionWarps + 1) * 32)

// Optional per-warp phase clocks (tools/profile_phase_timing.py builds the library with -DPP_PHASE_TIMING)
#ifdef PP_PHASE_TIMING
__device__ unsigned long long g_prof_phase[2][32 * 8];
#define PPK_CLK(var) const long long var = clock64()
#define PPK_ACC(slot, a, b) phase_acc[slot] += static_cast<unsigned long long>((b) - (a))
#else
#define PPK_CLK(var)
#define PPK_ACC(slot, a, b)
#endif

// byte offsets inside a warp block: 5 arrays of P slots and 5 single slots; a slot = 32 doubles, one per lane
template <int P>
struct ProfBlock {
    static constexpr int SLOT = kSlotBytes;
    static constexpr int MX = 0;                 // M_k
    static constexpr int AS = P * SLOT;          // M_k + t(M_k -> S_SKIP_{k+1}): pass-1 candidate of S_SKIP_{k+1}
    static constexpr int AB = 2 * P * SLOT;      // M_k + t(M_k -> B_BACK_k):     pass-1 candidate of B_BACK_k
    static constexpr int SK = 3 * P * SLOT;      // S_SKIP_k, B_BACK_k: written only in rows where an event ends
    static constexpr int BK = 4 * P * SLOT;
    static constexpr int EO = 5 * P * SLOT;      // E of the ascending chain over the block's own entries
    static constexpr int TT = EO + SLOT;         // T: the entry from the left neighbour's last M, carried through the block
    static constexpr int ED = EO + 2 * SLOT;     // E' of the descending chain
    static constexpr int VIN = EO + 3 * SLOT;    // S_SKIP_{k0-1}: value entering the block from below
    static constexpr int UIN = EO + 4 * SLOT;    // B_BACK_{k0+P}: value entering the block from above
    static constexpr int WS = EO + 5 * SLOT;     // bytes per block
};

__host__ __device__ inline size_t prof_smem_bytes(int Kr, int NW, int P) {
    // records | NW + 2 warp blocks | observation tile | row maxima [2][32]
    return static_cast<size_t>(Kr) * kProfRecDoubles * 8 + static_cast<size_t>(NW + 2) * (5 * P + 5) * kSlotBytes +
           static_cast<size_t>(kLaneTileCols) * (kLanes + 1) * 8 + 2 * kLanes * 4 + 64;
}

template <int P>
struct ProfSmem {
    double* rec;
    unsigned char* blocks;                       // block b = w + 1
    double* xt;
    int32_t* smax;                               // [2][32]
    int n_col;                                   // doubles in the blocks
    __device__ __forceinline__ ProfSmem(unsigned char* smem, int Kr, int NW) {
        rec = reinterpret_cast<double*>(smem);
        blocks = smem + static_cast<size_t>(Kr) * kProfRecDoubles * 8;
        n_col = (NW + 2) * (5 * P + 5) * kLanes;
        xt = reinterpret_cast<double*>(blocks) + n_col;
        smax = reinterpret_cast<int32_t*>(xt + kLaneTileCols * (kLanes + 1));
    }
};

// CTA-wide barrier for warp-specialised code: both roles arrive at barrier 1 with the CTA's thread count
__device__ __forceinline__ void prof_bar(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }
// power-of-two shift of the next row from the largest high word among the emitting values of this one
__device__ __forceinline__ int prof_row_shift(int hi) {
    const int ex = (hi >> 20) & 0x7ff;
    const int shift = (ex == 0 || ex == 0x7ff) ? 0 : (kFwdTargetExp + 1023 - ex);
    return max(-900, min(900, shift));
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double ldb(const unsigned char* p, int off) { return *reinterpret_cast<const double*>(p + off); }
__device__ __forceinline__ void stb(unsigned char* p, int off, double v) { *reinterpret_cast<double*>(p + off) = v; }
// slot of position k in array `arr` (this lane's slice when `blocks` already carries the lane offset)
template <int P>
__device__ __forceinline__ int prof_slot(int k, int arr) { return (k / P + 1) * ProfBlock<P>::WS + arr + (k % P) * kSlotBytes; }

// observation tile with a run-time warp count
__device__ __forceinline__ void prof_fill_tile(double* xt, const void* __restrict__ obs, bool f64, int64_t off, int n, int i0,
                                               int warp, int lane, int n_warps) {
    constexpr int TC = kLaneTileCols;
    for (int j = warp * (32 / TC) + lane / TC; j < kLanes; j += n_warps * (32 / TC)) {
        const int64_t oj = __shfl_sync(0xffffffffu, off, j);
        const int nj = __shfl_sync(0xffffffffu, n, j);
        const int c = i0 + (lane % TC);
        double v = 0.0;
        if (c < nj) v = f64 ? __ldg(static_cast<const double*>(obs) + oj + c)
                            : static_cast<double>(__ldg(static_cast<const float*>(obs) + oj + c));
        xt[(lane % TC) * (kLanes + 1) + j] = v;
    }
}

// ---- emissions ------------------------------------------------------------------------------------------------
// exact log-densities (NormalDistribution / UniformDistribution._log_probability, yahmm.pyx:377-390, 257-262)
__device__ __forceinline__ double prof_log_normal(const double* R, double x) {
    const double2 a = ld2(R), b = ld2(R + 2);                     // {mu, 2 sigma^2}, {RN(1 / (2 sigma^2)), log peak}
    const double d = __dsub_rn(x, a.x);
    const double d2 = __dmul_rn(d, d);
    const double q0 = __dmul_rn(d2, b.x);
    const double rem = __fma_rn(-q0, a.y, d2);
    const double q = __fma_rn(rem, b.x, q0);                      // = RN(d2 / (2 sigma^2)), pp_device_common.cuh
    return __dsub_rn(b.y, q);
}
__device__ __forceinline__ double prof_log_point(const double* R, double x) {
    return (fabs(__dsub_rn(x, R[0])) < 1E-4) ? 0.0 : neg_inf();
}
__device__ __forceinline__ double prof_log_uniform(const double* R, double x) {
    const double2 a = ld2(R);
    return (x >= a.x && x <= a.y) ? R[3] : neg_inf();
}
// linear-space factors p(x) 2^shift; scale2 = 2^shift as a double
__device__ __forceinline__ double prof_lin_normal(const double* R, double x, double shiftd) {
    const double2 a = ld2(R);                                     // {mu, log2(e) / (2 sigma^2)}
    const double d = x - a.x;
    return exp2_scaled(fma(-(d * d), a.y, R[2]) + shiftd);
}
__device__ __forceinline__ double prof_lin_point(const double* R, double x, double scale2) {
    return (fabs(x - R[0]) < 1E-4) ? R[3] * scale2 : 0.0;
}
__device__ __forceinline__ double prof_lin_uniform(const double* R, double x, double scale2) {
    const double2 a = ld2(R);
    return (x >= a.x && x <= a.y) ? R[3] * scale2 : 0.0;
}

// =================================================================================================================
// Viterbi
// =================================================================================================================
// Of the candidates of M_k(r + 1) only two depend on the silent chains of row r: the one from S_SKIP_{k-1} (slot 1) and
// the one from B_BACK_{k+1} (slot 7).  Everything else -- the emissions, the other five (six) candidates and their
// first-maximum tree, the insert and blip states, which have no silent source at all -- is computed by the position warps
// WHILE the chain warps walk the chains of row r (`early`); after the barrier only the two missing candidates are merged
// in (`late`).  The merge keeps the reference's order: slot 1 beats every later slot on equality and loses to slot 0;
// slot 7 replaces only if strictly greater.
//
// early: rows r -> r + 1 of the warp's P positions.  tb: this lane's slice of the warp's block; left / right: M of the
// neighbouring positions k0 - 1 and k0 + P in row r.  On return vM[p] holds the PARTIAL maximum of M_k(r + 1), eM[p] its
// emission, bits 5 p .. 5 p + 4 of the returned word {partial arg (3), insert ordinal, blip ordinal}.
template <int P>
__device__ __forceinline__ uint32_t prof_vit_early(const double* __restrict__ recw, const unsigned char* __restrict__ tb,
                                                   double left, double right, double x, double (&vM)[P], double (&vB)[P],
                                                   double (&vI)[P], double (&eMv)[P]) {
    using L = ProfBlock<P>;
    uint32_t ord = 0;
#pragma unroll
    for (int p = P - 1; p >= 0; --p) {                            // descending: vM[p - 1] is still row r
        const double* R = recw + p * kProfRecDoubles;
        const double2 w01 = ld2(R + PR_WM), w23 = ld2(R + PR_WM + 2), w45 = ld2(R + PR_WM + 4);
        const double2 wi = ld2(R + PR_WI), wb = ld2(R + PR_WB);
        const double2 wdk = ld2(R + PR_WD);
        const int kinds = __double2loint(wdk.y);
        const double mprev = p > 0 ? vM[p > 0 ? p - 1 : 0] : left;
        const double eM = prof_log_normal(R + PR_EM, x);
        const double eB = prof_log_uniform(R + PR_EB, x);
        const double eI = ((kinds >> 8) & 0xff) == EMIS_POINT ? prof_log_point(R + PR_EI, x) : prof_log_normal(R + PR_EI, x);
        const double drop = __dadd_rn(vM[p], wdk.x);              // S_DROPOFF_k(r): its one candidate (3564-3584)
        double s[5];
        s[0] = __dadd_rn(__dadd_rn(mprev, w01.x), eM);            // slot 0
        s[1] = __dadd_rn(__dadd_rn(vM[p], w23.x), eM);            // slots 2 .. 5
        s[2] = __dadd_rn(__dadd_rn(drop, w23.y), eM);
        s[3] = __dadd_rn(__dadd_rn(vB[p], w45.x), eM);
        s[4] = __dadd_rn(__dadd_rn(vI[p], w45.y), eM);
        double best;
        uint32_t arg;
        FirstMax<0, 5>::run(s, best, arg);
        arg += arg != 0 ? 1u : 0u;                                // tree index -> slot
        if (kinds >> 24) {
            // a candidate from M_{k+1} sits behind the insert's (bake merged a back-slip state away): strictly greater only.
            // M_{k+1}(r) is still in the exchange: the warps store their new M values in the late part.
            const double mnext = p < P - 1 ? ldb(tb, L::MX + (p + 1) * L::SLOT) : right;
            const double c = __dadd_rn(__dadd_rn(mnext, ld2(R + PR_WM + 6).x), eM);
            if (c > best) { best = c; arg = 7; }
        }
        const double i0 = __dadd_rn(__dadd_rn(mprev, wi.x), eI), i1 = __dadd_rn(__dadd_rn(vI[p], wi.y), eI);
        const double b0 = __dadd_rn(__dadd_rn(vM[p], wb.x), eB), b1 = __dadd_rn(__dadd_rn(vB[p], wb.y), eB);
        const bool pi = i1 > i0, pb = b1 > b0;
        vM[p] = best;
        eMv[p] = eM;
        vI[p] = pi ? i1 : i0;
        vB[p] = pb ? b1 : b0;
        ord |= (arg | (pi ? 8u : 0u) | (pb ? 16u : 0u)) << (5 * p);
    }
    return ord;
}
// fill: the warp's P links of each chain of row r, from the carries the chain warp published.  skin[j] = S_SKIP_{k0+j-1},
// bkin[j] = B_BACK_{k0+j+1}: what position j reads.  Bit j of the result = ordinal of S_SKIP_{k0+j}, bit P + j of
// B_BACK_{k0+j} (pass 2 replaces only if strictly greater, 3586-3605).  keep: an event ends at this row, the end state
// will read the values.
template <int P>
__device__ __forceinline__ uint32_t prof_vit_fill(const double* __restrict__ recw, unsigned char* __restrict__ tb, bool keep,
                                                  double (&skin)[P], double (&bkin)[P]) {
    using L = ProfBlock<P>;
    double A[P], D[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        A[j] = ldb(tb, j > 0 ? L::AS + (j - 1) * L::SLOT : -L::WS + L::AS + (P - 1) * L::SLOT);   // M_{k0+j-1} + a
        D[j] = ldb(tb, L::AB + j * L::SLOT);
    }
    double v = ldb(tb, L::VIN), u = ldb(tb, L::UIN);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int jd = P - 1 - j;
        skin[j] = v;
        bkin[jd] = u;
        const double t = __dadd_rn(v, recw[j * kProfRecDoubles + PR_SK + 1]);
        const double td = __dadd_rn(u, recw[jd * kProfRecDoubles + PR_BK + 1]);
        const bool p = t > A[j], pd = td > D[jd];
        v = p ? t : A[j];
        u = pd ? td : D[jd];
        bits |= (p ? 1u : 0u) << j;
        bits |= (pd ? 1u : 0u) << (P + jd);
        if (keep) {
            stb(tb, L::SK + j * L::SLOT, v);
            stb(tb, L::BK + jd * L::SLOT, u);
        }
    }
    return bits;
}
// late: merges the two chain-dependent candidates, publishes M_k(r + 1), the pass-1 candidates of the chain states it
// feeds (3564-3584) and the block summaries of row r + 1 for the chain warp.
template <int P>
__device__ __forceinline__ uint32_t prof_vit_late(const double* __restrict__ recw, unsigned char* __restrict__ tb, uint32_t ord,
                                                  double (&vM)[P], const double (&eMv)[P], const double (&skin)[P],
                                                  const double (&bkin)[P]) {
    using L = ProfBlock<P>;
    double as[P], ab[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double* R = recw + p * kProfRecDoubles;
        const double s1 = __dadd_rn(__dadd_rn(skin[p], R[PR_WM + 1]), eMv[p]);
        const double s7 = __dadd_rn(__dadd_rn(bkin[p], R[PR_WM + 7]), eMv[p]);
        double m = vM[p];
        const uint32_t a = (ord >> (5 * p)) & 7u;
        const bool t1 = s1 > m || (s1 == m && a != 0);
        m = t1 ? s1 : m;
        const bool t7 = s7 > m;
        m = t7 ? s7 : m;
        const uint32_t na = t7 ? 6u : (t1 ? 1u : a);              // stored ordinal: 6 = back-slip, 7 = M_{k+1}
        ord = (ord & ~(7u << (5 * p))) | (na << (5 * p));
        vM[p] = m;
        as[p] = __dadd_rn(m, R[PR_SK]);                           // into S_SKIP_{k+1}
        ab[p] = __dadd_rn(m, R[PR_BK]);                           // into B_BACK_k
        stb(tb, L::MX + p * L::SLOT, m);
        stb(tb, L::AS + p * L::SLOT, as[p]);
        stb(tb, L::AB + p * L::SLOT, ab[p]);
    }
    // E: links k0 + 1 .. k0 + P - 1 entered through one of their own pass-1 candidates as[0 .. P - 2]
    double e = as[0];
#pragma unroll
    for (int j = 1; j < P - 1; ++j) {
        const double t = __dadd_rn(e, recw[(j + 1) * kProfRecDoubles + PR_SK + 1]);
        e = t > as[j] ? t : as[j];
    }
    stb(tb, L::EO, e);
    // T: as[P - 1] is the first entry of the RIGHT neighbour's block; carried through that block's other links
    double t = as[P - 1];
#pragma unroll
    for (int j = 1; j < P; ++j) t = __dadd_rn(t, recw[(P + j) * kProfRecDoubles + PR_SK + 1]);
    stb(tb, L::WS + L::TT, t);
    // E': links k0 + P - 1 .. k0 of the descending chain
    double ed = ab[P - 1];
#pragma unroll
    for (int j = P - 2; j >= 0; --j) {
        const double td = __dadd_rn(ed, recw[j * kProfRecDoubles + PR_BK + 1]);
        ed = td > ab[j] ? td : ab[j];
    }
    stb(tb, L::ED, ed);
    return ord;
}

// The chain warp: block carries of both chains of one row, one block of each per step (two independent dependent
// sequences).  cb: this lane's slice of block 0.  Publishes the value entering every block.
template <int P>
__device__ __forceinline__ void prof_vit_carries(const double* __restrict__ rec, unsigned char* __restrict__ cb, int NW,
                                                 double startval) {
    using L = ProfBlock<P>;
    double v = startval, u = neg_inf();
    double c[P], d[P], tt, eo, ed;
    // operands of step i: the ascending chain's block i and the descending chain's block NW - 1 - i
    auto load = [&](int i, double (&cc)[P], double (&dd)[P], double& t1, double& e1, double& e2) {
        const unsigned char* bu = cb + (i + 1) * L::WS;
        const unsigned char* bd = cb + (NW - i) * L::WS;
        t1 = ldb(bu, L::TT);
        e1 = ldb(bu, L::EO);
        e2 = ldb(bd, L::ED);
#pragma unroll
        for (int j = 0; j < P; ++j) {
            cc[j] = rec[(i * P + j) * kProfRecDoubles + PR_SK + 1];
            dd[j] = rec[((NW - 1 - i) * P + (P - 1 - j)) * kProfRecDoubles + PR_BK + 1];
        }
    };
    load(0, c, d, tt, eo, ed);
    for (int i = 0; i < NW; ++i) {
        // the next step's operands are requested first and touched last: a warp issues in order, and the links below
        // must not wait behind a load that has nothing to do with them
        double cn[P], dn[P], ttn = 0.0, eon = 0.0, edn = 0.0;
        if (i + 1 < NW) load(i + 1, cn, dn, ttn, eon, edn);
        stb(cb + (i + 1) * L::WS, L::VIN, v);
        stb(cb + (NW - i) * L::WS, L::UIN, u);
        double th = v, td = u;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            th = __dadd_rn(th, c[j]);
            td = __dadd_rn(td, d[j]);
        }
        const double ee = tt > eo ? tt : eo;
        v = th > ee ? th : ee;
        u = td > ed ? td : ed;
#pragma unroll
        for (int j = 0; j < P; ++j) { c[j] = cn[j]; d[j] = dn[j]; }
        tt = ttn;
        eo = eon;
        ed = edn;
    }
}

// One chain per warp (two chain warps): half the instructions on the row's critical path
template <int P, bool UP>
__device__ __forceinline__ void prof_vit_carry_one(const double* __restrict__ rec, unsigned char* __restrict__ cb, int NW,
                                                   double startval) {
    using L = ProfBlock<P>;
    double v = startval;
    double c[P], tt = neg_inf(), eo;
    auto load = [&](int i, double (&cc)[P], double& t1, double& e1) {
        if (UP) {
            const unsigned char* bu = cb + (i + 1) * L::WS;
            t1 = ldb(bu, L::TT);
            e1 = ldb(bu, L::EO);
#pragma unroll
            for (int j = 0; j < P; ++j) cc[j] = rec[(i * P + j) * kProfRecDoubles + PR_SK + 1];
        } else {
            const unsigned char* bd = cb + (NW - i) * L::WS;
            e1 = ldb(bd, L::ED);
#pragma unroll
            for (int j = 0; j < P; ++j) cc[j] = rec[((NW - 1 - i) * P + (P - 1 - j)) * kProfRecDoubles + PR_BK + 1];
        }
    };
    load(0, c, tt, eo);
    for (int i = 0; i < NW; ++i) {
        double cn[P], ttn = neg_inf(), eon = 0.0;
        if (i + 1 < NW) load(i + 1, cn, ttn, eon);
        if (UP) stb(cb + (i + 1) * L::WS, L::VIN, v);
        else stb(cb + (NW - i) * L::WS, L::UIN, v);
        double th = v;
#pragma unroll
        for (int j = 0; j < P; ++j) th = __dadd_rn(th, c[j]);
        const double ee = UP ? (tt > eo ? tt : eo) : eo;
        v = th > ee ? th : ee;
#pragma unroll
        for (int j = 0; j < P; ++j) c[j] = cn[j];
        tt = ttn;
        eo = eon;
    }
}

template <int P, int NC>
__global__ void __launch_bounds__(PP_PROF_THREADS + (NC - 1) * 32, PP_PROF_MIN_CTAS)
k_profile_viterbi(const ProfSched S, const LaneBatchAny B, const VitOut O) {
    using L = ProfBlock<P>;
    extern __shared__ __align__(16) unsigned char smem[];
    const ProfSmem<P> M(smem, S.Kr, S.NW);
    constexpr int TC = kLaneTileCols;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int n_warps = S.NW + NC;
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    constexpr int UB = 5 * P <= 16 ? 2 : 4;

    for (int i = threadIdx.x; i < S.Kr * kProfRecDoubles; i += n_warps * 32) M.rec[i] = __ldg(S.rec + i);
    for (int i = threadIdx.x; i < M.n_col; i += n_warps * 32) reinterpret_cast<double*>(M.blocks)[i] = neg_inf();   // v[0, :] = -inf (3513)
    if (nmax > 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, 0, warp, lane, n_warps);
    __syncthreads();

    unsigned char* const cb = M.blocks + lane * 8;                               // block 0, this lane
    const int n_threads = n_warps * 32;

    // The two roles run separate loops (separate register allocations) that meet at the same barriers:
    //   barrier 1 after {carries of row r | early part of row r + 1}, [end state of events ending at row r, between two
    //   more barriers], barrier 2 after the late part of row r + 1.
    if (warp == 0) {                                                // the chain warp: the oldest warp of the CTA
#ifdef PP_PHASE_TIMING
        unsigned long long phase_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        for (int r = 0;; ++r) {
            // ---- silent chains of row r; v[0, start] = 0 (3514), the start state is -inf in every later row
            PPK_CLK(c0);
            if (NC == 1) prof_vit_carries<P>(M.rec, cb, S.NW, r == 0 ? 0.0 : neg_inf());
            else prof_vit_carry_one<P, true>(M.rec, cb, S.NW, r == 0 ? 0.0 : neg_inf());
            PPK_CLK(c1);
            PPK_ACC(0, c0, c1);
#ifdef PP_PHASE_TIMING
            phase_acc[6] += 1;
            if (r == nmax && lane == 0)
                for (int k = 0; k < 8; ++k) atomicAdd(&g_prof_phase[0][k], phase_acc[k]);
#endif
            prof_bar(n_threads);
            if (r >= 1 && __any_sync(0xffffffffu, r == n)) {
                // ---- events whose last row this is: the end state (3614-3616).  Every warp holds the same lengths, so
                //      the branch (and its barriers) is uniform over the CTA.
                prof_bar(n_threads);                                             // the position warps have filled in row r
                double best = neg_inf();
                int arg = 0;
                for (int j = 0; j < S.end_n; ++j) {
                    const ProfEndRec E = ldg_rec(S.end_list + j);
                    double v;
                    if (E.cls == PROF_M) v = ldb(cb, prof_slot<P>(E.k, L::MX));
                    else if (E.cls == PROF_DROP) v = __dadd_rn(ldb(cb, prof_slot<P>(E.k, L::MX)), M.rec[E.k * kProfRecDoubles + PR_WD]);
                    else if (E.cls == PROF_SKIP) v = ldb(cb, prof_slot<P>(E.k, L::SK));
                    else v = ldb(cb, prof_slot<P>(E.k, L::BK));
                    const double s = __dadd_rn(v, E.w);
                    if (s > best) { best = s; arg = j; }
                }
                if (r == n) {
                    O.score[ev] = best;
                    O.end_state[ev] = S.end_state;
                    O.end_ord[ev] = arg;
                }
                prof_bar(n_threads);
            }
            if (r == nmax) break;
            if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
            prof_bar(n_threads);
        }
        return;
    }

    if (NC == 2 && warp == 1) {                                     // the second chain warp: the descending chain
        for (int r = 0;; ++r) {
            prof_vit_carry_one<P, false>(M.rec, cb, S.NW, neg_inf());
            prof_bar(n_threads);
            if (r >= 1 && __any_sync(0xffffffffu, r == n)) {
                prof_bar(n_threads);
                prof_bar(n_threads);
            }
            if (r == nmax) break;
            if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
            prof_bar(n_threads);
        }
        return;
    }

    double vM[P], vB[P], vI[P], eM[P];
#pragma unroll
    for (int p = 0; p < P; ++p) vM[p] = vB[p] = vI[p] = neg_inf();
    const int pw = warp - NC;                                                    // position warp index
    unsigned char* const tb = cb + (pw + 1) * L::WS;                             // this warp's block
    const double* const recw = M.rec + pw * P * kProfRecDoubles;
    // this warp's units of row 0 in the pointer table: emitting ordinals (written with row r + 1), chain ordinals (row r)
    uint8_t* prow = reinterpret_cast<uint8_t*>(O.ptr) + static_cast<size_t>(B.group_row[g]) * S.row_bytes + pw * UB * kLanes + lane * UB;
    constexpr int CB = 2 * P <= 8 ? 1 : 2;                                       // bytes per lane of the chain ordinals
    const int sil_off = S.NW * UB * kLanes + pw * CB * kLanes + lane * CB - (pw * UB * kLanes + lane * UB);
#ifdef PP_PHASE_TIMING
    unsigned long long phase_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int r = 0;; ++r) {
        PPK_CLK(c0);
        uint32_t ord = 0;
        if (r < nmax) {
            // ---- beside the carry sweep of row r: everything of row r + 1 that does not read the chains (3545-3562)
            const double left = ldb(tb, -L::WS + L::MX + (P - 1) * L::SLOT);     // M_{k0-1}(r), M_{k0+P}(r)
            const double right = ldb(tb, L::WS + L::MX);
            const double x = M.xt[(r & (TC - 1)) * (kLanes + 1) + lane];         // obs[r] belongs to row r + 1
            ord = prof_vit_early<P>(recw, tb, left, right, x, vM, vB, vI, eM);
        }
        PPK_CLK(c1);
        prof_bar(n_threads);
        PPK_CLK(c2);
        PPK_ACC(0, c0, c1); PPK_ACC(1, c1, c2);
        // ---- the warp's links of row r's chains
        const bool ends = r >= 1 && __any_sync(0xffffffffu, r == n);
        double skin[P], bkin[P];
        const uint32_t sbits = prof_vit_fill<P>(recw, tb, ends, skin, bkin);
        if (CB == 1) prow[sil_off] = static_cast<uint8_t>(sbits);
        else *reinterpret_cast<uint16_t*>(prow + sil_off) = static_cast<uint16_t>(sbits);
        prow += S.row_bytes;
        if (ends) {                                                              // the chain warp evaluates the end state
            prof_bar(n_threads);
            prof_bar(n_threads);
        }
        if (r == nmax) break;
        PPK_CLK(c3);
        // ---- the two candidates that read the chains of row r; publish M(r + 1)
        ord = prof_vit_late<P>(recw, tb, ord, vM, eM, skin, bkin);
        if (UB == 2) *reinterpret_cast<uint16_t*>(prow) = static_cast<uint16_t>(ord);
        else *reinterpret_cast<uint32_t*>(prow) = ord;
        // the observation tile is free now (the early part of this row has read it): stage the next one when due
        if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
        PPK_CLK(c4);
        prof_bar(n_threads);
        PPK_CLK(c5);
        PPK_ACC(2, c2, c3); PPK_ACC(3, c3, c4); PPK_ACC(4, c4, c5);
#ifdef PP_PHASE_TIMING
        phase_acc[6] += 1;
#endif
    }
#ifdef PP_PHASE_TIMING
    if (lane == 0)
        for (int k = 0; k < 8; ++k) atomicAdd(&g_prof_phase[0][warp * 8 + k], phase_acc[k]);
#endif
}

// =================================================================================================================
// Forward (scaled linear space)
// =================================================================================================================
// Same split: `early` sums every term that does not read the chains of row r and finishes the insert and blip states,
// `late` adds the two chain terms and applies the emission.  Returns the largest high word among the insert / blip values.
template <int P>
__device__ __forceinline__ int prof_fwd_early(const double* __restrict__ recw, const unsigned char* __restrict__ tb, double left,
                                              double right, double x, double shiftd, double scale2, double (&vM)[P],
                                              double (&vB)[P], double (&vI)[P], double (&eMv)[P]) {
    using L = ProfBlock<P>;
    int mx = 0;
#pragma unroll
    for (int p = P - 1; p >= 0; --p) {
        const double* R = recw + p * kProfRecDoubles;
        const double2 w01 = ld2(R + PR_WM), w23 = ld2(R + PR_WM + 2), w45 = ld2(R + PR_WM + 4);
        const double2 wi = ld2(R + PR_WI), wb = ld2(R + PR_WB);
        const double2 wdk = ld2(R + PR_WD);
        const int kinds = __double2loint(wdk.y);
        const double mprev = p > 0 ? vM[p > 0 ? p - 1 : 0] : left;
        eMv[p] = prof_lin_normal(R + PR_EM, x, shiftd);
        const double eB = prof_lin_uniform(R + PR_EB, x, scale2);
        const double eI = ((kinds >> 8) & 0xff) == EMIS_POINT ? prof_lin_point(R + PR_EI, x, scale2)
                                                              : prof_lin_normal(R + PR_EI, x, shiftd);
        const double drop = vM[p] * wdk.x;
        double a0 = mprev * w01.x;                                  // two accumulators halve the dependent DFMA chain
        double a1 = drop * w23.y;
        a0 = fma(vM[p], w23.x, a0);
        a1 = fma(vB[p], w45.x, a1);
        a0 = fma(vI[p], w45.y, a0);
        if (kinds >> 24) a1 = fma(p < P - 1 ? ldb(tb, L::MX + (p + 1) * L::SLOT) : right, ld2(R + PR_WM + 6).x, a1);
        const double nI = fma(mprev, wi.x, vI[p] * wi.y) * eI;
        const double nB = fma(vM[p], wb.x, vB[p] * wb.y) * eB;
        vM[p] = a0 + a1;
        vI[p] = nI;
        vB[p] = nB;
        mx = max(mx, max(__double2hiint(nI), __double2hiint(nB)));
    }
    return mx;
}
// fill (linear): the warp's links of row r's chains from the published carries
template <int P>
__device__ __forceinline__ void prof_fwd_fill(const double* __restrict__ recw, unsigned char* __restrict__ tb, bool keep,
                                              double (&skin)[P], double (&bkin)[P]) {
    using L = ProfBlock<P>;
    double A[P], D[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        A[j] = ldb(tb, j > 0 ? L::AS + (j - 1) * L::SLOT : -L::WS + L::AS + (P - 1) * L::SLOT);
        D[j] = ldb(tb, L::AB + j * L::SLOT);
    }
    double v = ldb(tb, L::VIN), u = ldb(tb, L::UIN);
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int jd = P - 1 - j;
        skin[j] = v;
        bkin[jd] = u;
        v = fma(v, recw[j * kProfRecDoubles + PR_SK + 1], A[j]);
        u = fma(u, recw[jd * kProfRecDoubles + PR_BK + 1], D[jd]);
        if (keep) {
            stb(tb, L::SK + j * L::SLOT, v);
            stb(tb, L::BK + jd * L::SLOT, u);
        }
    }
}
template <int P>
__device__ __forceinline__ int prof_fwd_late(const double* __restrict__ recw, unsigned char* __restrict__ tb, int mx,
                                             double (&vM)[P], const double (&eMv)[P], const double (&skin)[P],
                                             const double (&bkin)[P]) {
    using L = ProfBlock<P>;
    double as[P], ab[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double* R = recw + p * kProfRecDoubles;
        const double m = fma(bkin[p], R[PR_WM + 7], fma(skin[p], R[PR_WM + 1], vM[p])) * eMv[p];
        vM[p] = m;
        as[p] = m * R[PR_SK];
        ab[p] = m * R[PR_BK];
        stb(tb, L::MX + p * L::SLOT, m);
        stb(tb, L::AS + p * L::SLOT, as[p]);
        stb(tb, L::AB + p * L::SLOT, ab[p]);
        mx = max(mx, __double2hiint(m));
    }
    double e = as[0];
#pragma unroll
    for (int j = 1; j < P - 1; ++j) e = fma(e, recw[(j + 1) * kProfRecDoubles + PR_SK + 1], as[j]);
    stb(tb, L::EO, e);
    stb(tb, L::WS + L::TT, as[P - 1] * recw[P * kProfRecDoubles + PR_CUT]);
    double ed = ab[P - 1];
#pragma unroll
    for (int j = P - 2; j >= 0; --j) ed = fma(ed, recw[j * kProfRecDoubles + PR_BK + 1], ab[j]);
    stb(tb, L::ED, ed);
    return mx;
}

// the chain warp: affine block carries of both chains
template <int P>
__device__ __forceinline__ void prof_fwd_carries(const double* __restrict__ rec, unsigned char* __restrict__ cb, int NW,
                                                 double startval) {
    using L = ProfBlock<P>;
    double v = startval, u = 0.0;
#pragma unroll 2
    for (int i = 0; i < NW; ++i) {
        unsigned char* bu = cb + (i + 1) * L::WS;
        unsigned char* bd = cb + (NW - i) * L::WS;
        const double ee = ldb(bu, L::TT) + ldb(bu, L::EO), ed = ldb(bd, L::ED);
        const double cu = rec[i * P * kProfRecDoubles + PR_CUP], cd = rec[(NW - 1 - i) * P * kProfRecDoubles + PR_CDN];
        stb(bu, L::VIN, v);
        stb(bd, L::UIN, u);
        v = fma(v, cu, ee);
        u = fma(u, cd, ed);
    }
}

template <int P>
__global__ void __launch_bounds__(PP_PROF_THREADS, PP_PROF_MIN_CTAS)
k_profile_forward(const ProfSched S, const LaneBatchAny B, const FwdOut O) {
    using L = ProfBlock<P>;
    extern __shared__ __align__(16) unsigned char smem[];
    const ProfSmem<P> M(smem, S.Kr, S.NW);
    constexpr int TC = kLaneTileCols;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int n_warps = S.NW + 1;
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];

    for (int i = threadIdx.x; i < S.Kr * kProfRecDoubles; i += n_warps * 32) M.rec[i] = __ldg(S.rec + i);
    for (int i = threadIdx.x; i < M.n_col; i += n_warps * 32) reinterpret_cast<double*>(M.blocks)[i] = 0.0;
    if (threadIdx.x < 2 * kLanes) M.smax[threadIdx.x] = 0;
    if (nmax > 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, 0, warp, lane, n_warps);
    __syncthreads();

    unsigned char* const cb = M.blocks + lane * 8;
    const int n_threads = n_warps * 32;

    if (warp == 0) {                                                // the chain warp: the oldest warp of the CTA
        long long scale_total = 0;                                  // sum of the shifts of rows 1 .. r
        for (int r = 0;; ++r) {
            const int q = r & 1;
            const int shift = prof_row_shift(M.smax[q * kLanes + lane]);
            prof_fwd_carries<P>(M.rec, cb, S.NW, r == 0 ? 1.0 : 0.0);
            M.smax[(q ^ 1) * kLanes + lane] = 0;                    // row r + 1 collects its maxima here in the late part
            prof_bar(n_threads);
            if (r >= 1 && __any_sync(0xffffffffu, r == n)) {        // log_probability (3378-3396); CTA-uniform branch
                prof_bar(n_threads);
                double Pr = 0.0;
                for (int j = 0; j < S.end_n; ++j) {
                    const ProfEndRec E = ldg_rec(S.end_list + j);
                    double v;
                    if (E.cls == PROF_M) v = ldb(cb, prof_slot<P>(E.k, L::MX));
                    else if (E.cls == PROF_DROP) v = ldb(cb, prof_slot<P>(E.k, L::MX)) * M.rec[E.k * kProfRecDoubles + PR_WD];
                    else if (E.cls == PROF_SKIP) v = ldb(cb, prof_slot<P>(E.k, L::SK));
                    else v = ldb(cb, prof_slot<P>(E.k, L::BK));
                    Pr = fma(v, E.w, Pr);
                }
                if (r == n) {
                    double lp;
                    if (Pr == 0.0) lp = neg_inf();
                    else if (!(Pr < __longlong_as_double(0x7FF0000000000000LL))) lp = __longlong_as_double(0x7FF8000000000000LL);
                    else lp = log(Pr) - static_cast<double>(scale_total) * 0.693147180559945309417;
                    O.logp[ev] = lp;
                }
                prof_bar(n_threads);
            }
            scale_total += shift;
            if (r == nmax) break;
            if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
            prof_bar(n_threads);
        }
        return;
    }

    double vM[P], vB[P], vI[P], eM[P];
#pragma unroll
    for (int p = 0; p < P; ++p) vM[p] = vB[p] = vI[p] = 0.0;
    const int pw = warp - 1;
    unsigned char* const tb = cb + (pw + 1) * L::WS;
    const double* const recw = M.rec + pw * P * kProfRecDoubles;
    for (int r = 0;; ++r) {
        const int q = r & 1;
        int mx = 0;
        if (r < nmax) {
            // the whole of row r + 1 is multiplied by 2^shift, chosen from the largest emitting value of row r
            const int shift = prof_row_shift(M.smax[q * kLanes + lane]);
            const double left = ldb(tb, -L::WS + L::MX + (P - 1) * L::SLOT);
            const double right = ldb(tb, L::WS + L::MX);
            const double x = M.xt[(r & (TC - 1)) * (kLanes + 1) + lane];
            const double scale2 = __hiloint2double((shift + 1023) << 20, 0);
            mx = prof_fwd_early<P>(recw, tb, left, right, x, static_cast<double>(shift), scale2, vM, vB, vI, eM);
        }
        prof_bar(n_threads);
        const bool ends = r >= 1 && __any_sync(0xffffffffu, r == n);
        double skin[P], bkin[P];
        prof_fwd_fill<P>(recw, tb, ends, skin, bkin);
        if (ends) {
            prof_bar(n_threads);
            prof_bar(n_threads);
        }
        if (r == nmax) break;
        mx = prof_fwd_late<P>(recw, tb, mx, vM, eM, skin, bkin);
        atomicMax(&M.smax[(q ^ 1) * kLanes + lane], mx);
        if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(M.xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
        prof_bar(n_threads);
    }
}

}  // namespace pp
