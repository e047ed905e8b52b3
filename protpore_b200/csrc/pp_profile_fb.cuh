// pp_profile_fb.cuh -- forward-backward (Baum-Welch E-step) of a LINEAR PROFILE HMM for sm_100a, register-resident.
//
// Model._forward_backward / _train_once_baum_welch (yahmm.pyx:3193-3364, 4107-4236) in sufficient-statistic form, for
// the models pp_profile.h recognises.  Same mathematics as k_fb_lane (pp_lane_fb.cuh: scaled linear-space sweeps with
// power-of-two row scaling, float64-accurate emissions, the posterior factor c_r = 2^(S_n - S_r - T_r) / P^), different
// machine mapping -- the one of pp_profile_kernels.cuh: lane j = event j, position warp w owns positions [4w, 4w + 4) and
// keeps what belongs to them in registers, the two silent chains cross the CTA as affine block carries propagated by one
// chain warp.  k_fb_lane interprets run records with fully unrolled bodies (46,832 instructions, 750 KB of code: its issue
// slots starve on instruction fetch at large batches); this kernel is two loops.
//
//   phase F  the forward sweep of k_profile_forward with float64-accurate 2^t, row maxima steered to 2^480; every row's
//            values f^_r (M, I_INS, I_BLIP, S_SKIP, B_BACK per position; S_DROPOFF = M * t is recomputed) and S_r go to a
//            per-CTA scratch in HBM ([row][warp][5 P slots][lane], 256-byte coalesced stores).
//   phase B  rows n .. 0.  H_l = p_l(x_{r+1}) 2^shift b^_{r+1}[l] of the warp's emitting states stays in registers; the
//            backward chains
//                S_SKIP_k = [t H(M_{k+1}) + t_end 1(r = n)] + t S_SKIP_{k+1}      (descending)
//                B_BACK_k = [t H(M_{k-1}) + t_end 1(r = n)] + t B_BACK_{k-1}      (ascending)
//            depend only on H, which is known when the row starts: each warp publishes the affine summary of its block,
//            the chain warp propagates the carries WHILE the position warps compute every term that does not read a
//            chain value (`early`), and after one barrier the warps fill in their links and finish M (`late`).
//            Every product w * value that makes up b^_r[k] is also an expected-count term: count(k -> l) += (f^_r[k] c_r)
//            (w value); per emitting state (1, x, x^2) gamma.  The 30 per-lane (= per-event) terms of a position are summed
//            over the 32 lanes eight at a time through a shared-memory transpose (8 stores, 4 16-byte loads, 7 adds), and
//            the result is added to REGISTER accumulators that live as long as the CTA: no read-modify-write in L2 per row.
// Events the scaled sweep cannot resolve (P^ = 0 / inf / NaN) contribute nothing and go to the exact log-space kernel; a
// scale factor outside the double range or a failed final check (posterior of the start state at row 0 = 1) hands the
// whole batch to it (pp_aux.cu), exactly as with k_fb_lane.
#pragma once
#include "pp_profile_kernels.cuh"

namespace pp {

constexpr int kPfbTargetExp = 480;       // row maxima of both sweeps are steered to ~2^480 (c_r ~ 2^-960 stays normal)
constexpr int kPfbT2 = 17;               // 2^(j/16), j = -8 .. 8
constexpr int kPfbStageRow = 40;         // doubles per staging row: 32 lanes + 8 (rows 16 banks apart: conflict-free 16-byte loads)
constexpr int kPfbStageDoubles = 8 * kPfbStageRow;          // per warp: one buffer of 8 terms


// backward-phase slots of a warp block (phase F uses the ProfBlock layout in the same memory)
template <int P>
struct PfbBlock {
    static constexpr int SLOT = kSlotBytes;
    static constexpr int WS = ProfBlock<P>::WS;
    static constexpr int HMF = 0;                // [2] H of the block's first match state, by row parity
    static constexpr int HML = 2 * SLOT;         // [2] ... of its last
    static constexpr int HIF = 4 * SLOT;         // [2] H of the block's first insert state
    static constexpr int EDN = 6 * SLOT;         // skip chain: the block's own entries, carried to its lower end
    static constexpr int TDN = 7 * SLOT;         //             the entry that comes from the right neighbour's first match state
    static constexpr int EUP = 8 * SLOT;         // back-slip chain
    static constexpr int TUP = 9 * SLOT;
    static constexpr int VIN = 10 * SLOT;        // S_SKIP_{k0+P}(r): enters the block from above
    static constexpr int UIN = 11 * SLOT;        // B_BACK_{k0-1}(r): enters from below
    static_assert(12 * SLOT <= WS, "backward slots fit the forward block");
};

__host__ __device__ inline size_t pfb_smem_bytes(int Kr, int NW, int P) {
    // forward records | backward records | NW + 2 warp blocks | observation tile | staging | row maxima, 1/P^, S_n, c_r, 2^(j/16)
    return static_cast<size_t>(Kr) * (kProfRecDoubles + kFbRecDoubles) * 8 + static_cast<size_t>(NW + 2) * (5 * P + 5) * kSlotBytes +
           static_cast<size_t>(kLaneTileCols) * (kLanes + 1) * 8 + static_cast<size_t>(NW) * kPfbStageDoubles * 8 +
           2 * kLanes * 4 + 4 * kLanes * 8 + kPfbT2 * 8 + 64;
}

__device__ __forceinline__ int pfb_row_shift(int hi) {
    const int ex = (hi >> 20) & 0x7ff;
    const int shift = (ex == 0 || ex == 0x7ff) ? 0 : (kPfbTargetExp + 1023 - ex);
    return max(-900, min(900, shift));
}

// Backward sweep: H(r - 1) is made from b^_r while only the maxima of b^_{r+1} are known (those of b^_r are being collected),
// so the shift allows for the one already applied in between: b^_r ~ b^_{r+1} 2^prev.  (Steering by the stale maximum
// alone applies every correction twice and oscillates between 2^0 and 2^960.)
// `first`: row r is the event's last row (b^_n is the end weights, ~2^-4, and nothing is known yet): the first shift lifts
// H(n - 1) to the target at once -- an outlier in the last observation (2^-1200) would otherwise flush the whole sweep.
__device__ __forceinline__ int pfb_row_shift_lagged(int hi, int prev, bool first) {
    const int ex = (hi >> 20) & 0x7ff;
    const int shift = (ex == 0 || ex == 0x7ff) ? (first ? kPfbTargetExp : 0) : (kPfbTargetExp + 1023 - ex - prev);
    return max(-900, min(900, shift));
}

// 2^t, float64-accurate (rel. error ~6e-16): t = n + j/16 + g, |g| <= 1/32; 2^g by a degree-6 polynomial of exp(g ln 2),
// 2^(j/16) from shared memory, 2^n into the exponent.  An exact 0 below 2^-1000 (and for -inf / NaN).
__device__ __forceinline__ double pfb_exp2(double t, const double* __restrict__ t2) {
    const double magic = 6755399441055744.0;
    const double tm = t + magic;
    const int n = __double2loint(tm);
    const double fr = t - (tm - magic);                        // [-0.5, 0.5]
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
    const bool ok = t > -1000.0;                                // false for NaN as well
    const double tv = t2[ok ? j + 8 : 8];
    const double scale = __hiloint2double((n + 1023) << 20, 0);
    const double r = (p * tv) * scale;
    return ok ? r : 0.0;
}
// linear-space emission factors p(x) 2^shift from a 4-double record (pp_profile.h)
__device__ __forceinline__ double pfb_lin_normal(const double* R, double x, double shiftd, const double* t2) {
    const double2 a = ld2(R);                                   // {mu, log2(e) / (2 sigma^2)}
    const double d = x - a.x;
    return pfb_exp2(fma(-(d * d), a.y, R[2]) + shiftd, t2);
}
__device__ __forceinline__ double pfb_emit_ins(const double* R, int kinds, double x, double shiftd, double scale2, const double* t2) {
    return ((kinds >> 8) & 0xff) == EMIS_POINT ? prof_lin_point(R, x, scale2) : pfb_lin_normal(R, x, shiftd, t2);
}

__device__ __forceinline__ double pfb_ldcg(const void* p) { return __ldcg(static_cast<const double*>(p)); }
__device__ __forceinline__ void pfb_stcg(void* p, double v) { __stcg(static_cast<double*>(p), v); }
__device__ __forceinline__ double pfb_load_obs(const LaneBatchAny& B, int64_t i) {
    return B.obs_f64 ? __ldg(static_cast<const double*>(B.obs) + i)
                     : static_cast<double>(__ldg(static_cast<const float*>(B.obs) + i));
}
__device__ __forceinline__ void pfb_atomic_min(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(static_cast<long long>(old)) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void pfb_atomic_max(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(static_cast<long long>(old)) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}

// Sum of 8 per-lane terms over the 32 lanes through a shared-memory transpose: lane L returns the sum of term L >> 2
// over lanes {2q, 2q+1, 2q+8, 2q+9, 2q+16, 2q+17, 2q+24, 2q+25}, q = L & 3 (the four lanes of a term hold its four partial
// sums).  `stage` = the warp's staging buffer; the leading __syncwarp orders the previous call's loads before these stores.
__device__ __forceinline__ double pfb_reduce8(const double (&t)[8], double* stage, int lane) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) stage[i * kPfbStageRow + lane] = t[i];
    __syncwarp();
    const double* rp = stage + (lane >> 2) * kPfbStageRow + (lane & 3) * 2;
    const double2 a = ld2(rp), b = ld2(rp + 8), c = ld2(rp + 16), d = ld2(rp + 24);
    return ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
}

// log of a posterior (f^ b^) c given log c: the product of the two large factors is formed first; where it is not a normal
// double although both factors are non-zero the logarithms are added instead
__device__ __forceinline__ double pfb_log_post(double bf, double b, double f, double lc) {
    if (bf >= 2.2250738585072014e-308) return log(bf) + lc;
    if (b > 0.0 && f > 0.0) return (log(b) + log(f)) + lc;
    return neg_inf();
}

// ---- phase F ------------------------------------------------------------------------------------------------------
// prof_fwd_early with float64-accurate exponentials
template <int P>
__device__ __forceinline__ int pfb_fwd_early(const double* __restrict__ recw, const unsigned char* __restrict__ tb, double left,
                                             double right, double x, double shiftd, double scale2, const double* t2,
                                             double (&vM)[P], double (&vB)[P], double (&vI)[P], double (&eMv)[P]) {
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
        eMv[p] = pfb_lin_normal(R + PR_EM, x, shiftd, t2);
        const double eB = prof_lin_uniform(R + PR_EB, x, scale2);
        const double eI = pfb_emit_ins(R + PR_EI, kinds, x, shiftd, scale2, t2);
        const double drop = vM[p] * wdk.x;
        double a0 = mprev * w01.x;
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
// prof_fwd_fill that also writes the chain values of row r to the scratch row (gs: this lane's S_SKIP slot of position k0)
template <int P>
__device__ __forceinline__ void pfb_fwd_fill(const double* __restrict__ recw, unsigned char* __restrict__ tb, bool keep,
                                             unsigned char* __restrict__ gs, double (&skin)[P], double (&bkin)[P]) {
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
        pfb_stcg(gs + j * kSlotBytes, v);
        pfb_stcg(gs + (P + jd) * kSlotBytes, u);
        if (keep) {
            stb(tb, L::SK + j * L::SLOT, v);
            stb(tb, L::BK + jd * L::SLOT, u);
        }
    }
}

// ---- phase B ------------------------------------------------------------------------------------------------------
// Publishes what the next row needs from this warp: the boundary H values (parity `q`), the affine summaries of its two
// chain blocks, and the entries its boundary match states contribute to the neighbouring blocks.  ind = 1(row == n).
template <int P>
__device__ __forceinline__ void pfb_publish(const double* __restrict__ recb, unsigned char* __restrict__ tb, int q, double ind,
                                            const double (&HM)[P], const double (&HI)[P]) {
    using Bk = PfbBlock<P>;
    stb(tb, Bk::HMF + q * Bk::SLOT, HM[0]);
    stb(tb, Bk::HML + q * Bk::SLOT, HM[P - 1]);
    stb(tb, Bk::HIF + q * Bk::SLOT, HI[0]);
    // skip chain, descending: S_SKIP_{k0+j} = a_j + c_j S_SKIP_{k0+j+1}; the block's way out is S_SKIP_{k0}
    double e = recb[(P - 1) * kFbRecDoubles + FR_WE + 2] * ind;        // a_{P-1} without the neighbour's H
#pragma unroll
    for (int j = P - 2; j >= 0; --j) {
        const double* R = recb + j * kFbRecDoubles;
        const double2 ws = ld2(R + FR_WS);
        e = fma(ws.y, e, fma(ws.x, HM[j + 1], R[FR_WE + 2] * ind));
    }
    stb(tb, Bk::EDN, e);
    stb(tb, -Bk::WS + Bk::TDN, recb[FR_TDN] * HM[0]);
    // back-slip chain, ascending: B_BACK_{k0+j} = a'_j + c'_j B_BACK_{k0+j-1}; the way out is B_BACK_{k0+P-1}
    double eu = recb[FR_WE + 3] * ind;
#pragma unroll
    for (int j = 1; j < P; ++j) {
        const double* R = recb + j * kFbRecDoubles;
        const double2 wk = ld2(R + FR_WK);
        eu = fma(wk.y, eu, fma(wk.x, HM[j - 1], R[FR_WE + 3] * ind));
    }
    stb(tb, Bk::EUP, eu);
    stb(tb, Bk::WS + Bk::TUP, recb[FR_TUP] * HM[P - 1]);
}

template <int P>
__device__ __forceinline__ void pfb_bwd_carries(const double* __restrict__ recb, unsigned char* __restrict__ cb, int NW) {
    using Bk = PfbBlock<P>;
    double v = 0.0, u = 0.0;
#pragma unroll 2
    for (int i = 0; i < NW; ++i) {
        unsigned char* bd = cb + (NW - i) * Bk::WS;                 // skip chain: blocks NW-1 .. 0
        unsigned char* bu = cb + (i + 1) * Bk::WS;                  // back-slip chain: blocks 0 .. NW-1
        const double ed = ldb(bd, Bk::EDN) + ldb(bd, Bk::TDN), eu = ldb(bu, Bk::EUP) + ldb(bu, Bk::TUP);
        const double cd = recb[(NW - 1 - i) * P * kFbRecDoubles + FR_CDN], cu = recb[i * P * kFbRecDoubles + FR_CUP];
        stb(bd, Bk::VIN, v);
        stb(bu, Bk::UIN, u);
        v = fma(cd, v, ed);
        u = fma(cu, u, eu);
    }
}

template <int P, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_profile_fb(const ProfFbSched S, const LaneBatchAny B) {
    using L = ProfBlock<P>;
    using Bk = PfbBlock<P>;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TC = kLaneTileCols;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int n_warps = S.NW + 1, n_threads = n_warps * 32;

    double* const s_recf = reinterpret_cast<double*>(smem);
    double* const s_recb = s_recf + static_cast<size_t>(S.Kr) * kProfRecDoubles;
    unsigned char* const s_blocks = reinterpret_cast<unsigned char*>(s_recb + static_cast<size_t>(S.Kr) * kFbRecDoubles);
    const int n_col = (S.NW + 2) * (5 * P + 5) * kLanes;
    double* const xt = reinterpret_cast<double*>(s_blocks) + n_col;
    double* const s_stage = xt + TC * (kLanes + 1);
    double* const s_invp = s_stage + static_cast<size_t>(S.NW) * kPfbStageDoubles;
    double* const s_sn = s_invp + kLanes;
    double* const s_cr = s_sn + kLanes;                          // [2][32]
    double* const s_t2 = s_cr + 2 * kLanes;
    int32_t* const smax = reinterpret_cast<int32_t*>(s_t2 + kPfbT2 + 1);
    __shared__ int s_group;

    for (int i = threadIdx.x; i < S.Kr * kProfRecDoubles; i += n_threads) s_recf[i] = __ldg(S.rec_f + i);
    for (int i = threadIdx.x; i < S.Kr * kFbRecDoubles; i += n_threads) s_recb[i] = __ldg(S.rec_b + i);
    for (int i = threadIdx.x; i < kPfbT2; i += n_threads) s_t2[i] = exp2(static_cast<double>(i - 8) * 0.0625);

    unsigned char* const cb = s_blocks + lane * 8;
    const int pw = warp - 1;                                     // position warp index (-1: the chain warp)
    unsigned char* const tb = cb + (pw + 1) * L::WS;
    const double* const recw = s_recf + pw * P * kProfRecDoubles;
    const double* const recb = s_recb + pw * P * kFbRecDoubles;
    double* const stage = s_stage + (pw < 0 ? 0 : pw) * kPfbStageDoubles;
    const size_t row_bytes = static_cast<size_t>(S.row_slots) * kSlotBytes;
    unsigned char* const scratch = S.scratch + static_cast<size_t>(blockIdx.x) * S.scratch_stride + lane * 8;
    const uint32_t scale_off = static_cast<uint32_t>(S.NW * 5 * P) * kSlotBytes;
    const uint32_t my_off = static_cast<uint32_t>((pw < 0 ? 0 : pw) * 5 * P) * kSlotBytes;

    double acc[P][kFbChunks];                                    // statistics of the warp's positions, all groups of this CTA
    double acc_start = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < kFbChunks; ++c) acc[p][c] = 0.0;

    for (;;) {
        __syncthreads();                                         // previous group is done with shared memory
        if (threadIdx.x == 0) s_group = atomicAdd(S.counter, 1);
        __syncthreads();
        if (s_group >= S.n_groups) break;
        const int g = s_group;
        const int ev = B.order[g * kLanes + lane];
        const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
        const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
        const int nmax = B.group_len[g];

        // ================= phase F =================
        for (int i = threadIdx.x; i < n_col; i += n_threads) reinterpret_cast<double*>(s_blocks)[i] = 0.0;
        if (threadIdx.x < 2 * kLanes) smax[threadIdx.x] = 0;
        if (warp == 0) { s_invp[lane] = 0.0; s_sn[lane] = 0.0; }
        if (nmax > 0) prof_fill_tile(xt, B.obs, B.obs_f64 != 0, off, n, 0, warp, lane, n_warps);
        __syncthreads();

        double xlo = __longlong_as_double(0x7FF0000000000000LL), xhi = neg_inf();
        if (warp == 0) {                                         // the chain warp
            long long scale_total = 0;                           // S_r
            for (int r = 0;; ++r) {
                const int q = r & 1;
                const int shift = pfb_row_shift(smax[q * kLanes + lane]);
                pfb_stcg(scratch + static_cast<size_t>(r) * row_bytes + scale_off, static_cast<double>(scale_total));
                prof_fwd_carries<P>(s_recf, cb, S.NW, r == 0 ? 1.0 : 0.0);
                smax[(q ^ 1) * kLanes + lane] = 0;
                prof_bar(n_threads);
                if (r >= 1 && __any_sync(0xffffffffu, r == n)) {      // P^ = sum over the end state's in-edges
                    prof_bar(n_threads);
                    double Pr = 0.0;
                    for (int j = 0; j < S.end_n; ++j) {
                        const ProfEndRec E = ldg_rec(S.end_list + j);
                        double v;
                        if (E.cls == PROF_M) v = ldb(cb, prof_slot<P>(E.k, L::MX));
                        else if (E.cls == PROF_DROP) v = ldb(cb, prof_slot<P>(E.k, L::MX)) * s_recf[E.k * kProfRecDoubles + PR_WD];
                        else if (E.cls == PROF_SKIP) v = ldb(cb, prof_slot<P>(E.k, L::SK));
                        else v = ldb(cb, prof_slot<P>(E.k, L::BK));
                        Pr = fma(v, E.w, Pr);
                    }
                    if (r == n) {
                        // P^ hundreds of orders below the level the rows are steered to (outliers in the last rows): 1 / P^
                        // must stay a double with room for the row factors.  Such an event is flagged (NaN) like one with
                        // P^ = inf and goes to the exact kernel, alone
                        // (2^-900; when posteriors are asked for, 2^-120: in such an event the supports of the two sweeps are
                        // nearly disjoint in the last rows and posteriors far above e^-250 would be lost)
                        const double floor_p = S.post != nullptr ? 7.52316384526264e-37 : 1.1832913578315177e-271;
                        const bool usable = Pr > floor_p && Pr < __longlong_as_double(0x7FF0000000000000LL);
                        double lp;
                        if (Pr == 0.0) lp = neg_inf();
                        else if (!usable) lp = __longlong_as_double(0x7FF8000000000000LL);
                        else lp = log(Pr) - static_cast<double>(scale_total) * 0.693147180559945309417;
                        S.logp[ev] = lp;
                        s_invp[lane] = usable ? 1.0 / Pr : 0.0;
                        s_sn[lane] = static_cast<double>(scale_total);
                    }
                    prof_bar(n_threads);
                }
                scale_total += shift;
                if (r == nmax) break;
                if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
                prof_bar(n_threads);
            }
        } else {
            double vM[P], vB[P], vI[P], eM[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                vM[p] = vB[p] = vI[p] = 0.0;
                pfb_stcg(scratch + my_off + p * kSlotBytes, 0.0);            // f^_0 of the emitting states
                pfb_stcg(scratch + my_off + (P + p) * kSlotBytes, 0.0);
                pfb_stcg(scratch + my_off + (2 * P + p) * kSlotBytes, 0.0);
            }
            for (int r = 0;; ++r) {
                const int q = r & 1;
                int mx = 0;
                if (r < nmax) {
                    const int shift = pfb_row_shift(smax[q * kLanes + lane]);
                    const double left = ldb(tb, -L::WS + L::MX + (P - 1) * L::SLOT);
                    const double right = ldb(tb, L::WS + L::MX);
                    const double x = xt[(r & (TC - 1)) * (kLanes + 1) + lane];
                    if (r < n) { xlo = fmin(xlo, x); xhi = fmax(xhi, x); }
                    const double scale2 = __hiloint2double((shift + 1023) << 20, 0);
                    mx = pfb_fwd_early<P>(recw, tb, left, right, x, static_cast<double>(shift), scale2, s_t2, vM, vB, vI, eM);
                }
                prof_bar(n_threads);
                const bool ends = r >= 1 && __any_sync(0xffffffffu, r == n);
                double skin[P], bkin[P];
                unsigned char* grow = scratch + static_cast<size_t>(r) * row_bytes + my_off;
                pfb_fwd_fill<P>(recw, tb, ends, grow + 3 * P * kSlotBytes, skin, bkin);
                if (ends) {
                    prof_bar(n_threads);
                    prof_bar(n_threads);
                }
                if (r == nmax) break;
                mx = prof_fwd_late<P>(recw, tb, mx, vM, eM, skin, bkin);
                grow += row_bytes;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    pfb_stcg(grow + p * kSlotBytes, vM[p]);
                    pfb_stcg(grow + (P + p) * kSlotBytes, vI[p]);
                    pfb_stcg(grow + (2 * P + p) * kSlotBytes, vB[p]);
                }
                atomicMax(&smax[(q ^ 1) * kLanes + lane], mx);
                if (r + 1 < nmax && ((r + 1) & (TC - 1)) == 0) prof_fill_tile(xt, B.obs, B.obs_f64 != 0, off, n, r + 1, warp, lane, n_warps);
                prof_bar(n_threads);
            }
        }
        __syncthreads();

        // ================= phase B =================
        for (int i = threadIdx.x; i < n_col; i += n_threads) reinterpret_cast<double*>(s_blocks)[i] = 0.0;
        if (threadIdx.x < 2 * kLanes) smax[threadIdx.x] = 0;
        __syncthreads();
        const double invp = s_invp[lane], sn = s_sn[lane];
        const bool valid = ev >= 0 && invp > 0.0;
        bool failed = false;

        if (warp == 0) {
            double T = 0.0;                                      // T_r: sum of the backward shifts of rows nmax-1 .. r
            // c_r = 2^(S_n - S_r - T_r) / P^, in two halves: the exponent alone may leave the double range although the
            // product does not (a tiny P^ after an outlier in the last rows)
            auto c_of = [&](int r, double Tr) -> double {
                const double sr = pfb_ldcg(scratch + static_cast<size_t>(r) * row_bytes + scale_off);
                const double ed = sn - sr - Tr;
                const bool active = valid && r <= n;
                const int ei = static_cast<int>(fmin(fmax(ed, -2000.0), 2000.0));
                const int e1 = ei / 2, e2 = ei - e1;
                const double cc = (invp * __hiloint2double((e1 + 1023) << 20, 0)) * __hiloint2double((e2 + 1023) << 20, 0);
                const bool in_range = cc > 1e-300 && cc < 1e300;
                if (active && !in_range) {
                    if (!failed && atomicCAS(S.fail + 1, 0, 1) == 0) {
                        S.fail[2] = ev; S.fail[3] = r;
                        reinterpret_cast<double*>(S.fail + 7)[0] = ed;
                        reinterpret_cast<double*>(S.fail + 7)[1] = invp;
                        reinterpret_cast<double*>(S.fail + 7)[2] = sn;
                        reinterpret_cast<double*>(S.fail + 7)[3] = Tr + 1e-3 * n + 1e-6 * nmax;
                    }
                    failed = true;
                }
                return (active && in_range) ? cc : 0.0;
            };
            s_cr[(nmax & 1) * kLanes + lane] = c_of(nmax, 0.0);
            int sh_prev = 0;
            for (int r = nmax; r >= 0; --r) {
                const int q = r & 1;
                prof_bar(n_threads);                             // A: the summaries of row r are published
                pfb_bwd_carries<P>(s_recb, cb, S.NW);
                if (r > 0) {
                    sh_prev = pfb_row_shift_lagged(smax[(q ^ 1) * kLanes + lane], sh_prev, r == n);   // shift of H(r - 1): from the maxima of b^_{r+1}
                    T += static_cast<double>(sh_prev);
                    s_cr[(q ^ 1) * kLanes + lane] = c_of(r - 1, T);
                }
                smax[q * kLanes + lane] = 0;                     // row r collects its maxima here in the late part
                prof_bar(n_threads);                             // B
            }
            if (__any_sync(0xffffffffu, failed)) {
                const uint32_t b = __ballot_sync(0xffffffffu, failed);
                if (lane == 0) atomicAdd(S.fail, __popc(b));
            }
            continue;
        }

        // ---- position warps ----
        double HM[P], HI[P], HB[P];
        uint32_t flg[P][3];
#pragma unroll
        for (int p = 0; p < P; ++p) { HM[p] = HI[p] = HB[p] = 0.0; flg[p][0] = flg[p][1] = flg[p][2] = 0u; }
        pfb_publish<P>(recb, tb, nmax & 1, (nmax == n && valid) ? 1.0 : 0.0, HM, HI);
        // xb = obs[r - 1]: the observation the posteriors of row r weight, and the one H(r - 1) is made with
        double xb = (nmax >= 1 && nmax - 1 < n) ? pfb_load_obs(B, off + nmax - 1) : 0.0;
        double skip0 = 0.0;
        int sh_prev = 0;
        for (int r = nmax; r >= 0; --r) {
            const int q = r & 1;
            double xpre = 0.0;
            if (r - 2 >= 0 && r - 2 < n) xpre = pfb_load_obs(B, off + r - 2);
            const unsigned char* srow = scratch + static_cast<size_t>(r) * row_bytes + my_off;
            if (r > 0) {                                         // pull row r - 1 of the scratch into L2 while row r is processed
                const unsigned char* nrow = srow - row_bytes - lane * 8;
                for (int o = lane * 128; o < 5 * P * static_cast<int>(kSlotBytes); o += 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + o));
            }
            prof_bar(n_threads);                                 // A
            const double c = s_cr[q * kLanes + lane];
            const double ind = (r == n && valid) ? 1.0 : 0.0;
            const double hmR = ldb(tb, Bk::WS + Bk::HMF + q * Bk::SLOT);
            const double hiR = ldb(tb, Bk::WS + Bk::HIF + q * Bk::SLOT);
            const double hmL = ldb(tb, -Bk::WS + Bk::HML + q * Bk::SLOT);
            const double xx = xb * xb;
            // posteriors of row r belong to observation r - 1 (only asked for by forward_backward(): a per-lane scattered store)
            const bool writes = r >= 1 && r <= n && c > 0.0;
            const double lc = (S.post != nullptr && writes) ? log(c) : 0.0;
            const int64_t prow = (off + r - 1) * S.ne;
            double bMp[P], bI[P], bB[P], aS[P], aK[P], fMv[P], fSv[P], fKv[P];
            int mx = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double* R = recb + p * kFbRecDoubles;
                const double fM = pfb_ldcg(srow + p * kSlotBytes), fI = pfb_ldcg(srow + (P + p) * kSlotBytes);
                const double fB = pfb_ldcg(srow + (2 * P + p) * kSlotBytes), fS = pfb_ldcg(srow + (3 * P + p) * kSlotBytes);
                const double fK = pfb_ldcg(srow + (4 * P + p) * kSlotBytes);
                const double2 w01 = ld2(R + FR_WM), w23 = ld2(R + FR_WM + 2), w45 = ld2(R + FR_WM + 4);
                const double2 we01 = ld2(R + FR_WE), we23 = ld2(R + FR_WE + 2);
                const double2 ws = ld2(R + FR_WS), wk = ld2(R + FR_WK), wd = ld2(R + FR_WD), wi = ld2(R + FR_WI), wb = ld2(R + FR_WB);
                const double hm_up = p < P - 1 ? HM[p < P - 1 ? p + 1 : 0] : hmR;
                const double hi_up = p < P - 1 ? HI[p < P - 1 ? p + 1 : 0] : hiR;
                const double hm_dn = p > 0 ? HM[p > 0 ? p - 1 : 0] : hmL;
                const double p0 = w01.x * hm_up, p1 = w01.y * HM[p], p2 = w23.x * hi_up, p3 = w23.y * HB[p];
                const double p7 = wd.x * HM[p], p8 = we01.y * ind;
                const double p4 = w45.x * (p7 + p8), p5 = w45.y * hm_dn, p6 = we01.x * ind;
                const double p9 = ws.x * hm_up, p10 = we23.x * ind, p11 = wk.x * hm_dn, p12 = we23.y * ind;
                const double p13 = wi.x * HM[p], p14 = wi.y * HI[p], p15 = wb.x * HM[p], p16 = wb.y * HB[p];
                bMp[p] = ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + p6);
                bI[p] = p13 + p14;
                bB[p] = p15 + p16;
                aS[p] = p9 + p10;
                aK[p] = p11 + p12;
                fMv[p] = fM; fSv[p] = fS; fKv[p] = fK;
                const double fD = fM * w45.x;                     // f^ of S_DROPOFF_k, as the forward sweep computed it
                // a term is (w value) f^ c: the two large factors first (<= 2^962), so that a state with a forward weight
                // of 1e-200 and an ordinary backward weight is not lost to f^ c leaving the double range (pp_lane_fb.cuh)
                {
                    const double t[8] = {(p0 * fM) * c, (p1 * fM) * c, (p2 * fM) * c, (p3 * fM) * c,
                                         (p4 * fM) * c, (p5 * fM) * c, (p6 * fM) * c, (p7 * fD) * c};
                    acc[p][0] += pfb_reduce8(t, stage, lane);
                }
                {
                    const double t[8] = {(p8 * fD) * c, (p9 * fS) * c, (p10 * fS) * c, (p11 * fK) * c,
                                         (p12 * fK) * c, (p13 * fI) * c, (p14 * fI) * c, (p15 * fB) * c};
                    acc[p][1] += pfb_reduce8(t, stage, lane);
                }
                {
                    const double gi = (bI[p] * fI) * c, gb = (bB[p] * fB) * c;
                    if (S.post != nullptr && writes) {
                        const int sI = __ldg(S.emit_state + (pw * P + p) * 3 + 1), sB = __ldg(S.emit_state + (pw * P + p) * 3 + 2);
                        if (sI >= 0) S.post[prow + sI] = pfb_log_post(bI[p] * fI, bI[p], fI, lc);
                        if (sB >= 0) S.post[prow + sB] = pfb_log_post(bB[p] * fB, bB[p], fB, lc);
                    }
                    const double t[8] = {(p16 * fB) * c, gi, gi * xb, gi * xx, gb, gb * xb, gb * xx, 0.0};
                    acc[p][2] += pfb_reduce8(t, stage, lane);
                    flg[p][1] |= __ballot_sync(0xffffffffu, gi > 0.0);
                    flg[p][2] |= __ballot_sync(0xffffffffu, gb > 0.0);
                }
                mx = max(mx, max(__double2hiint(bI[p]), __double2hiint(bB[p])));
            }
            prof_bar(n_threads);                                 // B: the carries of row r are published
            double sk_up[P], bk_dn[P], bk[P];
            {
                double v = ldb(tb, Bk::VIN), u = ldb(tb, Bk::UIN);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const int jd = P - 1 - j;
                    sk_up[jd] = v;                               // S_SKIP_{k0+jd+1}(r)
                    bk_dn[j] = u;                                // B_BACK_{k0+j-1}(r)
                    v = fma(recb[jd * kFbRecDoubles + FR_WS + 1], v, aS[jd]);
                    u = fma(recb[j * kFbRecDoubles + FR_WK + 1], u, aK[j]);
                    bk[j] = u;
                }
                skip0 = v;                                       // S_SKIP_{k0}(r)
            }
            double bM[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double* R = recb + p * kFbRecDoubles;
                const double2 w67 = ld2(R + FR_WM + 6);
                const double p24 = w67.x * sk_up[p], p25 = w67.y * bk[p];
                const double p26 = R[FR_WS + 1] * sk_up[p], p27 = R[FR_WK + 1] * bk_dn[p];
                bM[p] = bMp[p] + (p24 + p25);
                const double gm = (bM[p] * fMv[p]) * c;
                if (S.post != nullptr && writes) {
                    const int sM = __ldg(S.emit_state + (pw * P + p) * 3);
                    if (sM >= 0) S.post[prow + sM] = pfb_log_post(bM[p] * fMv[p], bM[p], fMv[p], lc);
                }
                const double t[8] = {(p24 * fMv[p]) * c, (p25 * fMv[p]) * c, (p26 * fSv[p]) * c, (p27 * fKv[p]) * c,
                                     gm, gm * xb, gm * xx, 0.0};
                acc[p][3] += pfb_reduce8(t, stage, lane);
                flg[p][0] |= __ballot_sync(0xffffffffu, gm > 0.0);
                mx = max(mx, __double2hiint(bM[p]));
            }
            atomicMax(&smax[q * kLanes + lane], mx);
            if (r == 0) break;
            // H(r - 1) = p(obs[r - 1]) 2^shift b^_r, shift from the maxima of b^_{r+1}
            {
                const int shift = pfb_row_shift_lagged(smax[(q ^ 1) * kLanes + lane], sh_prev, r == n);
                sh_prev = shift;
                const double shiftd = static_cast<double>(shift);
                const double scale2 = __hiloint2double((shift + 1023) << 20, 0);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const double* R = recb + p * kFbRecDoubles;
                    const int kinds = __double2loint(R[FR_KINDS]);
                    HM[p] = pfb_lin_normal(R + FR_EM, xb, shiftd, s_t2) * bM[p];
                    HI[p] = pfb_emit_ins(R + FR_EI, kinds, xb, shiftd, scale2, s_t2) * bI[p];
                    HB[p] = prof_lin_uniform(R + FR_EB, xb, scale2) * bB[p];
                }
                pfb_publish<P>(recb, tb, q ^ 1, (r - 1 == n && valid) ? 1.0 : 0.0, HM, HI);
            }
            xb = xpre;
        }
        // ---- end of the group: the start state's two out-edges and the final check (warp of position 0) ----
        if (pw == 0) {
            const double c0 = s_cr[lane];                        // row 0
            const double ps0 = S.start_wm * HM[0], ps1 = S.start_ws * skip0;      // f^_0[start] = 1, S_0 = 0
            const double chk = (ps0 + ps1) * c0;
            if (valid && !(fabs(chk - 1.0) < 1e-6)) {
                if (atomicCAS(S.fail + 1, 0, 2) == 0) {
                    S.fail[2] = ev; S.fail[3] = n;
                    reinterpret_cast<double*>(S.fail + 7)[0] = chk;
                    reinterpret_cast<double*>(S.fail + 7)[1] = c0;
                    reinterpret_cast<double*>(S.fail + 7)[2] = ps0;
                    reinterpret_cast<double*>(S.fail + 7)[3] = ps1;
                }
                failed = true;
            }
            const double t[8] = {ps0 * c0, ps1 * c0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            acc_start += pfb_reduce8(t, stage, lane);
            if (__any_sync(0xffffffffu, failed)) {
                const uint32_t b = __ballot_sync(0xffffffffu, failed);
                if (lane == 0) atomicAdd(S.fail, __popc(b));
            }
        }
        // observed range of every event that gave a state non-zero weight (304-326)
#pragma unroll
        for (int p = 0; p < P; ++p) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const uint32_t mask = flg[p][s];
                if (mask == 0u) continue;
                const int state = __ldg(S.emit_state + (pw * P + p) * 3 + s);
                if (state < 0) continue;
                double lo = ((mask >> lane) & 1u) ? xlo : __longlong_as_double(0x7FF0000000000000LL);
                double hi = ((mask >> lane) & 1u) ? xhi : neg_inf();
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                if (lane == 0) {
                    pfb_atomic_min(S.minmax + 2 * state, lo);
                    pfb_atomic_max(S.minmax + 2 * state + 1, hi);
                }
            }
        }
    }
    // the CTA's statistics: slot (position, chunk), lane = 4 * term + partial
    if (warp > 0) {
        double* a = S.acc + static_cast<size_t>(blockIdx.x) * S.n_acc_slots * kLanes + lane;
        const int pw0 = warp - 1;
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < kFbChunks; ++c) a[static_cast<size_t>((pw0 * P + p) * kFbChunks + c) * kLanes] = acc[p][c];
        if (pw0 == 0) a[static_cast<size_t>(S.Kp) * kFbChunks * kLanes] = acc_start;
    }
}

}  // namespace pp
