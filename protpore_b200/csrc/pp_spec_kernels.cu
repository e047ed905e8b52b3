// pp_spec_kernels.cu -- lane-per-event Viterbi / forward kernels compiled against ONE model's schedule.
//
// Same algorithm, layout and arithmetic as pp_lane_kernels.cuh (read that header first); the difference
// is that the schedule is compile-time structure (pp_codegen.cpp writes PP_SPEC_HEADER):
//   * the per-warp item/edge programs are X-macros expanded into straight-line code, so there are no
//     record loads, no loops and no address arithmetic -- a column slot is `base + immediate`;
//   * transition weights and emission parameters are operands from the __constant__ table PP_WTAB
//     (`DADD R, R, c[3][imm]`), refreshed by pp_spec_set_params after a Baum-Welch M-step;
//   * a silent chain link reads its predecessor from a register (EF) instead of shared memory;
//   * traceback ordinals are packed 8 (or 4) to a 32-bit word and stored with one STG per word.
// Compiled by pp_spec.cpp:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared
//                           -DPP_SPEC_HEADER='"<cache>/pp_spec_<hash>.h"' pp_spec_kernels.cu
#include PP_SPEC_HEADER

#include "pp_device_common.cuh"

#ifndef PP_CTAS_PER_SM
#define PP_CTAS_PER_SM 2
#endif

using namespace pp;

__constant__ double PP_WTAB[PP_N_WCONST];

#define PP_LD(base, idx) (*reinterpret_cast<const double*>((base) + (idx) * 256))
#define PP_WE(k, j) PP_WTAB[PP_EMIS_BASE + 8 * (k) + (j)]

// ---- exact float64 emission log-density (see emission_log in pp_device_common.cuh) ------------
#define PP_EMIS_N(k)                                                                             \
    __dadd_rn(__dsub_rn(PP_WE(k, 3), pp_quot(__dsub_rn(x, PP_WE(k, 0)), PP_WE(k, 1), PP_WE(k, 2))), PP_WE(k, 4))
#define PP_EMIS_P(k) __dadd_rn((fabs(__dsub_rn(x, PP_WE(k, 0))) < 1E-4) ? 0.0 : neg_inf(), PP_WE(k, 4))
#define PP_EMIS_U(k)                                                                             \
    __dadd_rn((x == PP_WE(k, 0) && x == PP_WE(k, 1)) ? 0.0                                       \
                                                     : ((x >= PP_WE(k, 0) && x <= PP_WE(k, 1)) ? PP_WE(k, 3) : neg_inf()), \
              PP_WE(k, 4))

__device__ __forceinline__ double pp_quot(double d, double b, double rinv) {   // correctly rounded d*d / b
    const double d2 = __dmul_rn(d, d);
    const double q0 = __dmul_rn(d2, rinv);
    const double rem = __fma_rn(-q0, b, d2);
    return __fma_rn(rem, rinv, q0);
}

// =============================================================================================
// Viterbi macro bodies
// =============================================================================================
#define V_OPEN(dstp, pshift)                                                                     \
    {                                                                                            \
        double best = neg_inf();                                                                 \
        unsigned arg = 0;                                                                        \
        double* const dst_ = reinterpret_cast<double*>(dstp);                                    \
        constexpr int ps_ = (pshift);                                                            \
        constexpr bool is_start_ = false;
#define V_IN(k, pshift) V_OPEN(en + (k) * 256, pshift) const double e_ = PP_EMIS_N(k);
#define V_IP(k, pshift) V_OPEN(en + (k) * 256, pshift) const double e_ = PP_EMIS_P(k);
#define V_IU(k, pshift) V_OPEN(en + (k) * 256, pshift) const double e_ = PP_EMIS_U(k);
#define V_IS(j, pshift) V_OPEN(sil + (j) * 256, pshift)
#define V_IT(j, pshift)                                                                          \
    {                                                                                            \
        double best = neg_inf();                                                                 \
        unsigned arg = 0;                                                                        \
        double* const dst_ = reinterpret_cast<double*>(sil + (j) * 256);                         \
        constexpr int ps_ = (pshift);                                                            \
        constexpr bool is_start_ = true;
#define V_CAND(s, ord)                                                                           \
    {                                                                                            \
        const double s_ = (s);                                                                   \
        if (s_ > best) { best = s_; arg = (ord); }                                               \
    }
#define V_EO(ord, k, e) V_CAND(__dadd_rn(__dadd_rn(PP_LD(eo, k), PP_WTAB[PP_LOGP_BASE + (e)]), e_), ord)
#define V_EQ(ord, j, e) V_CAND(__dadd_rn(__dadd_rn(PP_LD(sil, j), PP_WTAB[PP_LOGP_BASE + (e)]), e_), ord)
#define V_EN(ord, k, e) V_CAND(__dadd_rn(PP_LD(en, k), PP_WTAB[PP_LOGP_BASE + (e)]), ord)
#define V_ES(ord, j, e) V_CAND(__dadd_rn(PP_LD(sil, j), PP_WTAB[PP_LOGP_BASE + (e)]), ord)
#define V_EF(ord, e) V_CAND(__dadd_rn(prev_, PP_WTAB[PP_LOGP_BASE + (e)]), ord)
#define V_X()                                                                                    \
        if (ROW0 && is_start_) best = 0.0;                                                       \
        *dst_ = best;                                                                            \
        prev_ = best;                                                                            \
        pw |= arg << ps_;                                                                        \
    }
#define V_PW(word)                                                                               \
    if (active) pr[(word) * 32] = pw;                                                            \
    pw = 0;

#define V_EMIT_CASE(w) case w: { PP_EMIT_WARP_##w(V_IN, V_IP, V_IU, V_EO, V_EQ, V_X, V_PW) } break;
#define V_SIL_CASE(w) case w: { PP_SIL_WARP_##w(V_IS, V_IT, V_EN, V_ES, V_EF, V_X, V_PW) } break;

template <bool ROW0>
__device__ __forceinline__ void vit_silent_phase(int warp, unsigned char* sil, const unsigned char* en, bool active,
                                                 uint32_t* pr) {
    double prev_ = neg_inf();
    uint32_t pw = 0;
    switch (warp) { PP_FOR_EACH_WARP(V_SIL_CASE) default: break; }
    (void)prev_; (void)pw;
}

template <typename ObsT>
__global__ void __launch_bounds__(PP_W * 32, PP_CTAS_PER_SM)
k_viterbi_spec(const LaneBatch<ObsT> B, const VitOut O) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    ObsT* xt = reinterpret_cast<ObsT*>(smem + static_cast<size_t>(PP_NSLOTS) * kSlotBytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    unsigned char* sil = smem + lane * 8;
    unsigned char* eA = sil + PP_NS * 256;
    unsigned char* eB = eA + PP_NE * 256;
    uint32_t* prow = O.ptr + (static_cast<size_t>(B.group_row[g]) * PP_PTR_WORDS) * kLanes + lane;

    for (int s = threadIdx.x; s < PP_NSLOTS * kLanes; s += PP_W * 32) col[s] = neg_inf();
    __syncthreads();
    vit_silent_phase<true>(warp, sil, eA, ev >= 0, prow);        // row 0: start = 0, silent closure (3513-3542)
    __syncthreads();

    for (int r = 1; r <= nmax; ++r) {
        if (((r - 1) & (kTileCols - 1)) == 0) {
            fill_tile<ObsT, PP_W>(xt, B.obs, off, n, r - 1, warp, lane);
            __syncthreads();
        }
        const double x = static_cast<double>(xt[((r - 1) & (kTileCols - 1)) * kTileStride + lane]);
        const bool active = r <= n;
        const unsigned char* eo = (r & 1) ? eA : eB;             // row r-1
        unsigned char* en = (r & 1) ? eB : eA;                   // row r
        uint32_t* pr = prow + static_cast<size_t>(r) * (PP_PTR_WORDS * kLanes);
        {
            constexpr bool ROW0 = false;
            double prev_ = 0.0;
            uint32_t pw = 0;
            switch (warp) { PP_FOR_EACH_WARP(V_EMIT_CASE) default: break; }     // emitting states (3545-3562)
            (void)prev_; (void)pw;
        }
        __syncthreads();
        vit_silent_phase<false>(warp, sil, en, active, pr);                     // silent states (3564-3605)
        __syncthreads();

        if (warp == 0 && __any_sync(0xffffffffu, r == n)) {                     // final score (3614-3619)
            double score = neg_inf();
            int est = PP_END_STATE, eord = 0;
#if PP_FINITE && PP_END_FINAL_ONLY
            {
                constexpr bool ROW0 = false;
                double prev_ = 0.0;
                uint32_t pw = 0;
                double endv;                     // end's value stays in a register: nothing reads its slot
#define V_IS_END(j, pshift)                                                                     \
    {                                                                                            \
        double best = neg_inf();                                                                 \
        unsigned arg = 0;                                                                        \
        double* const dst_ = &endv;                                                              \
        constexpr int ps_ = 0;                                                                   \
        constexpr bool is_start_ = false;
#define V_PW_NONE(word)
                PP_END_ITEM(V_IS_END, V_IS_END, V_EN, V_ES, V_EF, V_X, V_PW_NONE)
                score = endv;
                eord = static_cast<int>(pw);
                (void)prev_;
            }
#elif PP_FINITE
            score = PP_LD(sil, PP_END_SIL);
#else
            score = PP_LD(en, 0);                                               // numpy.argmax(v[n]): first maximum
            est = 0;
            for (int k = 1; k < PP_M; ++k) {
                const double v = k < PP_NE ? PP_LD(en, k) : PP_LD(sil, k - PP_NE);
                if (v > score) { score = v; est = k; }
            }
#endif
            if (r == n) {
                O.score[ev] = score;
                O.end_state[ev] = est;
                O.end_ord[ev] = eord;
            }
        }
    }
}

// =============================================================================================
// Forward macro bodies (scaled linear space, see k_forward_lane)
// =============================================================================================
#define F_OPEN(dstp)                                                                             \
    {                                                                                            \
        double acc = 0.0;                                                                        \
        double* const dst_ = reinterpret_cast<double*>(dstp);                                    \
        constexpr bool is_start_ = false;
#define F_EMIS2_N(k) fma(-((x - PP_WE(k, 0)) * (x - PP_WE(k, 0))), PP_WE(k, 6), PP_WE(k, 5))
#define F_EMIS2_P(k) ((fabs(x - PP_WE(k, 0)) < 1E-4) ? PP_WE(k, 5) : neg_inf())
#define F_EMIS2_U(k)                                                                             \
    ((x == PP_WE(k, 0) && x == PP_WE(k, 1)) ? PP_WE(k, 7) : ((x >= PP_WE(k, 0) && x <= PP_WE(k, 1)) ? PP_WE(k, 5) : neg_inf()))
#define F_IN(k, pshift) F_OPEN(en + (k) * 256) const double p_ = exp2_scaled(F_EMIS2_N(k) + shiftd);
#define F_IP(k, pshift) F_OPEN(en + (k) * 256) const double p_ = exp2_scaled(F_EMIS2_P(k) + shiftd);
#define F_IU(k, pshift) F_OPEN(en + (k) * 256) const double p_ = exp2_scaled(F_EMIS2_U(k) + shiftd);
#define F_IS(j, pshift) F_OPEN(sil + (j) * 256) constexpr double p_ = 1.0;
#define F_IT(j, pshift)                                                                          \
    {                                                                                            \
        double acc = 0.0;                                                                        \
        double* const dst_ = reinterpret_cast<double*>(sil + (j) * 256);                         \
        constexpr bool is_start_ = true;                                                         \
        constexpr double p_ = 1.0;
#define F_EO(ord, k, e) acc = fma(PP_LD(eo, k), PP_WTAB[PP_PROB_BASE + (e)], acc);
#define F_EQ(ord, j, e) acc = fma(PP_LD(sil, j), PP_WTAB[PP_PROB_BASE + (e)], acc);
#define F_EN(ord, k, e) acc = fma(PP_LD(en, k), PP_WTAB[PP_PROB_BASE + (e)], acc);
#define F_ES(ord, j, e) acc = fma(PP_LD(sil, j), PP_WTAB[PP_PROB_BASE + (e)], acc);
#define F_EF(ord, e) acc = fma(prev_, PP_WTAB[PP_PROB_BASE + (e)], acc);
#define F_X()                                                                                    \
        acc *= p_;                                                                               \
        if (ROW0 && is_start_) acc = 1.0;                                                        \
        *dst_ = acc;                                                                             \
        prev_ = acc;                                                                             \
        mx = max(mx, __double2hiint(acc));                                                       \
    }
#define F_PW(word)

#define F_EMIT_CASE(w) case w: { PP_EMIT_WARP_##w(F_IN, F_IP, F_IU, F_EO, F_EQ, F_X, F_PW) } break;
#define F_SIL_CASE(w) case w: { PP_SIL_WARP_##w(F_IS, F_IT, F_EN, F_ES, F_EF, F_X, F_PW) } break;

template <bool ROW0>
__device__ __forceinline__ void fwd_silent_phase(int warp, unsigned char* sil, const unsigned char* en, int& mx) {
    double prev_ = 0.0;
    switch (warp) { PP_FOR_EACH_WARP(F_SIL_CASE) default: break; }
    (void)prev_;
}

template <typename ObsT>
__global__ void __launch_bounds__(PP_W * 32, PP_CTAS_PER_SM)
k_forward_spec(const LaneBatch<ObsT> B, const FwdOut O) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    ObsT* xt = reinterpret_cast<ObsT*>(smem + static_cast<size_t>(PP_NSLOTS) * kSlotBytes);
    int32_t* wmax = reinterpret_cast<int32_t*>(smem + static_cast<size_t>(PP_NSLOTS) * kSlotBytes +
                                               static_cast<size_t>(kTileCols) * kTileStride * sizeof(ObsT));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int ev = B.order[g * kLanes + lane];
    const int64_t off = ev >= 0 ? B.offsets[ev] : 0;
    const int n = ev >= 0 ? static_cast<int>(B.offsets[ev + 1] - off) : 0;
    const int nmax = B.group_len[g];
    unsigned char* sil = smem + lane * 8;
    unsigned char* eA = sil + PP_NS * 256;
    unsigned char* eB = eA + PP_NE * 256;

    for (int s = threadIdx.x; s < PP_NSLOTS * kLanes; s += PP_W * 32) col[s] = 0.0;
    __syncthreads();
    int mx = 0;
    fwd_silent_phase<true>(warp, sil, eA, mx);
    wmax[warp * kLanes + lane] = mx;
    __syncthreads();

    long long scale_total = 0;
    for (int r = 1; r <= nmax; ++r) {
        int M = 0;
#pragma unroll
        for (int w = 0; w < PP_W; ++w) M = max(M, wmax[w * kLanes + lane]);
        const int ex = (M >> 20) & 0x7ff;
        int shift = (ex == 0 || ex == 0x7ff) ? 0 : (kFwdTargetExp + 1023 - ex);
        shift = max(-900, min(900, shift));
        scale_total += shift;
        const double shiftd = static_cast<double>(shift);
        if (((r - 1) & (kTileCols - 1)) == 0) {
            fill_tile<ObsT, PP_W>(xt, B.obs, off, n, r - 1, warp, lane);
            __syncthreads();
        }
        const double x = static_cast<double>(xt[((r - 1) & (kTileCols - 1)) * kTileStride + lane]);
        const unsigned char* eo = (r & 1) ? eA : eB;
        unsigned char* en = (r & 1) ? eB : eA;
        mx = 0;
        {
            constexpr bool ROW0 = false;
            double prev_ = 0.0;
            switch (warp) { PP_FOR_EACH_WARP(F_EMIT_CASE) default: break; }     // emitting states (2874-2889)
            (void)prev_;
        }
        __syncthreads();
        fwd_silent_phase<false>(warp, sil, en, mx);                             // silent states (2891-2926)
        wmax[warp * kLanes + lane] = mx;
        __syncthreads();

        if (warp == 0 && __any_sync(0xffffffffu, r == n)) {                     // log_probability (3378-3396)
            double P = 0.0;
#if PP_FINITE && PP_END_FINAL_ONLY
            {
                constexpr bool ROW0 = false;
                double prev_ = 0.0;
                int mxe = 0;
#define F_IS_END(j, pshift)                                                                      \
    {                                                                                            \
        double acc = 0.0;                                                                        \
        double* const dst_ = &P;                                                                 \
        constexpr bool is_start_ = false;                                                        \
        constexpr double p_ = 1.0;
#define F_X_END()                                                                                \
        *dst_ = acc;                                                                             \
        prev_ = acc;                                                                             \
        mxe = 0;                                                                                 \
    }
                PP_END_ITEM(F_IS_END, F_IS_END, F_EN, F_ES, F_EF, F_X_END, F_PW)
                (void)prev_; (void)mxe;
            }
#elif PP_FINITE
            P = PP_LD(sil, PP_END_SIL);
#else
            for (int k = 0; k < PP_NE; ++k) P += PP_LD(en, k);
#endif
            if (r == n) {
                double lp;
                if (P == 0.0) lp = neg_inf();
                else if (!(P < __longlong_as_double(0x7FF0000000000000LL))) lp = __longlong_as_double(0x7FF8000000000000LL);
                else lp = log(P) - static_cast<double>(scale_total) * 0.693147180559945309417;
                O.logp[ev] = lp;
            }
        }
    }
}

// =============================================================================================
// host entry points (C ABI of the specialised module, bound with dlsym by pp_spec.cpp)
// =============================================================================================
extern "C" {

int pp_spec_info(int* out, int n) {
    const int v[8] = {1 /* abi */, PP_M, PP_NE, PP_NSLOTS, PP_W, PP_N_WCONST, PP_PTR_WORDS, PP_PTR_BITS};
    for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
    return 0;
}

int pp_spec_set_params(const double* w, int n, void* stream) {
    if (n != PP_N_WCONST) return -1;
    return static_cast<int>(cudaMemcpyToSymbolAsync(PP_WTAB, w, sizeof(double) * PP_N_WCONST, 0, cudaMemcpyHostToDevice,
                                                    static_cast<cudaStream_t>(stream)));
}

int pp_spec_viterbi(int obs_f64, const void* obs, const int64_t* offsets, const int32_t* order, const int32_t* group_len,
                    const int64_t* group_row, double* score, int32_t* end_state, int32_t* end_ord, uint32_t* ptr,
                    int n_groups, void* stream) {
    const VitOut O{score, end_state, end_ord, ptr};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (obs_f64) {
        const size_t smem = lane_smem_bytes(PP_NSLOTS, 8, PP_W);
        e = cudaFuncSetAttribute(k_viterbi_spec<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const LaneBatch<double> B{static_cast<const double*>(obs), offsets, order, group_len, group_row};
        k_viterbi_spec<double><<<n_groups, PP_W * 32, smem, s>>>(B, O);
    } else {
        const size_t smem = lane_smem_bytes(PP_NSLOTS, 4, PP_W);
        e = cudaFuncSetAttribute(k_viterbi_spec<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const LaneBatch<float> B{static_cast<const float*>(obs), offsets, order, group_len, group_row};
        k_viterbi_spec<float><<<n_groups, PP_W * 32, smem, s>>>(B, O);
    }
    return static_cast<int>(cudaGetLastError());
}

int pp_spec_forward(int obs_f64, const void* obs, const int64_t* offsets, const int32_t* order, const int32_t* group_len,
                    const int64_t* group_row, double* logp, int n_groups, void* stream) {
    const FwdOut O{logp};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (obs_f64) {
        const size_t smem = lane_smem_bytes(PP_NSLOTS, 8, PP_W);
        e = cudaFuncSetAttribute(k_forward_spec<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const LaneBatch<double> B{static_cast<const double*>(obs), offsets, order, group_len, group_row};
        k_forward_spec<double><<<n_groups, PP_W * 32, smem, s>>>(B, O);
    } else {
        const size_t smem = lane_smem_bytes(PP_NSLOTS, 4, PP_W);
        e = cudaFuncSetAttribute(k_forward_spec<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const LaneBatch<float> B{static_cast<const float*>(obs), offsets, order, group_len, group_row};
        k_forward_spec<float><<<n_groups, PP_W * 32, smem, s>>>(B, O);
    }
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
