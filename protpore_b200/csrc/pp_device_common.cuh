// pp_device_common.cuh -- device helpers shared by every kernel family: batch / output descriptors, the
// exact float64 emission log-density, the log2-domain emission of the linear-space sweeps, observation
// tile staging, and the traceback kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pp_lowering.h"

namespace pp {

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

// first maximum of s[lo, hi): a balanced tree of "right replaces left only if strictly greater" steps picks the
// same index as the reference's left-to-right strict-'>' scan (the lowest index attaining the maximum), with a
// dependent chain of log2(n) compare+select steps instead of n.
template <int LO, int HI>
struct FirstMax {
    __device__ static __forceinline__ void run(const double* s, double& best, uint32_t& arg) {
        constexpr int MID = (LO + HI) / 2;
        double b2;
        uint32_t a2;
        FirstMax<LO, MID>::run(s, best, arg);
        FirstMax<MID, HI>::run(s, b2, a2);
        if (b2 > best) { best = b2; arg = a2; }
    }
};
template <int LO>
struct FirstMax<LO, LO + 1> {
    __device__ static __forceinline__ void run(const double* s, double& best, uint32_t& arg) { best = s[LO]; arg = LO; }
};

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
        r = (x >= R.a0 && x <= R.a1) ? R.a3 : neg_inf();      // a3 = 0 when a0 == a1 (host), the x == a == b branch
    }
    return __dadd_rn(r, R.logw);
}

// 2^(t + shift) as a double, t in the log2 domain.  The fraction goes through MUFU.EX2 (fp32,
// rel. error ~2^-22); the integer part is written straight into the exponent field.  Below 2^-1000
// the factor is flushed to zero, above 2^1000 it becomes +inf (caught by the caller's status).
__device__ __forceinline__ double exp2_scaled(double t) {
    // clamp instead of branching: below -1000 the result is made an exact 0, above 1000 it saturates to 2^1000 * p
    // (the caller's row maximum then reaches inf / NaN and the event is re-run by the exact kernel)
    const double tc = fmin(fmax(t, -1001.0), 1000.0);          // fmax(NaN, -1001) = -1001 -> 0
    // round to nearest integer with the 1.5 * 2^52 trick: two DADDs, the integer lands in the low word
    const double magic = 6755399441055744.0;
    const double tm = tc + magic;
    const int n = __double2loint(tm);
    const float f = static_cast<float>(tc - (tm - magic));
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f));
    const double scale = __hiloint2double((n + 1023) << 20, 0);
    const double r = static_cast<double>(p) * scale;
    return (t > -1000.0) ? r : 0.0;
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
// Traceback (yahmm.pyx:3632-3646): one thread per event walks the ordinals back from (n, end).
//   k_traceback        path == nullptr: counts only; else writes the path front-to-back at path_off[ev] (needs the count)
//   k_traceback_walk   ONE walk: counts and writes the states, back to front, into a per-event scratch region of
//                      4 n + 64 entries (a path visits ~1.7 states per observation; an event that needs more raises
//                      *overflow and the caller falls back to the two passes); k_traceback_compact then copies the
//                      paths to their final, densely packed place with coalesced accesses.
// The per-state lookup (where the ordinal sits in a pointer row, where the state's in-edge list starts) is one 16-byte
// record, staged in shared memory together with the in-edge sources when the model is small enough, so a step of the
// walk costs one global load: the ordinal itself.
// ------------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) TraceRec {
    uint32_t edge_begin;               // first entry of the state's in-edge (slot) list
    uint32_t byte_off;                 // byte of lane 0's ordinal inside a pointer row
    uint32_t stride;                   // bytes per lane of the unit holding it
    uint32_t shift_mask;               // bit position << 4 | mask (ordinals are at most 4 bits wide)
};
struct TraceArgs {
    const uint32_t* ptr;
    const int32_t* order;
    const int64_t* group_row;
    const int64_t* offsets;
    const double* score;
    const int32_t* end_state;
    const int32_t* end_ord;
    const TraceRec* rec;               // [m]
    const uint16_t* src;               // [lowered edges / slots] source state
    int32_t m, n_src, ne, start, end_final_only;
    int64_t row_bytes;                 // bytes per pointer row (all 32 lanes)
    int64_t e0;                        // first event of the chunk (scratch regions are laid out from it)
};
constexpr int kTraceSmemBytes = 40 * 1024;
// dynamic shared memory a traceback launch needs: both tables when they fit, else none (they are read through L1)
inline size_t trace_smem_bytes(int m, int n_src) {
    const size_t need = static_cast<size_t>(m) * sizeof(TraceRec) + static_cast<size_t>(n_src) * sizeof(uint16_t);
    return need <= kTraceSmemBytes ? need : 0;
}

struct TraceTables {
    const TraceRec* rec;
    const uint16_t* src;
    // stage both tables in shared memory when they fit (all threads of the block call this)
    __device__ __forceinline__ TraceTables(const TraceArgs& T, unsigned char* smem) {
        const size_t need = static_cast<size_t>(T.m) * sizeof(TraceRec) + static_cast<size_t>(T.n_src) * sizeof(uint16_t);
        if (need <= kTraceSmemBytes) {
            TraceRec* r = reinterpret_cast<TraceRec*>(smem);
            uint16_t* q = reinterpret_cast<uint16_t*>(r + T.m);
            for (int i = threadIdx.x; i < T.m; i += blockDim.x) r[i] = T.rec[i];
            for (int i = threadIdx.x; i < T.n_src; i += blockDim.x) q[i] = T.src[i];
            __syncthreads();
            rec = r;
            src = q;
        } else {
            rec = T.rec;
            src = T.src;
        }
    }
};

// one step back from (px, py): the state the stored ordinal leads to
__device__ __forceinline__ int trace_step(const TraceArgs& T, const TraceTables& tab, const uint8_t* base, size_t row_stride, int lane,
                                          int px, int py, bool first, int ev) {
    const TraceRec R = tab.rec[py];
    int ord;
    if (first && T.end_final_only) ord = T.end_ord[ev];
    else {
        const uint8_t* a = base + static_cast<size_t>(px) * row_stride + R.byte_off + lane * R.stride;
        const int sh = R.shift_mask >> 4, mk = R.shift_mask & 15;
        uint32_t u = a[0];
        if ((mk << sh) > 0xff) u |= static_cast<uint32_t>(a[1]) << 8;
        ord = static_cast<int>((u >> sh) & mk);
    }
    return tab.src[R.edge_begin + ord];
}

// The traceback kernels are launched by pp_api.cu only: other translation units include this header for the helpers above
#ifdef PP_WITH_TRACEBACK_KERNELS
static __global__ void __launch_bounds__(128) k_traceback(const TraceArgs T, int64_t* __restrict__ path_len,
                                                          const int64_t* __restrict__ path_off, uint16_t* __restrict__ path, int n_slots) {
    extern __shared__ __align__(16) unsigned char smem[];
    const TraceTables tab(T, smem);
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
    const size_t row_stride = static_cast<size_t>(T.row_bytes);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(T.ptr) + static_cast<size_t>(T.group_row[g]) * row_stride;
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
        const int src = trace_step(T, tab, base, row_stride, lane, px, py, first, ev);
        first = false;
        if (py < T.ne) --px;                             // emitting states point into the previous row
        py = src;
        ++len;
    }
    if (out) out[pos] = static_cast<uint16_t>(py);
    else path_len[ev] = len;
}

__device__ __forceinline__ int64_t trace_tmp_base(const TraceArgs& T, int ev) {
    return 4 * (T.offsets[ev] - T.offsets[T.e0]) + 64 * static_cast<int64_t>(ev - T.e0);
}

static __global__ void __launch_bounds__(128) k_traceback_walk(const TraceArgs T, int64_t* __restrict__ path_len,
                                                               uint16_t* __restrict__ tmp, int32_t* __restrict__ overflow, int n_slots) {
    extern __shared__ __align__(16) unsigned char smem[];
    const TraceTables tab(T, smem);
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int ev = T.order[slot];
    if (ev < 0) return;
    const int g = slot >> 5, lane = slot & 31;
    if (T.score[ev] == neg_inf()) {
        path_len[ev] = -1;
        return;
    }
    const int n = static_cast<int>(T.offsets[ev + 1] - T.offsets[ev]);
    const size_t row_stride = static_cast<size_t>(T.row_bytes);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(T.ptr) + static_cast<size_t>(T.group_row[g]) * row_stride;
    const int64_t cap = 4 * static_cast<int64_t>(n) + 64;          // a multiple of 4: the region ends on an 8-byte boundary
    uint64_t* out = reinterpret_cast<uint64_t*>(tmp + trace_tmp_base(T, ev) + cap);   // filled back to front, 4 states per store
    int64_t len = 0;
    int px = n, py = T.end_state[ev];
    bool first = true;
    uint64_t word = 0;
    for (;;) {
        ++len;
        word = (word << 16) | static_cast<uint64_t>(py);            // the four latest states, the latest in the low bits:
        if ((len & 3) == 0 && len <= cap) *--out = word;            // little-endian, so memory reads front to back
        if (px == 0 && py == T.start) break;
        const int src = trace_step(T, tab, base, row_stride, lane, px, py, first, ev);
        first = false;
        if (py < T.ne) --px;
        py = src;
    }
    if ((len & 3) != 0 && len <= cap) *--out = word << (16 * (4 - (len & 3)));   // the last, partial word: its states in the upper slots
    if (len > cap) atomicAdd(overflow, 1);
    path_len[ev] = len;
}

// one warp per event: scratch region -> the event's place in the packed output
static __global__ void k_traceback_compact(const TraceArgs T, const int64_t* __restrict__ path_len, const int64_t* __restrict__ path_off,
                                           const uint16_t* __restrict__ tmp, uint16_t* __restrict__ path, int64_t e1) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t ev = T.e0 + wid; ev < e1; ev += nw) {
        const int64_t len = path_len[ev];
        if (len <= 0) continue;
        const int64_t cap = 4 * (T.offsets[ev + 1] - T.offsets[ev]) + 64;
        const uint16_t* s = tmp + trace_tmp_base(T, static_cast<int>(ev)) + cap - len;
        uint16_t* d = path + path_off[ev];
        for (int64_t i = lane; i < len; i += 32) d[i] = s[i];
    }
}

#endif  // PP_WITH_TRACEBACK_KERNELS

constexpr int kFwdTargetExp = 900;   // a row's largest value is steered to ~2^900

}  // namespace pp
