// pp_lane_types.h -- kernel-parameter structs and launch entry points of the lane-per-event kernels.
// The kernels themselves (pp_lane_kernels.cuh) are instantiated in three translation units that compile in
// parallel: pp_lane_viterbi4.cu (4-bit ordinals), pp_lane_forward.cu, and their global-column twins pp_lane_viterbi4g.cu /
// pp_lane_forwardg.cu (an 8-bit-ordinal build existed but was unreachable: lane-lowered items have <= 8 in-edges).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "pp_device_common.cuh"

namespace pp {

constexpr int kArenaDoubles = 7680;                       // 60 KB of the 64 KB constant bank

struct LaneSched {
    const EdgeRec* end_edges[2];     // lowered in-edges of the end state, per row parity (final score)
    int32_t run_ranges[kMaxLaneWarps * 6];       // per warp: {begin, end} of its late-emitting, silent and early-emitting runs
    const double* wtab;              // [n_low] transition weights in lowered-edge order: log p (Viterbi) or p (forward)
    const double* etab;              // [ne][4] emission records of this kernel family
    int32_t n_low;
    int32_t etab_doubles;            // doubles per emission record (5 Viterbi, 4 forward)
    int32_t m, ne, ns, n_slots;
    int32_t start_slot_off;          // byte offset of the start state's slot
    int32_t end_slot_off;            // byte offset of the end state's slot
    int32_t end_state, end_edge_begin, end_n_edges, finite, end_final_only;
    int32_t n_ptr_words, ptr_bits;
    unsigned char* gcol;             // global-column kernels (models too large for shared memory): per-CTA column scratch
    uint64_t gcol_stride;            // bytes per CTA
};

// tile of staged observations: 32 columns for float32, 16 for float64 (same bytes)
constexpr int kLaneTileCols = 16;   // the tile holds doubles whatever the input type: one kernel body for float32 and float64 events

// type-erased batch: the same kernel serves float32 and float64 observations (halves the instantiations)
struct LaneBatchAny {
    const void* obs;
    int32_t obs_f64;
    const int64_t* offsets;
    const int32_t* order;
    const int32_t* group_len;
    const int64_t* group_row;
};

__host__ __device__ inline size_t lane_smem_bytes2(int n_slots, int n_low, int ne, int warps) {
    return static_cast<size_t>(n_slots) * kSlotBytes + static_cast<size_t>(n_low) * 8 + static_cast<size_t>(ne) * 40 +
           static_cast<size_t>(32) * 33 * 4 + static_cast<size_t>(warps) * kLanes * sizeof(int32_t) + 64;
}


// One __constant__ arena per kernel translation unit (and device) holds the run records of ONE model at a time.  Handles
// driven from different host threads take turns: run() serialises "upload the image if another model's is resident, then
// launch" under a mutex, and orders the upload behind the last launch that read the old image (an event), so a kernel
// never has its run records replaced under it.
struct ArenaOwner {
    static constexpr int kMaxDevices = 64;
    std::mutex mu;
    const void* owner[kMaxDevices] = {};
    cudaEvent_t last[kMaxDevices] = {};
    template <typename Upload, typename Launch>
    int run(const void* who, cudaStream_t stream, Upload upload, Launch launch) {
        std::lock_guard<std::mutex> lock(mu);
        int dev = 0;
        cudaGetDevice(&dev);
        dev = dev < 0 || dev >= kMaxDevices ? 0 : dev;
        if (owner[dev] != who) {
            if (last[dev]) cudaStreamWaitEvent(stream, last[dev], 0);
            const cudaError_t e = upload();
            if (e != cudaSuccess) return static_cast<int>(e);
            owner[dev] = who;
        }
        const int rc = launch();
        if (!last[dev]) cudaEventCreateWithFlags(&last[dev], cudaEventDisableTiming);
        if (last[dev]) cudaEventRecord(last[dev], stream);
        return rc;
    }
    void invalidate(const void* who) {
        std::lock_guard<std::mutex> lock(mu);
        for (int d = 0; d < kMaxDevices; ++d)
            if (owner[d] == who) owner[d] = nullptr;
    }
};

// Launchers.  Each kernel translation unit owns a __constant__ arena holding the run records of ONE model;
// `arena` / `arena_doubles` / `owner` let it re-upload (stream-ordered) when another handle's image is resident.
int lane_viterbi4_launch(const LaneSched& S, const LaneBatchAny& B, const VitOut& O, int n_groups, size_t smem,
                         cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner);
int lane_forward_launch(const LaneSched& S, const LaneBatchAny& B, const FwdOut& O, int n_groups, size_t smem,
                        cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner);
// the same kernels with the column in global memory (S.gcol): pp_lane_gcol.cu
int lane_viterbi4g_launch(const LaneSched& S, const LaneBatchAny& B, const VitOut& O, int n_groups, size_t smem,
                          cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner);
int lane_forwardg_launch(const LaneSched& S, const LaneBatchAny& B, const FwdOut& O, int n_groups, size_t smem,
                         cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner);
// lane-per-event forward-backward kernel (pp_lane_fb.cuh)
struct FbSched {
    int32_t run_ranges[kMaxLaneWarps * 6];   // per warp: {begin, end} of its H, silent and emitting runs (arena indices)
    int32_t acc_base[kMaxLaneWarps];
    const double* wtab;                      // backward weights (probabilities) in backward-lowered edge order
    int32_t n_w;
    int32_t n_acc_slots;
    int32_t row_slots;                       // ns + ne + 1 (the last slot holds S_r)
    int32_t n_groups;
    int32_t group_stride;                    // coprime to n_groups: ticket t works on group (t * stride) mod n_groups
    unsigned char* scratch;                  // per CTA: (max_len + 1) rows of row_slots * 256 bytes
    size_t scratch_stride;                   // bytes per CTA
    double* acc;                             // per CTA: n_acc_slots * 32 doubles, zeroed by the host
    int32_t* counter;                        // work-stealing group counter (zeroed by the host)
    int32_t* fail;                           // number of events whose backward sweep failed its checks
    double* logp;                            // [n_events]
    double* minmax;                          // [2 * ne] min / max over events that gave the state non-zero weight
};
size_t lane_fb_smem(int n_slots, int n_w, int ne);
int lane_fb_launch(const LaneSched& S, const FbSched& F, const LaneBatchAny& B, int grid, size_t smem, cudaStream_t stream,
                   const double* arena, size_t arena_doubles, const void* owner);
int lane_fb_finalize(const double* acc, int n_ctas, int n_slots, const int32_t* acc_map, const int32_t* acc_log2,
                     int n_edges, int ne, double* edge_counts, double* emit_moments, const double* minmax_tmp,
                     double* emit_minmax, cudaStream_t stream);
void lane_arena_invalidate(const void* owner);           // forget `owner` in every arena (model changed / destroyed)
#ifdef PP_PHASE_TIMING
int lane_debug_phase_cycles(unsigned long long* out, int reset);
#endif

}  // namespace pp
