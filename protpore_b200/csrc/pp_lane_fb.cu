// pp_lane_fb.cu -- instantiation of the lane-per-event forward-backward kernel (pp_lane_fb.cuh), its own constant arena
// (forward runs followed by backward runs).
#include "pp_lane_fb.cuh"

namespace pp {

static ArenaOwner g_arena;

size_t lane_fb_smem(int n_slots, int n_w, int ne) { return fb_smem_bytes(n_slots, n_w, ne, kLaneWarps); }

int lane_fb_launch(const LaneSched& S, const FbSched& F, const LaneBatchAny& B, int grid, size_t smem,
                   cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner) {
    return g_arena.run(
        owner, stream,
        [&] { return cudaMemcpyToSymbolAsync(c_arena, arena, arena_doubles * sizeof(double), 0, cudaMemcpyHostToDevice, stream); },
        [&] {
            cudaError_t e = cudaFuncSetAttribute(k_fb_lane<kLaneWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            k_fb_lane<kLaneWarps><<<grid, kLaneWarps * 32, smem, stream>>>(S, F, B);
            return static_cast<int>(cudaGetLastError());
        });
}

int lane_fb_finalize(const double* acc, int n_ctas, int n_slots, const int32_t* acc_map, const int32_t* acc_log2,
                     int n_edges, int ne, double* edge_counts, double* emit_moments, const double* minmax_tmp,
                     double* emit_minmax, cudaStream_t stream) {
    k_fb_finalize<<<n_slots, kLanes, 0, stream>>>(acc, n_ctas, n_slots, acc_map, acc_log2, n_edges, edge_counts, emit_moments);
    k_fb_minmax_merge<<<(ne + 127) / 128, 128, 0, stream>>>(minmax_tmp, ne, emit_minmax);
    return static_cast<int>(cudaGetLastError());
}

void lane_fb_invalidate(const void* owner) { g_arena.invalidate(owner); }

}  // namespace pp
