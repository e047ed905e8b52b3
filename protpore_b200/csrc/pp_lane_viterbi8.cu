// pp_lane_viterbi8.cu -- one instantiation of the lane-per-event kernels (pp_lane_kernels.cuh), its own constant arena.
#include "pp_lane_kernels.cuh"

namespace pp {

static const void* g_owner = nullptr;

int lane_viterbi8_launch(const LaneSched& S, const LaneBatchAny& B, const VitOut& O, int n_groups, size_t smem,
                     cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner) {
    cudaError_t e;
    if (g_owner != owner) {
        e = cudaMemcpyToSymbolAsync(c_arena, arena, arena_doubles * sizeof(double), 0, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return static_cast<int>(e);
        g_owner = owner;
    }
    e = cudaFuncSetAttribute(k_viterbi_lane<kLaneWarps, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    k_viterbi_lane<kLaneWarps, 8><<<n_groups, kLaneWarps * 32, smem, stream>>>(S, B, O);
    return static_cast<int>(cudaGetLastError());
}

void lane_viterbi8_invalidate(const void* owner) { if (g_owner == owner) g_owner = nullptr; }

}  // namespace pp
