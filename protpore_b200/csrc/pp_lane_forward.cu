// pp_lane_forward.cu -- one instantiation of the lane-per-event kernels (pp_lane_kernels.cuh), its own constant arena.
#include "pp_lane_kernels.cuh"

namespace pp {

static ArenaOwner g_arena;

int lane_forward_launch(const LaneSched& S, const LaneBatchAny& B, const FwdOut& O, int n_groups, size_t smem,
                     cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner) {
    return g_arena.run(
        owner, stream,
        [&] { return cudaMemcpyToSymbolAsync(c_arena, arena, arena_doubles * sizeof(double), 0, cudaMemcpyHostToDevice, stream); },
        [&] {
            cudaError_t e = cudaFuncSetAttribute(k_forward_lane<kLaneWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            k_forward_lane<kLaneWarps><<<n_groups, kLaneWarps * 32, smem, stream>>>(S, B, O);
            return static_cast<int>(cudaGetLastError());
        });
}

void lane_forward_invalidate(const void* owner) { g_arena.invalidate(owner); }

}  // namespace pp
