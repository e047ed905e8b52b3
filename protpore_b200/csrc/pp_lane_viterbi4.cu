// pp_lane_viterbi4.cu -- one instantiation of the lane-per-event kernels (pp_lane_kernels.cuh), its own constant arena.
#include "pp_lane_kernels.cuh"

namespace pp {

static ArenaOwner g_arena;

int lane_viterbi4_launch(const LaneSched& S, const LaneBatchAny& B, const VitOut& O, int n_groups, size_t smem,
                     cudaStream_t stream, const double* arena, size_t arena_doubles, const void* owner) {
    return g_arena.run(
        owner, stream,
        [&] { return cudaMemcpyToSymbolAsync(c_arena, arena, arena_doubles * sizeof(double), 0, cudaMemcpyHostToDevice, stream); },
        [&] {
            cudaError_t e = cudaFuncSetAttribute(k_viterbi_lane<kLaneWarps, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            k_viterbi_lane<kLaneWarps, 4><<<n_groups, kLaneWarps * 32, smem, stream>>>(S, B, O);
            return static_cast<int>(cudaGetLastError());
        });
}

void lane_viterbi4_invalidate(const void* owner) { g_arena.invalidate(owner); }

#ifdef PP_PHASE_TIMING
int lane_debug_phase_cycles(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 128);
    if (reset) { unsigned long long z[128] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); }
    return 0;
}
#endif

}  // namespace pp
