// pp_profile.cu -- instantiation and launchers of the linear-profile kernels (pp_profile_kernels.cuh).
#ifndef PP_PROF_NCHAIN
#define PP_PROF_NCHAIN 1              // chain warps of the Viterbi sweep (2: one chain each -- measured slower)
#endif
#include "pp_profile_kernels.cuh"

namespace pp {

int profile_viterbi_launch(const ProfSched& S, const LaneBatchAny& B, const VitOut& O, int n_groups, cudaStream_t stream) {
    const size_t smem = prof_smem_bytes(S.Kr, S.NW, kProfPositionsPerWarp);
    cudaError_t e = cudaFuncSetAttribute(k_profile_viterbi<kProfPositionsPerWarp, PP_PROF_NCHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    k_profile_viterbi<kProfPositionsPerWarp, PP_PROF_NCHAIN><<<n_groups, (S.NW + PP_PROF_NCHAIN) * 32, smem, stream>>>(S, B, O);
    return static_cast<int>(cudaGetLastError());
}

int profile_forward_launch(const ProfSched& S, const LaneBatchAny& B, const FwdOut& O, int n_groups, cudaStream_t stream) {
    const size_t smem = prof_smem_bytes(S.Kr, S.NW, kProfPositionsPerWarp);
    cudaError_t e = cudaFuncSetAttribute(k_profile_forward<kProfPositionsPerWarp>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    k_profile_forward<kProfPositionsPerWarp><<<n_groups, (S.NW + 1) * 32, smem, stream>>>(S, B, O);
    return static_cast<int>(cudaGetLastError());
}

#ifdef PP_PHASE_TIMING
int profile_debug_phase_cycles(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_prof_phase, sizeof(unsigned long long) * 2 * 256);
    if (reset) { static unsigned long long z[512]; cudaMemcpyToSymbol(g_prof_phase, z, sizeof z); }
    return 0;
}
#endif

}  // namespace pp
#ifdef PP_PHASE_TIMING
extern "C" int pp_debug_profile_phase_cycles(unsigned long long* out, int reset) { return pp::profile_debug_phase_cycles(out, reset); }
#endif
