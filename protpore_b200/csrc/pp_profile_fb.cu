// pp_profile_fb.cu -- instantiation and launcher of the linear-profile forward-backward kernel (pp_profile_fb.cuh).
#include "pp_profile_fb.cuh"

namespace pp {

size_t profile_fb_smem(int Kr, int NW) { return pfb_smem_bytes(Kr, NW, kPfbPositionsPerWarp); }

int profile_fb_launch(const ProfFbSched& S, const LaneBatchAny& B, int grid, cudaStream_t stream) {
    const size_t smem = pfb_smem_bytes(S.Kr, S.NW, kPfbPositionsPerWarp);
    cudaError_t e = cudaFuncSetAttribute(k_profile_fb<kPfbPositionsPerWarp, kPfbThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    k_profile_fb<kPfbPositionsPerWarp, kPfbThreads><<<grid, (S.NW + 1) * 32, smem, stream>>>(S, B);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace pp
