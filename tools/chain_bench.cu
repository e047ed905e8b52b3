// Isolated timing of the silent-chain run body (vit_items<2, MODE_CHAIN>) and of a late-emitting N7 run body:
// one warp (or W warps) alone on an SM, so the numbers are pure dependent-latency / code-quality figures.
#include <cstdio>
#include "../protpore_b200/csrc/pp_lane_kernels.cuh"
using namespace pp;

template <int NE, int MODE, int KIND, bool PIPE = false>
__global__ void k_body(int count, int iters, int nsrc_stride, long long* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    for (int i = threadIdx.x; i < 400 * 32; i += blockDim.x) col[i] = -1.0 - (i % 97) * 0.01;
    double* wt = col + 400 * 32;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) wt[i] = -0.5 - (i % 13) * 0.1;
    double* et = wt + 1024;
    for (int i = threadIdx.x; i < 128 * 5; i += blockDim.x) et[i] = (i % 5 == 1) ? 2.0 : ((i % 5 == 2) ? 0.5 : 0.1 * (i % 7));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t colb = static_cast<uint32_t>(__cvta_generic_to_shared(smem + lane * 8));
    const uint32_t wtab = static_cast<uint32_t>(__cvta_generic_to_shared(wt));
    const uint32_t etab = static_cast<uint32_t>(__cvta_generic_to_shared(et));
    static uint8_t* sink;
    __shared__ uint8_t rowbuf[8192];
    VitCtx C{rowbuf, lane, true, 30.0};
    double prev = -3.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        RunPtrs<NE> P;
#pragma unroll
        for (int j = 0; j < NE; ++j) { P.sp[j] = colb + (129 + j) * 256; P.ss[j] = nsrc_stride; }
        P.dp = colb; P.ds = 256; P.wp = wtab; P.ws = NE * 8; P.ep = etab; P.es = 40; P.pbase = 0;
        if (PIPE) vit_chain_pipelined<NE, 4>(P, count, C, prev);
        else vit_items<NE, MODE, KIND, 4>(P, count, C, false, prev);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (prev == 12345.0) sink = rowbuf;
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    const size_t smem = 400 * 256 + 1024 * 8 + 128 * 40;
    long long h;
    auto run = [&](const char* name, void (*k)(int, int, int, long long*), int threads, int count, int stride) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<1, threads, smem>>>(count, 200, stride, d); cudaDeviceSynchronize();
        k<<<1, threads, smem>>>(count, 200, stride, d); cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %2d warp(s): %7.1f cycles per item\n", name, threads / 32, double(h) / (200.0 * count));
    };
    for (int threads : {32, 128, 256}) {
        run("chain link  (NE=2, MODE_CHAIN, G=4)", k_body<2, MODE_CHAIN, RUN_SILENT>, threads, 40, 768);
        run("chain link  pipelined", k_body<2, MODE_CHAIN, RUN_SILENT, true>, threads, 40, 768);
        run("silent item (NE=1, MODE_SILENT, G=4)", k_body<1, MODE_SILENT, RUN_SILENT>, threads, 40, 768);
        run("emit item   (NE=2, point mass, G=4)", k_body<2, MODE_EMIT, EMIS_POINT>, threads, 40, 768);
        run("emit item   (NE=2, uniform, G=4)", k_body<2, MODE_EMIT, EMIS_UNIFORM>, threads, 40, 768);
        run("emit item   (NE=7, normal, G=2)", k_body<7, MODE_EMIT, EMIS_NORMAL>, threads, 6, 768);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
