// pp_peaks.cu -- measured float64 issue peaks of the device: the denominators of the compute rooflines bench.py reports.
// The sweeps of this library are bound by float64 add / compare / FMA issue, for which MEASURED_PEAKS.json has no entry;
// pp_device_peaks measures them live (a few milliseconds) with independent dependency chains, 8 per thread.
#include "pp_common.h"

namespace pp {

template <int MODE>
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, double w, double e, int iters) {
    constexpr int CH = 8;
    double s[CH];
    const double seed = out[threadIdx.x & 31];
#pragma unroll
    for (int c = 0; c < CH; ++c) s[c] = seed + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) s[c] = MODE == 0 ? __dadd_rn(s[c], w) : __fma_rn(s[c], w, e);
    }
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc += s[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static double measure(double* d, int sms, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = sms * 8;
    k_fp64_peak<MODE><<<grid, 256>>>(d, -0.25, 1e-3, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_fp64_peak<MODE><<<grid, 256>>>(d, -0.25, 1e-3, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return static_cast<double>(grid) * 256.0 * iters * 8.0 / (best * 1e-3);
}

}  // namespace pp

extern "C" int pp_device_peaks(int device, double* out4) {
    using namespace pp;
    if (!out4) { set_error("null argument"); return PP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device"); return PP_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("device out of range"); return PP_ERR_ARG; }
    PP_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    PP_CUDA(cudaGetDeviceProperties(&p, device));
    int khz = 0;
    PP_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    double* d = nullptr;
    PP_CUDA(cudaMalloc(&d, static_cast<size_t>(p.multiProcessorCount) * 8 * 256 * sizeof(double)));
    PP_CUDA(cudaMemset(d, 0, static_cast<size_t>(p.multiProcessorCount) * 8 * 256 * sizeof(double)));
    out4[0] = measure<0>(d, p.multiProcessorCount, 4000);     // float64 adds per second
    out4[1] = measure<1>(d, p.multiProcessorCount, 4000);     // float64 FMAs per second
    out4[2] = p.multiProcessorCount;
    out4[3] = khz * 1e3;
    cudaFree(d);
    PP_CUDA(cudaGetLastError());
    return PP_OK;
}
