// fp64_peak.cu -- measured float64 issue peaks of one B200: the denominators of bench.py's compute roofline.
//   dadd    : independent DADD chains (what a Viterbi candidate add is)
//   dfma    : independent DFMA chains (what a forward edge is)
//   cand    : the Viterbi candidate mix with perfect ILP and no loads: DADD, DADD, DSETP, 2 x FSEL, SEL
//   cand+f  : the same with one forward DFMA per candidate (the fused sweep's inner step)
// Every thread runs CH independent chains for `iters` rounds; grid = SMs x ctas_per_sm, 256 threads per CTA.
// Prints warp-instructions per clock per SM (from the kernel's wall time and the SM clock it reads) and Gop/s.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(double* out, double w, double e, int iters) {
    double s[CH], f[CH];
    unsigned arg[CH];
    const double seed = out[threadIdx.x & 31];
#pragma unroll
    for (int c = 0; c < CH; ++c) { s[c] = seed + c; f[c] = seed * 0.5 + c; arg[c] = c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) {
                s[c] = __dadd_rn(s[c], w);
            } else if (MODE == 1) {
                s[c] = __fma_rn(s[c], w, e);
            } else {
                // candidate: t = (s[c'] + w) + e ; if (t > s[c]) { s[c] = t; arg = i }
                const double t = __dadd_rn(__dadd_rn(s[(c + 1) % CH], w), e);
                if (t > s[c]) { s[c] = t; arg[c] = i; }
                if (MODE == 3) f[c] = __fma_rn(f[(c + 1) % CH], w, f[c]);
            }
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc += s[c] + f[c] + arg[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(const char* name, double ops_per_step, double instr_per_step, int sms, int ctas_per_sm, double* d, int iters,
                double clock_ghz) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_peak<MODE><<<sms * ctas_per_sm, 256>>>(d, -0.25, 1e-3, 100);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_peak<MODE><<<sms * ctas_per_sm, 256>>>(d, -0.25, 1e-3, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double steps = double(sms) * ctas_per_sm * 256.0 * iters * CH;       // per-thread steps in total
    const double s = best * 1e-3;
    printf("{\"name\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"fp64_ops_per_s\": %.4e, \"warp_instr_per_clk_per_sm\": %.3f, "
           "\"steps_per_s\": %.4e}\n",
           name, ctas_per_sm, best, steps * ops_per_step / s, steps / 32.0 * instr_per_step / s / (clock_ghz * 1e9) / sms,
           steps / s);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, size_t(sms) * 8 * 256 * 8); cudaMemset(d, 0, size_t(sms) * 8 * 256 * 8);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, sms, ghz);
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    for (int c : {1, 2, 4, 8}) {
        run<0>("dadd", 1, 1, sms, c, d, iters, ghz);
        run<1>("dfma", 1, 1, sms, c, d, iters, ghz);
        run<2>("viterbi_candidate(2 DADD + DSETP + 2 FSEL + SEL)", 3, 6, sms, c, d, iters, ghz);
        run<3>("fused_candidate(+1 DFMA)", 4, 7, sms, c, d, iters, ghz);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
