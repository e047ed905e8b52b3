// Micro-benchmarks of the latencies that bound a silent-chain link: dependent DADD, DADD->DSETP->FSEL, LDS.64.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dadd(double* out, double w, int iters) {
    double s = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) s = __dadd_rn(s, w);
    }
    long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) printf("dependent DADD: %.2f cycles\n", double(t1 - t0) / (iters * 16.0));
}
__global__ void k_link(double* out, const double* a, double w, int iters) {
    double s = out[threadIdx.x];
    double av[16];
    for (int j = 0; j < 16; ++j) av[j] = a[j * 32 + threadIdx.x];
    unsigned arg = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double t = __dadd_rn(s, w);
            const bool p = t > av[j];
            s = p ? t : av[j];
            arg += p;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = s + arg;
    if (threadIdx.x == 0) printf("chain link DADD->DSETP->FSEL: %.2f cycles\n", double(t1 - t0) / (iters * 16.0));
}
__global__ void k_link_spec(double* out, const double* a, double w, int iters) {
    double t = out[threadIdx.x];
    double av[16], uv[16];
    for (int j = 0; j < 16; ++j) { av[j] = a[j * 32 + threadIdx.x]; uv[j] = av[j] + w; }
    unsigned arg = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const bool p = t > av[j];
            const double z = __dadd_rn(t, w);
            t = p ? z : uv[j];
            arg += p;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = t + arg;
    if (threadIdx.x == 0) printf("speculative link max(DSETP,DADD)->FSEL: %.2f cycles\n", double(t1 - t0) / (iters * 16.0));
}
__global__ void k_lds(double* out, int iters) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 7 + 3) % 1024;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) idx = (int)sm[idx];
    }
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) printf("dependent LDS.64 (+F2I): %.2f cycles\n", double(t1 - t0) / (iters * 16.0));
}
__global__ void k_sel(int* out, int iters) {
    int s = out[threadIdx.x], b = out[32 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) s = (s > b) ? s - 3 : s + 5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) printf("dependent ISETP+SEL/IADD: %.2f cycles per step\n", double(t1 - t0) / (iters * 16.0));
}
int main() {
    double *d, *a; int* di;
    cudaMalloc(&d, 1024 * 8); cudaMalloc(&a, 16 * 32 * 8); cudaMalloc(&di, 1024);
    cudaMemset(d, 0, 1024 * 8); cudaMemset(a, 0, 16 * 32 * 8); cudaMemset(di, 0, 1024);
    for (int rep = 0; rep < 2; ++rep) {
        k_dadd<<<1, 32>>>(d, -0.25, 1000); cudaDeviceSynchronize();
        k_link<<<1, 32>>>(d, a, -0.25, 1000); cudaDeviceSynchronize();
        k_link_spec<<<1, 32>>>(d, a, -0.25, 1000); cudaDeviceSynchronize();
        k_lds<<<1, 32>>>(d, 1000); cudaDeviceSynchronize();
        k_sel<<<1, 32>>>(di, 1000); cudaDeviceSynchronize();
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
