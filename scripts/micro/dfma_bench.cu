// fp64 pipe on one SM: dependent-issue latency and throughput of DFMA, latency of a 64-bit shuffle, a __syncthreads round with
// a shared-memory exchange (the ingredients of one Chebyshev sweep in kpm_square.cu).  nvcc -arch=sm_100a -O3 dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dep_chain(double* out, double a, double b, int n, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void thr(double* out, double a, double b, int n, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = out[threadIdx.x] + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_chain(double* out, int n, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x = __shfl_xor_sync(0xffffffffu, x, 1);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_fma_chain(double* out, double a, int n, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x = fma(a, __shfl_xor_sync(0xffffffffu, x, 1), x);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void bar_chain(double* out, int n, long long* cyc) {
    extern __shared__ double sm[];
    double x = out[threadIdx.x];
    const int nb = (threadIdx.x + 32) % blockDim.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        sm[(i & 1) * blockDim.x + threadIdx.x] = x;
        __syncthreads();
        x += sm[(i & 1) * blockDim.x + nb];
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* d; long long* c; long long h[4];
    cudaMalloc(&d, 1024 * 8); cudaMemset(d, 0, 1024 * 8); cudaMalloc(&c, 64);
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        dep_chain<<<1, 32>>>(d, 0.999, 1e-3, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DFMA: %.2f cycles each (1 warp)\n", (double)h[0] / (16.0 * n));
        for (int T : {32, 128, 256, 512, 1024}) {
            thr<8><<<1, T>>>(d, 0.999, 1e-3, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("DFMA throughput, %4d threads x ILP 8: %.2f cycles per warp-instruction per SM-quarter; %.1f lanes/cycle/SM\n", T,
                            (double)h[0] / (8.0 * n) / ((T + 127) / 128), 8.0 * n * T / (double)h[0]);
        }
        thr<2><<<1, 512>>>(d, 0.999, 1e-3, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA, 512 threads x ILP 2: %.1f lanes/cycle/SM\n", 2.0 * n * 512 / (double)h[0]);
        shfl_chain<<<1, 32>>>(d, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent 64-bit shuffle: %.2f cycles\n", (double)h[0] / (8.0 * n));
        shfl_fma_chain<<<1, 32>>>(d, 1e-3, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent shuffle + DFMA: %.2f cycles\n", (double)h[0] / (8.0 * n));
        for (int T : {128, 256, 512, 1024}) {
            bar_chain<<<1, T, 2 * T * 8>>>(d, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("STS + __syncthreads + LDS + DADD round, %4d threads: %.1f cycles\n", T, (double)h[0] / n);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
