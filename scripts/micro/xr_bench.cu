// micro-benchmark: why does cg_xr_kernel take ~8-10 us?  (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double bsum(double x, double* red) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads(); if (l == 0) red[w] = x; __syncthreads();
    double t = 0; if (threadIdx.x == 0) for (int k = 0; k < (blockDim.x >> 5); ++k) t += red[k];
    return t;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double* x, double* r, const double* p, const double* ap, long long n, double* partial, double* S, unsigned* ticket) {
    __shared__ double red[32]; __shared__ bool flag;
    double alpha = S[0]; double sr = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]); double rv = fma(-alpha, ap[i], r[i]); r[i] = rv; sr += rv * rv;
    }
    if (MODE == 0) { if (sr == 123.456) partial[0] = sr; return; }
    double t = bsum(sr, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
    if (MODE == 1) return;
    if (threadIdx.x == 0) { __threadfence(); unsigned nn = atomicAdd(ticket, 1u); flag = (nn == gridDim.x - 1); }
    __syncthreads();
    if (flag) {
        __threadfence();
        double s = 0; for (int q = threadIdx.x; q < gridDim.x; q += blockDim.x) s += ((volatile double*)partial)[q];
        double rr = bsum(s, red);
        if (threadIdx.x == 0) { S[1] = sqrt(rr); S[2] = log(2.0 * rr); *ticket = 0; }
    }
}
int main() {
    long long n = 204800; double *x, *r, *p, *ap, *partial, *S; unsigned* ticket;
    cudaMalloc(&x, n * 8); cudaMalloc(&r, n * 8); cudaMalloc(&p, n * 8); cudaMalloc(&ap, n * 8); cudaMalloc(&partial, 8192); cudaMalloc(&S, 64); cudaMalloc(&ticket, 4);
    cudaMemset(x, 0, n * 8); cudaMemset(r, 0, n * 8); cudaMemset(p, 0, n * 8); cudaMemset(ap, 0, n * 8); cudaMemset(S, 0, 64); cudaMemset(ticket, 0, 4);
    cudaStream_t st; cudaStreamCreate(&st); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148, 296, 800}) for (int mode = 0; mode < 3; ++mode) {
        auto run = [&]() { if (mode == 0) k<0><<<blocks, 256, 0, st>>>(x, r, p, ap, n, partial, S, ticket); else if (mode == 1) k<1><<<blocks, 256, 0, st>>>(x, r, p, ap, n, partial, S, ticket); else k<2><<<blocks, 256, 0, st>>>(x, r, p, ap, n, partial, S, ticket); };
        for (int i = 0; i < 50; ++i) run();
        cudaStreamSynchronize(st); cudaEventRecord(e0, st);
        for (int i = 0; i < 1000; ++i) run();
        cudaEventRecord(e1, st); cudaStreamSynchronize(st); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("blocks %4d mode %d : %.2f us/launch\n", blocks, mode, ms);
    }
    // graph of 16 launches
    cudaGraph_t g; cudaGraphExec_t ge; cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
    for (int i = 0; i < 16; ++i) k<2><<<296, 256, 0, st>>>(x, r, p, ap, n, partial, S, ticket);
    cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st); cudaEventRecord(e0, st);
    for (int i = 0; i < 100; ++i) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st); cudaStreamSynchronize(st); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("graph of 16 x mode2(296): %.2f us/launch\n", ms * 1000 / 1600);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
