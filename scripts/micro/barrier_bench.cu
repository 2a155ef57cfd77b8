// micro-benchmark: cost of one grid barrier + deterministic sum inside a persistent cooperative kernel
// (development aid for cg_persistent.cu).  Variants:
//   0: barrier only (release-arrive on one counter, one polling thread per CTA)
//   1: + partial store before / ordered read-back of all partials after (what the CG kernels use)
//   2: as 1, each thread also stores NST doubles to global before the barrier and reads a neighbour CTA's after it
//   3: as 2 without the barrier's partial read-back (isolates the halo traffic)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a barrier_bench.cu -o /tmp/barrier_bench && /tmp/barrier_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE, int NST>
__global__ void __launch_bounds__(256) bench(double* partial, unsigned* bar, double* halo, int iters, double* out) {
    __shared__ double bcast;
    const int nb = gridDim.x;
    double acc = threadIdx.x * 1e-3;
    const int up = (blockIdx.x + 1) % nb;
    for (int j = 1; j <= iters; ++j) {
        if (MODE >= 2) {
#pragma unroll
            for (int k = 0; k < NST; ++k) halo[((size_t)blockIdx.x * NST + k) * 256 + threadIdx.x] = acc + k;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (MODE >= 1) partial[blockIdx.x] = acc;
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
            const unsigned target = (unsigned)j * nb;
            while (ld_acquire(bar) < target) {}
        }
        __syncthreads();
        if (MODE == 1 || MODE == 2) {
            if (threadIdx.x < 32) {
                double s = 0.0;
                for (int k = threadIdx.x; k < nb; k += 32) s += __ldcg(partial + k);
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (threadIdx.x == 0) bcast = s;
            }
            __syncthreads();
            acc = acc * 0.5 + bcast * 1e-9;
        }
        if (MODE >= 2) {
#pragma unroll
            for (int k = 0; k < NST; ++k) acc += 1e-9 * __ldcg(halo + ((size_t)up * NST + k) * 256 + threadIdx.x);
        }
    }
    if (acc == 123.456) out[0] = acc;
}

template <int MODE, int NST>
float run(int nb, int iters, double* partial, unsigned* bar, double* halo, double* out, cudaStream_t st) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemsetAsync(bar, 0, 4, st);
        void* args[] = {&partial, &bar, &halo, &iters, &out};
        cudaEventRecord(e0, st);
        cudaLaunchCooperativeKernel((const void*)bench<MODE, NST>, dim3(nb), dim3(256), args, 0, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e3f / iters;
}

int main() {
    double *partial, *halo, *out;
    unsigned* bar;
    cudaMalloc(&partial, 8192);
    cudaMalloc(&halo, 512ull * 8 * 256 * 8);
    cudaMalloc(&out, 64);
    cudaMalloc(&bar, 4);
    cudaStream_t st;
    cudaStreamCreate(&st);
    const int iters = 2000;
    for (int nb : {20, 40, 148, 200, 296}) {
        printf("nb=%3d  barrier %.2f us | +sum %.2f us | +halo(4 st/ld) %.2f us | halo, no sum %.2f us | +halo(8) %.2f us\n", nb,
               run<0, 1>(nb, iters, partial, bar, halo, out, st), run<1, 1>(nb, iters, partial, bar, halo, out, st),
               run<2, 4>(nb, iters, partial, bar, halo, out, st), run<3, 4>(nb, iters, partial, bar, halo, out, st),
               run<2, 8>(nb, iters, partial, bar, halo, out, st));
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
