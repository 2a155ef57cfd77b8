// cycles of the shared-memory Stockham transform (fft_smem.cuh) as the fused PCG kernel uses it: SB = 8 columns, 512 threads.
// nvcc -arch=sm_100a -O3 -I elphdynamics_b200/csrc -I include scripts/micro/fft_stage_bench.cu
#include "fft_smem.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace fftsm;

template <int SB, bool FIXED>
__global__ void k(FftPlan plan, const cplx* twg, long long* cyc, double* sink) {
    extern __shared__ __align__(16) unsigned char raw[];
    cplx* b0 = reinterpret_cast<cplx*>(raw);
    cplx* b1 = b0 + plan.L * SB;
    cplx* tw = b1 + plan.L * SB;
    for (int i = threadIdx.x; i < plan.L; i += blockDim.x) tw[i] = twg[i];
    for (int i = threadIdx.x; i < plan.L * SB; i += blockDim.x) b0[i] = make_double2(i * 0.001, 1.0 - i * 0.002);
    __syncthreads();
    for (int rep = 0; rep < 3; ++rep) {
        long long t0 = clock64();
        cplx* res = FIXED ? fft_smem_auto<SB>(b0, b1, plan, tw, rep & 1) : fft_smem<SB>(b0, b1, plan, tw, rep & 1);
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[rep] = t1 - t0;
        if (res != b0) { cplx* t = b0; b0 = b1; b1 = t; }
        __syncthreads();
    }
    sink[threadIdx.x] = b0[threadIdx.x].x;
}

int main() {
    for (int L : {100, 200, 64, 128}) {
        FftPlan p; p.L = L; p.nrad = 0; int n = L;
        while (n % 4 == 0) { p.rad[p.nrad++] = 4; n /= 4; }
        for (int f = 2; f * f <= n; ++f) while (n % f == 0) { p.rad[p.nrad++] = f; n /= f; }
        if (n > 1) p.rad[p.nrad++] = n;
        std::vector<cplx> tw(L);
        for (int i = 0; i < L; ++i) tw[i] = make_double2(cos(-2 * M_PI * i / L), sin(-2 * M_PI * i / L));
        cplx* dtw; long long* dc; double* ds; long long h[3];
        cudaMalloc(&dtw, L * 16); cudaMemcpy(dtw, tw.data(), L * 16, cudaMemcpyHostToDevice);
        cudaMalloc(&dc, 64); cudaMalloc(&ds, 1024 * 8);
        for (int T : {512, 256}) {
            size_t smem = (2ull * L * 8 + L) * 16;
            cudaFuncSetAttribute(k<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<8, false><<<1, T, smem>>>(p, dtw, dc, ds);
            cudaMemcpy(h, dc, 24, cudaMemcpyDeviceToHost);
            printf("L %d (%d stages) SB 8, %d threads: generic %lld %lld %lld", L, p.nrad, T, h[0], h[1], h[2]);
            k<8, true><<<1, T, smem>>>(p, dtw, dc, ds);
            cudaMemcpy(h, dc, 24, cudaMemcpyDeviceToHost);
            printf("  compile-time plan %lld %lld %lld cycles per transform  (%s)\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
