// micro-benchmark: one-way latency of a word from one SM to a polling thread on another SM through L2, for the access
// flavours the persistent CG kernels could use (development aid for cg_pipe.cu), and the same under contention (many CTAs
// polling the word).  CTA 0 and CTA `peer` bounce a counter; everybody else idles or polls.
//   mode 0: st.volatile / ld.volatile            (what ll_words.cuh uses)
//   mode 1: st.relaxed.gpu / ld.relaxed.gpu
//   mode 2: st.global.cg / ld.global.cg
//   mode 3: atom.exch (relaxed, gpu) / ld.relaxed.gpu
//   mode 4: st.release.gpu / ld.acquire.gpu
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a pingpong_bench.cu -o /tmp/pingpong && /tmp/pingpong
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void put(unsigned* p, unsigned v) {
    if (MODE == 0) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 2) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 3) { unsigned o; asm volatile("atom.relaxed.gpu.global.exch.b32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); }
    if (MODE == 4) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int MODE>
__device__ __forceinline__ unsigned get(const unsigned* p) {
    unsigned v;
    if (MODE == 0) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 1 || MODE == 3) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 4) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// a, b: two words in different 128-byte lines.  CTA 0 writes a = j and waits for b = j; CTA peer waits for a = j and writes
// b = j.  CTAs 1 .. npoll (except peer) poll `a` as well (contention) until the end.
template <int MODE>
__global__ void pingpong(unsigned* a, unsigned* b, int peer, int npoll, int iters, long long* cycles) {
    if (threadIdx.x != 0) return;
    const int c = blockIdx.x;
    if (c == 0) {
        const long long t0 = clock64();
        for (int j = 1; j <= iters; ++j) {
            put<MODE>(a, (unsigned)j);
            { unsigned sp = 0; while (get<MODE>(b) != (unsigned)j && ++sp < (1u << 22)) {} if (sp >= (1u << 22)) { cycles[1] = j; break; } }
        }
        cycles[0] = clock64() - t0;
        put<MODE>(a, 0xffffffffu);
    } else if (c == peer) {
        for (int j = 1; j <= iters; ++j) {
            { unsigned sp = 0; while (get<MODE>(a) != (unsigned)j && ++sp < (1u << 22)) {} if (sp >= (1u << 22)) break; }
            put<MODE>(b, (unsigned)j);
        }
    } else if (c <= npoll) {
        { unsigned sp = 0; while (get<MODE>(a) != 0xffffffffu && ++sp < (1u << 24)) {} }
    }
}

template <int MODE>
void run(const char* name, unsigned* w, long long* cyc) {
    const int iters = 2000;
    for (int npoll : {0, 147, 400}) {
        double lo = 1e30, hi = 0, sum = 0;
        int n = 0;
        for (int peer : {1, 2, 5, 17, 40, 73, 74, 90, 120, 147}) {
            cudaMemset(w, 0, 4096);
            const int grid = (npoll > 147) ? npoll + 1 : 148;
            pingpong<MODE><<<grid, 32>>>(w, w + 64, peer, npoll, iters, cyc);
            long long c2[2];
            cudaMemcpy(c2, cyc, sizeof(c2), cudaMemcpyDeviceToHost);
            const long long c = c2[0];
            if (c2[1]) { printf("%s: timeout at iteration %lld (peer %d, pollers %d)\n", name, c2[1], peer, npoll); fflush(stdout); cudaMemset(cyc, 0, 64); continue; }
            const double one_way = (double)c / iters / 2;
            lo = one_way < lo ? one_way : lo;
            hi = one_way > hi ? one_way : hi;
            sum += one_way;
            ++n;
        }
        printf("%-28s pollers %3d: one-way cycles min %.0f mean %.0f max %.0f\n", name, npoll, lo, sum / (n ? n : 1), hi);
        fflush(stdout);
    }
}

int main() {
    unsigned* w;
    long long* cyc;
    cudaMalloc(&w, 4096);
    cudaMalloc(&cyc, 64);
    cudaMemset(cyc, 0, 64);
    run<0>("st/ld.volatile", w, cyc);
    run<1>("st/ld.relaxed.gpu", w, cyc);
    run<2>("st/ld.cg", w, cyc);
    run<3>("atom.exch / ld.relaxed.gpu", w, cyc);
    run<4>("st.release / ld.acquire", w, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
