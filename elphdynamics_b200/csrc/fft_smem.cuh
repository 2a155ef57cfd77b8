// Shared-memory Stockham FFT along imaginary time (device functions shared by fft.cu and pcg_fused.cu).
// Conventions and the mapping of threads to butterflies: see the header comment of fft.cu.
#pragma once

#include "elph_internal.cuh"

namespace fftsm {

constexpr int kMaxRad = 24;

struct FftPlan {
    int L;
    int nrad;
    int rad[kMaxRad];
};

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
// multiply by -i*sg (sg = +1 forward, -1 inverse): (x + iy)(-i sg) = sg*y - i sg*x
__device__ __forceinline__ cplx mul_mi(cplx a, double sg) { return make_double2(sg * a.y, -sg * a.x); }
__device__ __forceinline__ cplx twid(const cplx* tw, int idx, bool inverse) {
    cplx w = tw[idx];
    if (inverse) w.y = -w.y;
    return w;
}

// One Stockham stage of radix R: butterfly (p, q) reads x[q + s(p + m k)], writes y[q + s(R p + j)] * W_L^{p j s}.
template <int SB, int R>
__device__ __forceinline__ void stage_radix(const cplx* __restrict__ x, cplx* __restrict__ y, int L, int n, int s,
                                            const cplx* __restrict__ tw, bool inverse, int site, int slot, int nslots) {
    const int m = n / R;
    const int nb = L / R;  // butterflies per sequence
    const double sg = inverse ? -1.0 : 1.0;
    for (int bb = slot; bb < nb; bb += nslots) {
        const int q = bb % s;
        const int p = bb / s;
        const int base = q + s * p;
        cplx a[R];
#pragma unroll
        for (int k = 0; k < R; ++k) a[k] = x[(size_t)(base + s * m * k) * SB + site];
        cplx b[R];
        if (R == 2) {
            b[0] = cadd(a[0], a[1]);
            b[1] = csub(a[0], a[1]);
        } else if (R == 3) {
            const cplx t = cadd(a[1], a[2]);
            b[0] = cadd(a[0], t);
            const cplx mm = make_double2(a[0].x - 0.5 * t.x, a[0].y - 0.5 * t.y);
            const double h = 0.86602540378443864676;  // sin(2 pi/3)
            const cplx d = csub(a[1], a[2]);
            const cplx nn = mul_mi(make_double2(h * d.x, h * d.y), sg);  // -i sg h (a1 - a2)
            b[1] = cadd(mm, nn);
            b[2] = csub(mm, nn);
        } else if (R == 4) {
            const cplx t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = mul_mi(csub(a[1], a[3]), sg);
            b[0] = cadd(t0, t2);
            b[2] = csub(t0, t2);
            b[1] = cadd(t1, t3);
            b[3] = csub(t1, t3);
        } else {  // R == 5
            const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2pi/5), cos(4pi/5)
            const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin(2pi/5), sin(4pi/5)
            const cplx t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
            b[0] = make_double2(a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y);
            const cplx m1 = make_double2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
            const cplx m2 = make_double2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
            const cplx n1 = mul_mi(make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y), sg);
            const cplx n2 = mul_mi(make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y), sg);
            b[1] = cadd(m1, n1);
            b[4] = csub(m1, n1);
            b[2] = cadd(m2, n2);
            b[3] = csub(m2, n2);
        }
        const int obase = q + s * R * p;
        y[(size_t)obase * SB + site] = b[0];
#pragma unroll
        for (int j = 1; j < R; ++j) {
            // twiddle W_L^{p j s}; p j s <= (m-1)(R-1)s < L, no reduction needed
            y[(size_t)(obase + s * j) * SB + site] = cmul(b[j], twid(tw, p * j * s, inverse));
        }
    }
}

// generic prime radix r: one thread per output element, O(r) work each
template <int SB>
__device__ __forceinline__ void stage_generic(const cplx* __restrict__ x, cplx* __restrict__ y, int L, int n, int s, int r,
                                              const cplx* __restrict__ tw, bool inverse, int site, int slot, int nslots) {
    const int m = n / r;
    const int wstep = L / r;
    for (int e = slot; e < L; e += nslots) {
        const int q = e % s;
        const int pj = e / s;
        const int j = pj % r;
        const int p = pj / r;
        cplx acc = make_double2(0.0, 0.0);
        const int base = q + s * p;
        int widx = 0;
        const int winc = (int)(((long long)j * wstep) % L);
        for (int k = 0; k < r; ++k) {
            acc = cfma(x[(size_t)(base + s * m * k) * SB + site], twid(tw, widx, inverse), acc);
            widx += winc;
            if (widx >= L) widx -= L;
        }
        y[(size_t)e * SB + site] = cmul(acc, twid(tw, p * j * s, inverse));
    }
}

// Stockham FFT of SB interleaved sequences of length L held in shared memory (x, y: [L][SB]).
// Returns the buffer holding the result.  All threads of the CTA must call it.
template <int SB>
__device__ cplx* fft_smem(cplx* x, cplx* y, const FftPlan& plan, const cplx* __restrict__ tw, bool inverse) {
    const int L = plan.L;
    const int site = threadIdx.x % SB;
    const int slot = threadIdx.x / SB;
    const int nslots = blockDim.x / SB;
    int n = L, s = 1;
    for (int st = 0; st < plan.nrad; ++st) {
        const int r = plan.rad[st];
        switch (r) {
            case 2: stage_radix<SB, 2>(x, y, L, n, s, tw, inverse, site, slot, nslots); break;
            case 3: stage_radix<SB, 3>(x, y, L, n, s, tw, inverse, site, slot, nslots); break;
            case 4: stage_radix<SB, 4>(x, y, L, n, s, tw, inverse, site, slot, nslots); break;
            case 5: stage_radix<SB, 5>(x, y, L, n, s, tw, inverse, site, slot, nslots); break;
            default: stage_generic<SB>(x, y, L, n, s, r, tw, inverse, site, slot, nslots); break;
        }
        __syncthreads();
        cplx* tmp = x;
        x = y;
        y = tmp;
        n /= r;
        s *= r;
    }
    return x;
}

// The same transform with the length and the radices known at compile time: the butterfly index arithmetic (two integer
// divisions per butterfly and stage in the generic form), the radix switch and the stage bookkeeping fold into constants.
// Measured on B200 (scripts/micro/fft_stage_bench.cu): ~1250 cycles per stage in the generic form whatever the length.
template <int SB, int L, int N_, int S_>
__device__ __forceinline__ cplx* fft_fixed_stages(cplx* x, cplx* y, const cplx* __restrict__, bool, int, int, int) {
    return x;
}
template <int SB, int L, int N_, int S_, int R, int... Rest>
__device__ __forceinline__ cplx* fft_fixed_stages(cplx* x, cplx* y, const cplx* __restrict__ tw, bool inverse, int site, int slot,
                                                  int nslots) {
    stage_radix<SB, R>(x, y, L, N_, S_, tw, inverse, site, slot, nslots);
    __syncthreads();
    return fft_fixed_stages<SB, L, N_ / R, S_ * R, Rest...>(y, x, tw, inverse, site, slot, nslots);
}
template <int SB, int L, int... Radices>
__device__ __forceinline__ cplx* fft_smem_fixed(cplx* x, cplx* y, const cplx* __restrict__ tw, bool inverse) {
    return fft_fixed_stages<SB, L, L, 1, Radices...>(x, y, tw, inverse, threadIdx.x % SB, threadIdx.x / SB, blockDim.x / SB);
}
// lengths of the shipped examples and of the benchmark configurations, factorised as elph_fft_init does (4 first, then primes)
template <int SB>
__device__ __forceinline__ cplx* fft_smem_auto(cplx* x, cplx* y, const FftPlan& plan, const cplx* __restrict__ tw, bool inverse) {
    switch (plan.L) {
        case 100: return fft_smem_fixed<SB, 100, 4, 5, 5>(x, y, tw, inverse);
        case 200: return fft_smem_fixed<SB, 200, 4, 2, 5, 5>(x, y, tw, inverse);
        case 20: return fft_smem_fixed<SB, 20, 4, 5>(x, y, tw, inverse);
        case 10: return fft_smem_fixed<SB, 10, 2, 5>(x, y, tw, inverse);
        default: return fft_smem<SB>(x, y, plan, tw, inverse);
    }
}

}  // namespace fftsm
