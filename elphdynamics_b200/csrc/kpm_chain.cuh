// Chebyshev chains of the KPM preconditioner on register tiles (shared by kpm_square.cu and pcg_fused.cu).
// Reference: src/KPMPreconditioners.jl:606-693,758-778.
#pragma once

#include "square_tiles.cuh"

namespace kpmch {

using namespace sqt;

struct KsqParams {
    const cplx* __restrict__ in;
    cplx* __restrict__ out;
    const double* __restrict__ eVbar;
    const cplx* __restrict__ coeff;
    const int* __restrict__ order;
    const int* __restrict__ coeff_off;
    const int* __restrict__ schedule;
    const int* skip;
    int L, Ly;
    double inv_mag, avg_over_mag;
    double c0, s0, c1, s1, c2, s2, c3, s3;
    double t0, t1, t2, t3, cprod;   // tanh form (square_tiles.cuh): t_g = s_g / c_g, cprod = c0 c1 c2 c3
    int fast;
    const double2* tab;         // SSH: tau-averaged (cosh, sinh) per bond, tile layout [direction][site]
    unsigned long long* prof;   // development aid (tuning key 12): clock64 stamps of the cluster of the longest polynomial
};

// ---- plain form (any coefficients per colour, or per-bond tables) -------------------------------------------------------
// TAB: per-bond (cosh, sinh) from shared-memory tables in the tile layout [direction][site] (SSH: the tau-averaged hoppings of
// src/KPMPreconditioners.jl:355-381); otherwise one pair per colour (Holstein).
template <int NSEG, int PY, bool TRANSPOSED, bool TAB = false>
__device__ __forceinline__ void apply_A_real(Tile<NSEG, PY>& s, const Tile<NSEG, PY>& ev, const KsqParams& P, double* strips,
                                             int& xbuf, int warp, int nwarps, int lane, const double2* tabs = nullptr) {
    constexpr int LX = 32 * NSEG;
    double ab[NSEG], be[NSEG];
    if (TAB) {
        const int N = LX * P.Ly, y0 = warp * PY;
        const double2* tx = tabs + (size_t)y0 * LX;
        const double2* ty = tabs + N + (size_t)y0 * LX;
        const double2* ty_halo = tabs + N + (size_t)((y0 + P.Ly - 1) % P.Ly) * LX;
        if (!TRANSPOSED) {
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) s.a[r][q] *= ev.a[r][q];
            g0_tab(s, tx, lane);
            g1_tab(s, tx, lane);
            g2_tab(s, ty, lane);
            exchange_edges1(s, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, ab, be);
            xbuf ^= 1;
            g3_tab(s, ty, ty_halo, lane, ab, be);
        } else {
            exchange_edges1(s, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, ab, be);
            xbuf ^= 1;
            g3_tab(s, ty, ty_halo, lane, ab, be);
            g2_tab(s, ty, lane);
            g1_tab(s, tx, lane);
            g0_tab(s, tx, lane);
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) s.a[r][q] *= ev.a[r][q];
        }
        return;
    }
    if (!TRANSPOSED) {
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) s.a[r][q] *= ev.a[r][q];
        g0_x_even(s, P.c0, P.s0);
        g1_x_odd(s, P.c1, P.s1, lane);
        g2_y_even(s, P.c2, P.s2);
        exchange_edges1(s, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, ab, be);
        xbuf ^= 1;
        g3_y_odd(s, P.c3, P.s3, ab, be);
    } else {
        exchange_edges1(s, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, ab, be);
        xbuf ^= 1;
        g3_y_odd(s, P.c3, P.s3, ab, be);
        g2_y_even(s, P.c2, P.s2);
        g1_x_odd(s, P.c1, P.s1, lane);
        g0_x_even(s, P.c0, P.s0);
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) s.a[r][q] *= ev.a[r][q];
    }
}

// A = sum_n cr_n T_n v ,  B = sum_n ci'_n T_n v   (ci' = -ci for the transposed/conjugated pass)
template <int NSEG, int PY, bool TRANSPOSED, bool TAB = false>
__device__ __forceinline__ void poly_real(Tile<NSEG, PY>& A, Tile<NSEG, PY>& B, const Tile<NSEG, PY>& vin, const Tile<NSEG, PY>& ev,
                                          const cplx* c_s, int order, const KsqParams& P, double* strips, int& xbuf, int warp,
                                          int nwarps, int lane, const double2* tabs = nullptr) {
    Tile<NSEG, PY> un, uprev, s;
    const double sg = TRANSPOSED ? -1.0 : 1.0;
    const double c0r = c_s[0].x, c0i = sg * c_s[0].y;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double v = vin.a[r][q];
            A.a[r][q] = c0r * v;
            B.a[r][q] = c0i * v;
            un.a[r][q] = v;
            uprev.a[r][q] = 0.0;
        }
    const double k1 = P.inv_mag, k2 = P.avg_over_mag;
    for (int n = 1; n < order; ++n) {
        s = un;
        apply_A_real<NSEG, PY, TRANSPOSED, TAB>(s, ev, P, strips, xbuf, warp, nwarps, lane, tabs);
        const double cr = c_s[n].x, ci = sg * c_s[n].y;
        const double two = (n > 1) ? 2.0 : 1.0, one = (n > 1) ? 1.0 : 0.0;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                double a = k1 * s.a[r][q] - k2 * un.a[r][q];
                a = two * a - one * uprev.a[r][q];
                uprev.a[r][q] = un.a[r][q];
                un.a[r][q] = a;
                A.a[r][q] += cr * a;
                B.a[r][q] += ci * a;
            }
    }
}


// The same polynomial with the sweep in tanh form and the constants folded (9 / 8 fp64 operations per site and Chebyshev term
// instead of 15; the chain of the lowest frequency is what an apply waits for):
//     evs = 2 (c0 c1 c2 c3 / mag) eVbar ;  Kt = prod_g (1 + t_g X_g)
//     T_1 = (1/2) S v - (avg/mag) v ;  T_n = S T_{n-1} - (2 (avg/mag) T_{n-1} + T_{n-2}) ;  S u = Kt (evs .* u)   [A'^T: evs .* (Kt^T u)]
// The bracket does not depend on the sweep, so it is off the dependent chain.  Differs from poly_real by rounding only.
// WIDE: the slice is split into row strips over the CTAs of a cluster (exchange_edges1_wide; wc = the neighbours' cluster ranks)
template <int NSEG, int PY, bool TRANSPOSED, bool WIDE = false>
__device__ __forceinline__ void sweep_t(Tile<NSEG, PY>& s, const KsqParams& P, double* strips, int& xbuf, int warp, int nwarps,
                                        int lane, WideCtx* wc = nullptr) {
    constexpr int LX = 32 * NSEG;
    double ab[NSEG], be[NSEG];
    auto exchange = [&]() {
        if (WIDE) exchange_edges1_wide(s, strips + (size_t)xbuf * nwarps * 2 * LX, xbuf, warp, nwarps, lane, ab, be, *wc);
        else exchange_edges1(s, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, ab, be);
        xbuf ^= 1;
    };
    if (!TRANSPOSED) {
        g0_x_even_t(s, P.t0);
        g1_x_odd_t(s, P.t1, lane);
        g2_y_even_t(s, P.t2);
        exchange();
        g3_y_odd_t(s, P.t3, ab, be);
    } else {
        exchange();
        g3_y_odd_t(s, P.t3, ab, be);
        g2_y_even_t(s, P.t2);
        g1_x_odd_t(s, P.t1, lane);
        g0_x_even_t(s, P.t0);
    }
}

template <int NSEG, int PY, bool TRANSPOSED, bool WIDE = false>
__device__ __forceinline__ void poly_real_fast(Tile<NSEG, PY>& A, Tile<NSEG, PY>& B, const Tile<NSEG, PY>& vin,
                                               const Tile<NSEG, PY>& evs, const cplx* c_s, int order, const KsqParams& P,
                                               double* strips, int& xbuf, int warp, int nwarps, int lane,
                                               WideCtx* wc = nullptr) {
    Tile<NSEG, PY> un, uprev, s, pre;
    const double sg = TRANSPOSED ? -1.0 : 1.0;
    const double c0r = c_s[0].x, c0i = sg * c_s[0].y;
    const double k2 = P.avg_over_mag, k22 = 2.0 * P.avg_over_mag;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double v = vin.a[r][q];
            A.a[r][q] = c0r * v;
            B.a[r][q] = c0i * v;
            un.a[r][q] = v;
        }
    if (order < 2) return;
    {   // n = 1
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                pre.a[r][q] = k2 * un.a[r][q];
                s.a[r][q] = TRANSPOSED ? un.a[r][q] : (0.5 * evs.a[r][q]) * un.a[r][q];
            }
        sweep_t<NSEG, PY, TRANSPOSED, WIDE>(s, P, strips, xbuf, warp, nwarps, lane, wc);
        const double cr = c_s[1].x, ci = sg * c_s[1].y;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double a = TRANSPOSED ? fma(0.5 * evs.a[r][q], s.a[r][q], -pre.a[r][q]) : (s.a[r][q] - pre.a[r][q]);
                uprev.a[r][q] = un.a[r][q];
                un.a[r][q] = a;
                A.a[r][q] = fma(cr, a, A.a[r][q]);
                B.a[r][q] = fma(ci, a, B.a[r][q]);
            }
    }
    for (int n = 2; n < order; ++n) {
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                pre.a[r][q] = fma(k22, un.a[r][q], uprev.a[r][q]);
                s.a[r][q] = TRANSPOSED ? un.a[r][q] : evs.a[r][q] * un.a[r][q];
            }
        sweep_t<NSEG, PY, TRANSPOSED, WIDE>(s, P, strips, xbuf, warp, nwarps, lane, wc);
        const double cr = c_s[n].x, ci = sg * c_s[n].y;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double a = TRANSPOSED ? fma(evs.a[r][q], s.a[r][q], -pre.a[r][q]) : (s.a[r][q] - pre.a[r][q]);
                uprev.a[r][q] = un.a[r][q];
                un.a[r][q] = a;
                A.a[r][q] = fma(cr, a, A.a[r][q]);
                B.a[r][q] = fma(ci, a, B.a[r][q]);
            }
    }
}

}  // namespace kpmch
