// Register-resident KPM apply for periodic square lattices (Holstein, uniform hopping per colour).
//
// Same arithmetic as kpm_apply_kernel in kpm.cu (reference: src/KPMPreconditioners.jl:606-693,758-778):
// for every frequency w <= ceil(L/2):  u = sum_m conj(c_m) T_m(A'^T) v ;  out = sum_m c_m T_m(A') u,
// A' = (A - lam_avg)/lam_mag,  A = K diag(eVbar),  mirror out[L-1-w] = conj(out[w]).
// One CTA per frequency (longest polynomial first).  The complex N-vector lives in registers as a pair of
// real tiles (lane = x, PY rows per warp); a sweep is shuffles + register rotations + ONE __syncthreads for the
// tile-edge rows, instead of ngroups+1 barriers and a shared-memory round trip per bond.  The chain for the
// lowest frequency (2*(order-1) sequential sweeps, order ~ 70 at config B) is latency bound; this kernel cuts
// the latency per sweep by an order of magnitude.
#include "kpm_chain.cuh"
#include "square_tiles.cuh"

namespace {

using namespace sqt;
using namespace kpmch;


template <int NSEG, int PY>
struct CTile {
    Tile<NSEG, PY> re, im;
};

template <int NSEG, int PY, bool TRANSPOSED>
__device__ __forceinline__ void apply_A(CTile<NSEG, PY>& s, const Tile<NSEG, PY>& ev, const KsqParams& P, double* strips, int& xbuf,
                                        int warp, int nwarps, int lane) {
    constexpr int LX = 32 * NSEG;
    double ar[NSEG], ai[NSEG], br[NSEG], bi[NSEG];
    if (!TRANSPOSED) {
        // A s = K (eVbar .* s): g0, g1, g2, g3
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                s.re.a[r][q] *= ev.a[r][q];
                s.im.a[r][q] *= ev.a[r][q];
            }
        g0_x_even(s.re, P.c0, P.s0);
        g0_x_even(s.im, P.c0, P.s0);
        g1_x_odd(s.re, P.c1, P.s1, lane);
        g1_x_odd(s.im, P.c1, P.s1, lane);
        g2_y_even(s.re, P.c2, P.s2);
        g2_y_even(s.im, P.c2, P.s2);
        exchange_edges2(s.re, s.im, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, ar, ai, br, bi);
        xbuf ^= 1;
        g3_y_odd(s.re, P.c3, P.s3, ar, br);
        g3_y_odd(s.im, P.c3, P.s3, ai, bi);
    } else {
        // A^T s = eVbar .* (K^T s): g3, g2, g1, g0
        exchange_edges2(s.re, s.im, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, ar, ai, br, bi);
        xbuf ^= 1;
        g3_y_odd(s.re, P.c3, P.s3, ar, br);
        g3_y_odd(s.im, P.c3, P.s3, ai, bi);
        g2_y_even(s.re, P.c2, P.s2);
        g2_y_even(s.im, P.c2, P.s2);
        g1_x_odd(s.re, P.c1, P.s1, lane);
        g1_x_odd(s.im, P.c1, P.s1, lane);
        g0_x_even(s.re, P.c0, P.s0);
        g0_x_even(s.im, P.c0, P.s0);
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                s.re.a[r][q] *= ev.a[r][q];
                s.im.a[r][q] *= ev.a[r][q];
            }
    }
}

// acc = sum_m c_m T_m(A') vin   (TRANSPOSED: A'^T and conjugated coefficients), three-term recurrence (:625-676)
template <int NSEG, int PY, bool TRANSPOSED>
__device__ __forceinline__ void poly(CTile<NSEG, PY>& acc, const CTile<NSEG, PY>& vin, const Tile<NSEG, PY>& ev, const cplx* c_s,
                                     int order, const KsqParams& P, double* strips, int& xbuf, int warp, int nwarps, int lane) {
    CTile<NSEG, PY> un, uprev, s;
    const double c0r = c_s[0].x, c0i = TRANSPOSED ? -c_s[0].y : c_s[0].y;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double vr = vin.re.a[r][q], vi = vin.im.a[r][q];
            acc.re.a[r][q] = c0r * vr - c0i * vi;
            acc.im.a[r][q] = c0r * vi + c0i * vr;
            un.re.a[r][q] = vr;
            un.im.a[r][q] = vi;
            uprev.re.a[r][q] = 0.0;
            uprev.im.a[r][q] = 0.0;
        }
    for (int n = 1; n < order; ++n) {
        s = un;
        apply_A<NSEG, PY, TRANSPOSED>(s, ev, P, strips, xbuf, warp, nwarps, lane);
        const double cr = c_s[n].x, ci = TRANSPOSED ? -c_s[n].y : c_s[n].y;
        const double two = (n > 1) ? 2.0 : 1.0, one = (n > 1) ? 1.0 : 0.0;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                // A' u = (1/mag) A u - (avg/mag) u ; T_{n+1} = 2 A' T_n - T_{n-1} (first step: T_1 = A' T_0)
                double ar_ = P.inv_mag * s.re.a[r][q] - P.avg_over_mag * un.re.a[r][q];
                double ai_ = P.inv_mag * s.im.a[r][q] - P.avg_over_mag * un.im.a[r][q];
                ar_ = two * ar_ - one * uprev.re.a[r][q];
                ai_ = two * ai_ - one * uprev.im.a[r][q];
                uprev.re.a[r][q] = un.re.a[r][q];
                uprev.im.a[r][q] = un.im.a[r][q];
                un.re.a[r][q] = ar_;
                un.im.a[r][q] = ai_;
                acc.re.a[r][q] += cr * ar_ - ci * ai_;
                acc.im.a[r][q] += cr * ai_ + ci * ar_;
            }
    }
}

template <int NSEG, int PY, int MAXT>
__global__ void __launch_bounds__(MAXT) kpm_square_kernel(KsqParams P, int max_order) {
    constexpr int LX = 32 * NSEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.skip && *P.skip) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int N = LX * P.Ly;
    const int w = P.schedule[blockIdx.x];
    const int order = P.order[w];
    cplx* c_s = reinterpret_cast<cplx*>(smem_raw);                                  // [max_order]
    double* strips = reinterpret_cast<double*>(smem_raw + (size_t)max_order * sizeof(cplx));  // 2 x [nwarps][4][LX]
    for (int k = threadIdx.x; k < order; k += blockDim.x) c_s[k] = P.coeff[P.coeff_off[w] + k];
    const size_t tile_off = (size_t)warp * PY * LX;
    CTile<NSEG, PY> v, t1, t2;
    Tile<NSEG, PY> ev;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = tile_off + r * LX + 32 * q + lane;
            const cplx z = P.in[(size_t)w * N + e];
            v.re.a[r][q] = z.x;
            v.im.a[r][q] = z.y;
            ev.a[r][q] = P.eVbar[e];
        }
    __syncthreads();  // coefficients staged
    int xbuf = 0;
    poly<NSEG, PY, true>(t1, v, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane);    // M^-T[w,w]
    poly<NSEG, PY, false>(t2, t1, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane);  // M^-1[w,w]
    const int wm = P.L - 1 - w;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = tile_off + r * LX + 32 * q + lane;
            if (wm != w) P.out[(size_t)w * N + e] = make_double2(t2.re.a[r][q], t2.im.a[r][q]);
            P.out[(size_t)wm * N + e] = make_double2(t2.re.a[r][q], -t2.im.a[r][q]);
        }
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster-split variant.  A is real, so T_n(A') acts on the real and imaginary parts of the frequency-space vector
// independently; only the coefficient sums mix them:
//     sum_n c_n T_n (vr + i vi):   re = sum cr T vr - sum ci T vi ,   im = sum cr T vi + sum ci T vr .
// A 2-CTA thread-block cluster handles one frequency: CTA 0 runs the chain on vr, CTA 1 on vi, each accumulating
// A = sum cr T v and B = sum ci T v; the B tiles are swapped through distributed shared memory (DSMEM) ONCE per
// polynomial (twice per apply).  The work per sweep on the critical chain (the lowest frequency, ~70 terms at
// config B) is halved.
// ---------------------------------------------------------------------------------------------------------------
template <int NSEG, int PY, int MAXT, bool TAB = false>
__global__ void __launch_bounds__(MAXT) kpm_square_split_kernel(KsqParams P, int max_order) {
    constexpr int LX = 32 * NSEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // NOTE: both CTAs of a cluster take the same early exit (the flag is written before this kernel starts)
    if (P.skip && *P.skip) return;
    unsigned int crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    unsigned long long* prof = (P.prof && (blockIdx.x >> 1) == 0 && threadIdx.x == 0) ? P.prof + 8 * crank : nullptr;
    int nstamp = 0;
    auto stamp = [&]() { if (prof) prof[nstamp++] = (unsigned long long)clock64(); };
    stamp();
    // both CTAs of the pair must be running before either touches the other's shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int N = LX * P.Ly;
    const int w = P.schedule[blockIdx.x >> 1];
    const int order = P.order[w];
    cplx* c_s = reinterpret_cast<cplx*>(smem_raw);                                            // [max_order]
    double* strips = reinterpret_cast<double*>(smem_raw + (size_t)max_order * sizeof(cplx));  // 2 x [nwarps][2][LX]
    double* xch = strips + 2ull * nwarps * 2 * LX;                                            // [N] written by the partner CTA
    double2* tabs = reinterpret_cast<double2*>(xch + N);                                      // TAB: [2][N] (cosh, sinh)
    if (TAB)
        for (int k = threadIdx.x; k < 2 * N; k += blockDim.x) tabs[k] = P.tab[k];
    for (int k = threadIdx.x; k < order; k += blockDim.x) c_s[k] = P.coeff[P.coeff_off[w] + k];
    const size_t tile_off = (size_t)warp * PY * LX;
    Tile<NSEG, PY> v, A, B, ev;
    const double* in_comp = reinterpret_cast<const double*>(P.in) + crank;   // re (CTA 0) or im (CTA 1) of the complex input
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = tile_off + r * LX + 32 * q + lane;
            v.a[r][q] = in_comp[2 * ((size_t)w * N + e)];
            ev.a[r][q] = P.eVbar[e];
        }
    __syncthreads();
    stamp();
    // DSMEM address of the partner's exchange buffer
    const uint32_t my_xch = (uint32_t)__cvta_generic_to_shared(xch);
    uint32_t remote_xch;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_xch) : "r"(my_xch), "r"(crank ^ 1u));
    auto cluster_sync = []() {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    };
    // swap B tiles and combine:  CTA 0 (re): A - B_partner ;  CTA 1 (im): A + B_partner
    auto swap_combine = [&](Tile<NSEG, PY>& out) {
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const uint32_t addr = remote_xch + (uint32_t)((tile_off + r * LX + 32 * q + lane) * sizeof(double));
                asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(B.a[r][q]) : "memory");
            }
        cluster_sync();
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double bp = xch[tile_off + r * LX + 32 * q + lane];
                out.a[r][q] = (crank == 0) ? (A.a[r][q] - bp) : (A.a[r][q] + bp);
            }
        cluster_sync();   // the partner may overwrite xch again only after both have read
    };
    int xbuf = 0;
    Tile<NSEG, PY> t1, t2;
    if (TAB) {
        poly_real<NSEG, PY, true, true>(A, B, v, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane, tabs);    // M^-T[w,w]
        swap_combine(t1);
        poly_real<NSEG, PY, false, true>(A, B, t1, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane, tabs);  // M^-1[w,w]
        swap_combine(t2);
    } else if (P.fast) {
        const double sc = 2.0 * P.inv_mag * P.cprod;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) ev.a[r][q] *= sc;
        poly_real_fast<NSEG, PY, true>(A, B, v, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane);    // M^-T[w,w]
        stamp();
        swap_combine(t1);
        stamp();
        poly_real_fast<NSEG, PY, false>(A, B, t1, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane);  // M^-1[w,w]
        stamp();
        swap_combine(t2);
        stamp();
    } else {
        poly_real<NSEG, PY, true>(A, B, v, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane);    // M^-T[w,w]
        swap_combine(t1);
        poly_real<NSEG, PY, false>(A, B, t1, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane);  // M^-1[w,w]
        swap_combine(t2);
    }
    const int wm = P.L - 1 - w;
    double* out_comp = reinterpret_cast<double*>(P.out) + crank;
    const double msign = (crank == 0) ? 1.0 : -1.0;   // mirror frequency = complex conjugate
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = tile_off + r * LX + 32 * q + lane;
            if (wm != w) out_comp[2 * ((size_t)w * N + e)] = t2.a[r][q];
            out_comp[2 * ((size_t)wm * N + e)] = msign * t2.a[r][q];
        }
    stamp();
}

// ---------------------------------------------------------------------------------------------------------------
// Wide variant for 64-wide lattices: the polynomial order falls as 1/(w + 1/2), so one apply waits for the chain of the lowest
// frequency (2 x 143 dependent sweeps at 64x64xL400) while most SMs idle; a sweep of a 64x64 slice on ONE CTA is bound by that
// SM's fp64 pipe and shuffle unit (~1700 cycles).  Here a frequency takes a cluster of 2 x STRIPS CTAs: (re | im) x STRIPS row strips
// of Ly / STRIPS rows.  Per sweep the first / last rows of a strip travel to the neighbouring strips through distributed shared
// memory and the barrier of the edge exchange becomes the cluster barrier; the B tiles of the (re, im) pair are swapped as in
// the 2-CTA kernel.  Tanh-form sweeps only (Holstein, uniform hopping per colour).
// MEASURED (profiles/r2_kpm_wide_summary.md): 175 us against 178 us for the chain of w = 0 at 64x64xL400 -- no gain.  The arithmetic
// per CTA drops fourfold, but every sweep now waits for a strip-edge round trip through distributed shared memory (~700 cycles
// whether it is a cluster barrier or tagged words polled by the receiver), and with 2 warps per scheduler the dependent
// instruction chain of a sweep is not hidden.  Kept behind tuning key 26 (default off) with its parity test; what would help is
// exchanging k edge rows every k sweeps (redundant halo sweeps), not more CTAs per slice.
// ---------------------------------------------------------------------------------------------------------------
template <int NSEG, int PY, int MAXT, int STRIPS>
__global__ void __launch_bounds__(MAXT) kpm_square_wide_kernel(KsqParams P, int max_order) {
    constexpr int LX = 32 * NSEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.skip && *P.skip) return;   // all CTAs of a cluster take the same exit (the flag is written before this kernel starts)
    unsigned int crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const unsigned int part = crank & 1u, strip = crank >> 1;
    auto cluster_sync = []() {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int N = LX * P.Ly, rows = P.Ly / STRIPS, Nloc = LX * rows;
    const int w = P.schedule[blockIdx.x / (2 * STRIPS)];
    const int order = P.order[w];
    cplx* c_s = reinterpret_cast<cplx*>(smem_raw);                                            // [max_order]
    // (the cluster barrier at the top comes after the zeroing of the halo words below in every CTA: moved there)
    const size_t c_bytes = ((size_t)max_order * sizeof(cplx) + 15) & ~size_t(15);
    ulonglong2* halo = reinterpret_cast<ulonglong2*>(smem_raw + c_bytes);                     // 2 x [2][LX] tagged words
    double* strips = reinterpret_cast<double*>(halo + 4 * LX);                                // 2 x [nwarps][2][LX]
    double* xch = strips + 2ull * nwarps * 2 * LX;                                            // [Nloc] written by the partner CTA
    for (int k = threadIdx.x; k < 4 * LX; k += blockDim.x) halo[k] = make_ulonglong2(0ull, 0ull);
    for (int k = threadIdx.x; k < order; k += blockDim.x) c_s[k] = P.coeff[P.coeff_off[w] + k];
    __syncthreads();
    cluster_sync();   // halo words are zero and every CTA of the cluster is running before anyone writes into another one's memory
    const size_t tile_off = (size_t)warp * PY * LX;                  // within the strip
    const size_t strip_off = (size_t)strip * Nloc;                   // of the strip within the slice
    Tile<NSEG, PY> v, A, B, ev;
    const double* in_comp = reinterpret_cast<const double*>(P.in) + part;   // re (even ranks) or im (odd ranks)
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = strip_off + tile_off + r * LX + 32 * q + lane;
            v.a[r][q] = in_comp[2 * ((size_t)w * N + e)];
            ev.a[r][q] = P.eVbar[e];
        }
    __syncthreads();
    WideCtx wc;
    wc.up_rank = (((strip + STRIPS - 1) % STRIPS) << 1) | part;
    wc.dn_rank = (((strip + 1) % STRIPS) << 1) | part;
    wc.seq = 0ull;
    wc.halo = halo;
    const uint32_t my_xch = (uint32_t)__cvta_generic_to_shared(xch);
    uint32_t remote_xch;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_xch) : "r"(my_xch), "r"(crank ^ 1u));
    // swap B tiles with the (re | im) partner and combine:  re: A - B_partner ;  im: A + B_partner
    auto swap_combine = [&](Tile<NSEG, PY>& out) {
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const uint32_t addr = remote_xch + (uint32_t)((tile_off + r * LX + 32 * q + lane) * sizeof(double));
                asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(B.a[r][q]) : "memory");
            }
        cluster_sync();
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double bp = xch[tile_off + r * LX + 32 * q + lane];
                out.a[r][q] = (part == 0) ? (A.a[r][q] - bp) : (A.a[r][q] + bp);
            }
        cluster_sync();   // the partner may overwrite xch again only after both have read
    };
    int xbuf = 0;
    Tile<NSEG, PY> t1, t2;
    const double sc = 2.0 * P.inv_mag * P.cprod;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) ev.a[r][q] *= sc;
    poly_real_fast<NSEG, PY, true, true>(A, B, v, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane, &wc);    // M^-T[w,w]
    swap_combine(t1);
    poly_real_fast<NSEG, PY, false, true>(A, B, t1, ev, c_s, order, P, strips, xbuf, warp, nwarps, lane, &wc);  // M^-1[w,w]
    swap_combine(t2);
    const int wm = P.L - 1 - w;
    double* out_comp = reinterpret_cast<double*>(P.out) + part;
    const double msign = (part == 0) ? 1.0 : -1.0;   // mirror frequency = complex conjugate
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = strip_off + tile_off + r * LX + 32 * q + lane;
            if (wm != w) out_comp[2 * ((size_t)w * N + e)] = t2.a[r][q];
            out_comp[2 * ((size_t)wm * N + e)] = msign * t2.a[r][q];
        }
}

template <int NSEG, int PY, int MAXT, int STRIPS>
bool launch_ksq_wide(elph_handle* h, const KsqParams& P, int max_order) {
    constexpr int LX = 32 * NSEG;
    const int rows = P.Ly / STRIPS, nwarps = rows / PY;
    if (P.Ly % STRIPS || rows % PY || nwarps < 1 || nwarps * 32 > MAXT) return false;
    size_t smem = (((size_t)max_order * sizeof(cplx) + 15) & ~size_t(15)) + 4ull * LX * sizeof(ulonglong2) +
                  2ull * nwarps * 2 * LX * sizeof(double) + (size_t)LX * rows * sizeof(double);
    if (h->kpm_exclusive) smem = std::max(smem, std::min<size_t>(h->smem_optin, 120 * 1024));   // one chain CTA per SM
    if (smem > h->smem_optin) return false;
    auto kern = kpm_square_wide_kernel<NSEG, PY, MAXT, STRIPS>;
    elph_enable_smem(h, kern);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * STRIPS * h->kpm.nsched);
    cfg.blockDim = dim3(nwarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * STRIPS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        return false;
    }
    ELPH_CUDA(cudaLaunchKernelEx(&cfg, kern, P, max_order));
    h->launches++;
    return true;
}

template <int NSEG, int PY, int MAXT, bool TAB = false>
void launch_ksq_split(elph_handle* h, const KsqParams& P, int nwarps, int max_order) {
    constexpr int LX = 32 * NSEG;
    size_t smem = (size_t)max_order * sizeof(cplx) + 2ull * nwarps * 2 * LX * sizeof(double) +
                  (size_t)LX * P.Ly * sizeof(double) + (TAB ? 2ull * LX * P.Ly * sizeof(double2) : 0);
    // one CTA per SM: a second cluster on the SM of the longest chain would share its fp64 pipe (the chain is what an apply waits for)
    if (h->kpm_exclusive) smem = std::max(smem, std::min<size_t>(h->smem_optin, 120 * 1024));
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "KPM split kernel does not fit in shared memory");
    elph_enable_smem(h, kpm_square_split_kernel<NSEG, PY, MAXT, TAB>);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * h->kpm.nsched);
    cfg.blockDim = dim3(nwarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ELPH_CUDA(cudaLaunchKernelEx(&cfg, kpm_square_split_kernel<NSEG, PY, MAXT, TAB>, P, max_order));
    h->launches++;
}

template <int NSEG, int PY, int MAXT>
void launch_ksq(elph_handle* h, const KsqParams& P, int nwarps, int max_order) {
    constexpr int LX = 32 * NSEG;
    const size_t smem = (size_t)max_order * sizeof(cplx) + 2ull * nwarps * 4 * LX * sizeof(double);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "KPM square kernel: polynomial order too large for shared memory");
    elph_enable_smem(h, kpm_square_kernel<NSEG, PY, MAXT>);
    kpm_square_kernel<NSEG, PY, MAXT><<<h->kpm.nsched, nwarps * 32, smem, h->stream>>>(P, max_order);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace

// Returns false when the model is not served by the square-lattice kernel (caller falls back to kpm_apply_kernel).
bool elph_launch_kpm_square(elph_handle* h, const cplx* nu_in, cplx* nu_out, const int* skip) {
    const KpmState& K = h->kpm;
    if (h->model == ELPH_MODEL_SSH) {
        // SSH on a periodic square lattice 32 sites wide: the same 2-CTA cluster per frequency with the tau-averaged (cosh, sinh)
        // of every bond in shared memory (tile layout, written by taumean2_kernel next to the bond-order copy)
        if (!h->ssq.enabled || h->sq_disable || !K.d_csbar_tile || !h->kpm_split) return false;
        const int Lx = h->ssq.Lx, Ly = h->ssq.Ly;
        if (Lx != 32 || Ly % 2 || Ly / 2 < 2 || Ly / 2 > 16) return false;
        int max_order = 1;
        for (int w = 0; w < K.Lo2; ++w) max_order = std::max(max_order, K.order[w]);
        KsqParams P;
        P.in = nu_in; P.out = nu_out; P.eVbar = K.d_eVbar; P.coeff = K.d_coeff; P.order = K.d_order; P.coeff_off = K.d_coeff_off;
        P.schedule = K.d_schedule; P.skip = skip; P.L = h->L; P.Ly = Ly;
        P.inv_mag = 1.0 / K.lam_mag; P.avg_over_mag = K.lam_avg / K.lam_mag;
        P.c0 = P.c1 = P.c2 = P.c3 = 1.0; P.s0 = P.s1 = P.s2 = P.s3 = 0.0;
        P.t0 = P.t1 = P.t2 = P.t3 = 0.0; P.cprod = 1.0; P.fast = 0;
        P.prof = nullptr;
        P.tab = K.d_csbar_tile;
        launch_ksq_split<1, 2, 512, true>(h, P, Ly / 2, max_order);
        return true;
    }
    if (!h->sq.enabled || h->sq_disable || h->model != ELPH_MODEL_HOLSTEIN) return false;
    const int Lx = h->sq.Lx, Ly = h->sq.Ly;
    // rows per warp: the recurrences are latency bound (one dependent sweep after the other), so on 32-wide lattices the
    // smallest tile wins: 2 rows per warp = 16 warps per CTA at 32x32 (measured 51.2 us per apply against 55.3 with 4)
    int PY = (Lx == 32 && Ly % 2 == 0 && Ly / 2 <= 32) ? 2 : 4;
    if (Lx == 32 && h->sq_py == 8 && Ly % 8 == 0) PY = 8;
    if (Lx == 32 && h->sq_py == 4) PY = 4;
    if (Lx == 32 && h->sq_py == 2) PY = 2;
    if (Ly % PY) return false;
    const int nwarps = Ly / PY;
    if (nwarps < 2 || nwarps > 32) return false;
    int max_order = 1;
    for (int w = 0; w < K.Lo2; ++w) max_order = std::max(max_order, K.order[w]);
    KsqParams P;
    P.in = nu_in; P.out = nu_out; P.eVbar = K.d_eVbar; P.coeff = K.d_coeff; P.order = K.d_order; P.coeff_off = K.d_coeff_off;
    P.schedule = K.d_schedule; P.skip = skip; P.L = h->L; P.Ly = Ly;
    P.inv_mag = 1.0 / K.lam_mag; P.avg_over_mag = K.lam_avg / K.lam_mag;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    P.t0 = P.s0 / P.c0; P.t1 = P.s1 / P.c1; P.t2 = P.s2 / P.c2; P.t3 = P.s3 / P.c3;
    P.cprod = P.c0 * P.c1 * P.c2 * P.c3;
    P.fast = h->kpm_fast ? 1 : 0;
    if (const char* e = getenv("ELPH_KPM_EXCL")) h->kpm_exclusive = atoi(e) != 0;
    P.prof = h->pipe_prof ? h->pipe_prof_buf : nullptr;
    P.tab = nullptr;
    // 64-wide lattices: every frequency on an 8-CTA cluster, (re | im) x 4 row strips (tuning key 26; default off, see below)
    if (Lx == 64 && h->kpm_split && h->kpm_fast && h->kpm_wide && launch_ksq_wide<2, 2, 512, 4>(h, P, max_order)) return true;
#define KSQ_CASE(NS, PYV, MAXT)                                    \
    if (Lx == 32 * NS && PY == PYV && nwarps * 32 <= MAXT) {       \
        if (h->kpm_split) launch_ksq_split<NS, PYV, MAXT>(h, P, nwarps, max_order); \
        else launch_ksq<NS, PYV, MAXT>(h, P, nwarps, max_order);   \
        return true;                                               \
    }
    KSQ_CASE(1, 4, 512)
    KSQ_CASE(1, 8, 256)
    KSQ_CASE(1, 2, 1024)
    KSQ_CASE(2, 4, 512)
    KSQ_CASE(3, 4, 1024)
    KSQ_CASE(4, 4, 1024)
#undef KSQ_CASE
    return false;
}
