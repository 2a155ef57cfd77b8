// K5: KPM (Chebyshev) preconditioner for M^T M.
//
// Reference: src/KPMPreconditioners.jl
//   KPMExpansion :101-146, setup! :269-321, update_A! :332-381, expansion mul!/ldiv! :387-420,
//   apply ldiv! :426-481, SymmetricKPMPreconditioner mul! :606-679, mulA'! :685-693, mulA! :758-778,
//   kpm_coefficients! :789-839, arnoldi_eigenvalue_bounds! :845-942, scalar_invM :948-951.
//
// Split of work:
//  * tau-mean of the operator tables: device kernel (sequential-in-tau sum per site, like the reference loop).
//  * Arnoldi bounds (2 x <=20 single-slice products on an N-vector, Gram-Schmidt, eigenvalues of a <=20x20
//    Hessenberg matrix) and the Chebyshev coefficient tables: host, once per setup!.  The start vectors are
//    injected by the caller (the reference draws them from model.rng).
//  * apply: twisted FFT (fft.cu) -> one CTA per Matsubara-like frequency runs the two three-term recurrences
//    on an N-vector (thread-owned sites in registers, the checkerboard sweep in shared memory), longest
//    polynomial first -> inverse FFT.  In the engine layout the frequency-space vector is [omega][site], so
//    the reference's two (L,N)<->(N,L) transposes disappear.
#include "elph_internal.cuh"
#include "square_tiles.cuh"

#include <future>
#include <algorithm>
#include <cmath>
#include <numeric>

typedef std::complex<double> zc;

// ------------------------------------------------------------------------------------------------
// host: eigenvalues of a small real upper-Hessenberg matrix (replaces LAPACK eigvals!, :891,:935)
// Complex single-shift (Wilkinson) QR with Givens rotations and deflation.
// ------------------------------------------------------------------------------------------------
// Plain (re, im) arithmetic: std::complex products compile to __muldc3 calls with NaN/Inf recovery, which made the two
// 20 x 20 problems of a set-up cost 2 x 55 us.  Only eigenvalues are wanted, so every rotation is applied to the active
// window [lo, hi] alone.
namespace {
struct cz {
    double re, im;
};
inline cz operator+(cz a, cz b) { return {a.re + b.re, a.im + b.im}; }
inline cz operator-(cz a, cz b) { return {a.re - b.re, a.im - b.im}; }
inline cz operator*(cz a, cz b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline cz operator*(double a, cz b) { return {a * b.re, a * b.im}; }
inline cz conjz(cz a) { return {a.re, -a.im}; }
inline cz negz(cz a) { return {-a.re, -a.im}; }
inline double norm2(cz a) { return a.re * a.re + a.im * a.im; }
inline double absz(cz a) { return std::sqrt(a.re * a.re + a.im * a.im); }   // entries are O(1): no hypot needed
inline cz sqrtz(cz a) {
    const zc r = std::sqrt(zc(a.re, a.im));
    return {r.real(), r.imag()};
}
}  // namespace

static std::vector<zc> hessenberg_eigvals(const std::vector<double>& hreal, int n) {
    std::vector<cz> H((size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H[(size_t)i * n + j] = cz{(i <= j + 1) ? hreal[(size_t)i * n + j] : 0.0, 0.0};
    auto at = [&](int i, int j) -> cz& { return H[(size_t)i * n + j]; };
    std::vector<zc> ev(n);
    std::vector<cz> cs(n), sn(n);
    double hnorm = 0.0;
    for (auto& z : H) hnorm += norm2(z);
    hnorm = std::sqrt(hnorm);
    if (hnorm == 0.0) hnorm = 1.0;
    const double eps = 2.220446049250313e-16;
    int hi = n - 1;
    int iter = 0;
    while (hi >= 0) {
        if (hi == 0) {
            ev[0] = zc(at(0, 0).re, at(0, 0).im);
            break;
        }
        // look for a negligible sub-diagonal entry
        int lo = hi;
        while (lo > 0) {
            double s = absz(at(lo - 1, lo - 1)) + absz(at(lo, lo));
            if (s == 0.0) s = hnorm;
            if (absz(at(lo, lo - 1)) <= eps * s) {
                at(lo, lo - 1) = cz{0.0, 0.0};
                break;
            }
            --lo;
        }
        if (lo == hi) {
            ev[hi] = zc(at(hi, hi).re, at(hi, hi).im);
            --hi;
            iter = 0;
            continue;
        }
        // Wilkinson shift from the trailing 2x2 block
        const cz a = at(hi - 1, hi - 1), b = at(hi - 1, hi), c = at(hi, hi - 1), d = at(hi, hi);
        const cz tr = a + d, det = a * d - b * c;
        const cz disc = sqrtz(tr * tr - 4.0 * det);
        const cz l1 = 0.5 * (tr + disc), l2 = 0.5 * (tr - disc);
        cz mu = (absz(l1 - d) < absz(l2 - d)) ? l1 : l2;
        ++iter;
        if (iter % 11 == 10) mu = cz{absz(at(hi, hi - 1)) + absz(at(hi - 1, hi - 2 >= lo ? hi - 2 : lo)), 0.0};  // exceptional shift
        if (iter > 30 * n + 300) break;  // give up (values so far are returned; caller treats NaN as inactive)
        // QR step on the active block [lo, hi]
        for (int k = lo; k <= hi; ++k) at(k, k) = at(k, k) - mu;
        for (int k = lo; k < hi; ++k) {
            const cz x = at(k, k), y = at(k + 1, k);
            const double r = std::sqrt(norm2(x) + norm2(y));
            cz cc, ss;
            if (r == 0.0) {
                cc = cz{1.0, 0.0};
                ss = cz{0.0, 0.0};
            } else {
                cc = (1.0 / r) * x;
                ss = (1.0 / r) * y;
            }
            cs[k - lo] = cc;
            sn[k - lo] = ss;
            // rows k, k+1:  [ conj(c) conj(s); -s c ]
            for (int j = k; j <= hi; ++j) {
                const cz t1 = at(k, j), t2 = at(k + 1, j);
                at(k, j) = conjz(cc) * t1 + conjz(ss) * t2;
                at(k + 1, j) = cc * t2 - ss * t1;
            }
        }
        for (int k = lo; k < hi; ++k) {
            const cz cc = cs[k - lo], ss = sn[k - lo];
            const int top = std::min(k + 2, hi);
            for (int i = lo; i <= top; ++i) {
                const cz t1 = at(i, k), t2 = at(i, k + 1);
                at(i, k) = t1 * cc + t2 * ss;
                at(i, k + 1) = t2 * conjz(cc) - t1 * conjz(ss);
            }
        }
        for (int k = lo; k <= hi; ++k) at(k, k) = at(k, k) + mu;
    }
    return ev;
}

std::vector<zc> elph_debug_hess_eig(const std::vector<double>& h, int n) { return hessenberg_eigvals(h, n); }

// ------------------------------------------------------------------------------------------------
// host: single-slice products with the tau-averaged operator (for Arnoldi only)
// ------------------------------------------------------------------------------------------------
static void host_mulA(const elph_handle* h, const KpmState& K, const std::vector<double>& v, std::vector<double>& out) {
    const int N = h->N;
    for (int i = 0; i < N; ++i) out[i] = K.eVbar[i] * v[i];
    for (int n = 0; n < h->Nb; ++n) {  // checkerboard_mul!, bond 1 first (src/Checkerboard.jl:123-141)
        const int i = h->bonds_host[n].x, j = h->bonds_host[n].y;
        const double c = K.cbar[n], s = K.sbar[n];
        const double t1 = out[i], t2 = out[j];
        out[i] = c * t1 + s * t2;
        out[j] = c * t2 + s * t1;
    }
}

static void host_ldivA(const elph_handle* h, const KpmState& K, const std::vector<double>& v, std::vector<double>& out) {
    const int N = h->N;
    out = v;
    for (int n = h->Nb - 1; n >= 0; --n) {  // checkerboard_inverse_mul! (src/Checkerboard.jl:298-316)
        const int i = h->bonds_host[n].x, j = h->bonds_host[n].y;
        const double c = K.cbar[n], s = K.sbar[n];
        const double t1 = out[i], t2 = out[j];
        out[i] = c * t1 - s * t2;
        out[j] = c * t2 - s * t1;
    }
    for (int i = 0; i < N; ++i) out[i] /= K.eVbar[i];
}

// dot product with four independent accumulators (the strict left-to-right sum does not vectorise; the Arnoldi bounds
// only enter through isapprox(rtol = buf) and floor() of the polynomial orders, far above this rounding difference)
static double host_dot(const double* a, const double* b, int n) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = 0;
    for (; i + 3 < n; i += 4) {
        s0 += a[i] * b[i];
        s1 += a[i + 1] * b[i + 1];
        s2 += a[i + 2] * b[i + 2];
        s3 += a[i + 3] * b[i + 3];
    }
    for (; i < n; ++i) s0 += a[i] * b[i];
    return (s0 + s1) + (s2 + s3);
}

// one half of arnoldi_eigenvalue_bounds! (:845-942): returns max real eigenvalue of the projected operator
static double host_arnoldi(const elph_handle* h, const KpmState& K, const double* start, bool inverse) {
    const int N = h->N, n = K.n;
    std::vector<double> Q((size_t)(n + 1) * N, 0.0), hm((size_t)(n + 1) * n, 0.0), b(N), v(N);
    double nb = 0.0;
    for (int i = 0; i < N; ++i) nb += start[i] * start[i];
    nb = std::sqrt(nb);
    for (int i = 0; i < N; ++i) b[i] = start[i] / nb;
    std::copy(b.begin(), b.end(), Q.begin());
    int l = n;
    for (int k = 0; k < n; ++k) {
        if (inverse) host_ldivA(h, K, b, v); else host_mulA(h, K, b, v);
        for (int j = 0; j <= k; ++j) {
            const double* Qj = &Q[(size_t)j * N];
            const double d = host_dot(Qj, v.data(), N);
            hm[(size_t)j * n + k] = d;
            for (int i = 0; i < N; ++i) v[i] -= d * Qj[i];
        }
        const double nv = std::sqrt(host_dot(v.data(), v.data(), N));
        hm[(size_t)(k + 1) * n + k] = nv;
        if (nv > 1e-12) {
            for (int i = 0; i < N; ++i) b[i] = v[i] / nv;
            std::copy(b.begin(), b.end(), Q.begin() + (size_t)(k + 1) * N);
        } else {
            l = k + 1;
            break;
        }
    }
    std::vector<double> hh((size_t)l * l);
    bool finite = true;
    for (int i = 0; i < l; ++i)
        for (int j = 0; j < l; ++j) {
            hh[(size_t)i * l + j] = hm[(size_t)i * n + j];
            if (!std::isfinite(hh[(size_t)i * l + j])) finite = false;
        }
    if (!finite) return INFINITY;
    std::vector<zc> ev = hessenberg_eigvals(hh, l);
    double mx = -INFINITY;
    for (auto& e : ev) mx = std::max(mx, e.real());
    return mx;
}

// largest real part of the eigenvalues of the leading l x l block of the Arnoldi Hessenberg matrix (leading dimension n)
static double hessenberg_bound(const double* hm, int n, int l) {
    std::vector<double> hh((size_t)l * l);
    for (int i = 0; i < l; ++i)
        for (int j = 0; j < l; ++j) {
            hh[(size_t)i * l + j] = hm[(size_t)i * n + j];
            if (!std::isfinite(hh[(size_t)i * l + j])) return INFINITY;
        }
    std::vector<zc> ev = hessenberg_eigvals(hh, l);
    double mx = -INFINITY;
    for (auto& e : ev) mx = std::max(mx, e.real());
    return mx;
}

// kpm_coefficients! (:789-839): c_0 = S_0/(2M), c_m = 2 S_m/(2M), S_m = sum_n f(x_n) cos(pi m (n+1/2)/(2M))
static void host_kpm_coefficients(zc* c, int order, double lam_lo, double lam_hi, double phi) {
    const int M = order, NM = 2 * M;
    const double lam_avg = (lam_hi + lam_lo) / 2, lam_mag = (lam_hi - lam_lo) / 2;
    const double pi = 3.14159265358979323846;
    std::vector<zc> f(NM);
    const zc eph = std::exp(zc(0.0, -phi));
    for (int n = 0; n < NM; ++n) {
        const double xn = lam_mag * std::cos(pi * (n + 0.5) / NM) + lam_avg;
        f[n] = 1.0 / (1.0 - eph * xn);
    }
    for (int m = 0; m < M; ++m) {
        zc S(0.0, 0.0);
        for (int n = 0; n < NM; ++n) S += f[n] * std::cos(pi * m * (n + 0.5) / NM);
        c[m] = (m == 0) ? S / (double)NM : 2.0 * S / (double)NM;
    }
}

static bool isapprox_rtol(double x, double y, double rtol) { return x == y || std::fabs(x - y) <= rtol * std::max(std::fabs(x), std::fabs(y)); }

// ------------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------------
namespace {

// Holstein: eVbar[i] = (sum_tau expnV[tau][i]) / L, summed in tau order (:340-347)
__global__ void taumean_kernel(const double* __restrict__ tab, double* __restrict__ out, int ncols, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols) return;
    double s = 0.0;
    for (int t = 0; t < L; ++t) s += tab[(size_t)t * ncols + i];
    out[i] = s / (double)L;
}
// SSH: (cbar,sbar)[b] = mean_tau (cosh,sinh)[tau][b] (:367-376)
// tile_out (optional): second copy in the tile layout [direction][site] of ssh_square.cu (slot = position of the bond there)
__global__ void taumean2_kernel(const double2* __restrict__ tab, double2* __restrict__ out, int ncols, int L,
                                const int* __restrict__ slot = nullptr, double2* __restrict__ tile_out = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols) return;
    double sc = 0.0, ss = 0.0;
    for (int t = 0; t < L; ++t) {
        const double2 v = tab[(size_t)t * ncols + i];
        sc += v.x;
        ss += v.y;
    }
    out[i] = make_double2(sc / (double)L, ss / (double)L);
    if (tile_out) tile_out[slot[i]] = out[i];
}

// SSH set-up from an external tau-mean (tau-sharded lattice): copy of the (cbar, sbar) pairs in the tile layout of ssh_square.cu
__global__ void csbar_tile_kernel(const double2* __restrict__ in, const int* __restrict__ slot, double2* __restrict__ tile_out, int ncols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncols) tile_out[slot[i]] = in[i];
}

// arnoldi_eigenvalue_bounds! (src/KPMPreconditioners.jl:845-942) on the device: blockIdx.x = 0 runs the Krylov iteration on A
// (e_max), 1 on A^-1 (1/e_min).  Output per run: the (n+1) x n Hessenberg matrix (row-major, leading dimension n) and the
// number of completed steps; the 20 x 20 eigenvalue problem stays on the host.
//
// Latency is everything here (two CTAs, one dependent chain each).  Measured on B200 with the reference's loop structure
// (modified Gram-Schmidt: k+1 dependent block reductions in step k; products through the bond list): 1100 cycles per
// reduction and 5300 per product, 190 us per kernel -- no better than the host loops it replaced.  Hence:
//   * a thread keeps its EPT elements of the current vector in registers for the whole iteration;
//   * orthogonalisation = classical Gram-Schmidt, applied twice: all k+1 dots of a pass are formed at once and folded by ONE
//     block reduction, h(j,k) = the sum of the two passes.  Same Krylov space, same projected matrix in exact arithmetic and
//     orthogonality at rounding level like the modified form (Giraud et al. 2005, "twice is enough"); the eigenvalue bounds
//     agree with the sequential form to ~1e-10 and only enter through isapprox(rtol = buf) and floor() of the orders;
//   * on periodic square lattices the products are register-tile sweeps (square_tiles.cuh), otherwise bond-list sweeps on a
//     copy in shared memory.
struct ArnoldiParams {
    const double* __restrict__ eVbar;
    const double2* __restrict__ csbar;
    const int2* __restrict__ bonds;
    const int* __restrict__ goff;
    const double* __restrict__ start;   // [2][N]
    double* Qg;                         // [2][n+1][N] when the basis is kept in global memory
    double* hm;                         // [2][(n+1) * n + 1]
    int ngroups, N, n, q_in_smem, Ly;
    double c[4], s[4];                  // square lattices: (cosh, sinh) per colour
    unsigned long long* prof;
};

constexpr int kArnMaxN = 32;            // Krylov steps the kernel is sized for (the reference default is 20)
constexpr int kArnWarps = 8;            // at most 8 warps per CTA
constexpr int kArnStride = kArnMaxN + 4;   // partial sums per warp: one per basis vector, padded to whole chunks of four

// fold nval per-thread values over the block: every thread receives the nval sums (same bits everywhere) in out[]
template <int MAXV>
__device__ __forceinline__ void arnoldi_block_sums(double (&val)[MAXV], int nval, double* red, int& rbuf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
        if (j < nval) {
            double x = val[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            val[j] = x;
        }
    double* mine = red + (size_t)rbuf * 32 * MAXV;
    rbuf ^= 1;
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < MAXV; ++j)
            if (j < nval) mine[warp * MAXV + j] = val[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
        if (j < nval) {
            double t = 0.0;
            for (int w = 0; w < nw; ++w) t += mine[w * MAXV + j];
            val[j] = t;
        }
}

// QSMEM: the Krylov basis lives in shared memory (addressed as such: generic loads of it were what the first versions waited for)
template <int EPT, bool SQUARE, bool QSMEM>
__global__ void __launch_bounds__(256) arnoldi_kernel(ArnoldiParams P) {
    extern __shared__ __align__(16) double sm[];
    const int N = P.N, n = P.n, T = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = T >> 5;
    const bool inverse = (blockIdx.x == 1);
    // shared memory: [reduction buffers 2 x kArnWarps x kArnStride] [vs: N (bond-list sweeps) or edge strips (tiles)] [Q: n + 4 rows]
    double* red = sm;
    double* vs = red + 2 * kArnWarps * kArnStride;
    double* Qs = vs + (SQUARE ? 2 * nwarps * 2 * 32 : N);
    double* Q = QSMEM ? Qs : P.Qg + (size_t)blockIdx.x * (n + 4) * N;   // [n+4][N], rows beyond the current basis are zero
    // two matrices per run: the coefficients of the first and of the second orthogonalisation pass (the host adds them)
    const size_t hstride = (size_t)(n + 1) * n + 1;
    double* hm = P.hm + (size_t)blockIdx.x * 2 * hstride;
    const double* start = P.start + (size_t)blockIdx.x * N;
    __shared__ double red1[64];
    int rbuf = 0, rbuf1 = 0, xbuf = 0;
    for (int i = threadIdx.x; i < 2 * (int)hstride; i += T) hm[i] = 0.0;
    for (int i = threadIdx.x; i < 2 * kArnWarps * kArnStride; i += T) red[i] = 0.0;
    for (int i = threadIdx.x; i < (n + 4) * N; i += T) Q[i] = 0.0;
    __syncthreads();
    // element e of this thread: square lattices: site (EPT * warp + e, lane) -- a register tile; otherwise tid + e * T
    auto idx = [&](int e) { return SQUARE ? (EPT * warp + e) * 32 + lane : (int)threadIdx.x + e * T; };
    double v[EPT], ev[EPT];
    double one[1];
    one[0] = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int i = idx(e);
        v[e] = (i < N) ? start[i] : 0.0;
        ev[e] = (i < N) ? P.eVbar[i] : 1.0;
        if (inverse) ev[e] = 1.0 / ev[e];            // A^-1 multiplies by 1/eVbar: one division per element and kernel, not per step
        one[0] = fma(v[e], v[e], one[0]);
    }
    arnoldi_block_sums<1>(one, 1, red1, rbuf1);
    const double rnb = 1.0 / sqrt(one[0]);
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int i = idx(e);
        v[e] *= rnb;                                  // v = q_0
        if (i < N) Q[i] = v[e];
    }
    int l = n;
    long long c_mul = 0, c_gs = 0, c_t = clock64();
    for (int k = 0; k < n; ++k) {
        // v holds q_k.  v = A q_k (A = K diag(eVbar)) or A^-1 q_k
        if (SQUARE) {
            sqt::Tile<1, EPT> t;
            double ab[1], be[1];
            if (!inverse) {   // checkerboard_mul!: colours in order (src/Checkerboard.jl:123-141)
#pragma unroll
                for (int e = 0; e < EPT; ++e) t.a[e][0] = ev[e] * v[e];
                sqt::g0_x_even(t, P.c[0], P.s[0]);
                sqt::g1_x_odd(t, P.c[1], P.s[1], lane);
                sqt::g2_y_even(t, P.c[2], P.s[2]);
                sqt::exchange_edges1(t, vs + (size_t)xbuf * nwarps * 2 * 32, warp, nwarps, lane, ab, be);
                xbuf ^= 1;
                sqt::g3_y_odd(t, P.c[3], P.s[3], ab, be);
#pragma unroll
                for (int e = 0; e < EPT; ++e) v[e] = t.a[e][0];
            } else {          // checkerboard_inverse_mul!: reverse order, sinh -> -sinh (src/Checkerboard.jl:298-316), then 1/eVbar
#pragma unroll
                for (int e = 0; e < EPT; ++e) t.a[e][0] = v[e];
                sqt::exchange_edges1(t, vs + (size_t)xbuf * nwarps * 2 * 32, warp, nwarps, lane, ab, be);
                xbuf ^= 1;
                sqt::g3_y_odd(t, P.c[3], -P.s[3], ab, be);
                sqt::g2_y_even(t, P.c[2], -P.s[2]);
                sqt::g1_x_odd(t, P.c[1], -P.s[1], lane);
                sqt::g0_x_even(t, P.c[0], -P.s[0]);
#pragma unroll
                for (int e = 0; e < EPT; ++e) v[e] = t.a[e][0] * ev[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = idx(e);
                if (i < N) vs[i] = inverse ? v[e] : ev[e] * v[e];
            }
            __syncthreads();
            for (int gg = 0; gg < P.ngroups; ++gg) {
                const int g = inverse ? P.ngroups - 1 - gg : gg;
                for (int b = P.goff[g] + threadIdx.x; b < P.goff[g + 1]; b += T) {
                    const int2 ij = P.bonds[b];
                    const double2 cs = P.csbar[b];
                    const double sn = inverse ? -cs.y : cs.y;
                    const double t1 = vs[ij.x], t2 = vs[ij.y];
                    vs[ij.x] = cs.x * t1 + sn * t2;
                    vs[ij.y] = cs.x * t2 + sn * t1;
                }
                __syncthreads();
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = idx(e);
                v[e] = (i < N) ? (inverse ? vs[i] * ev[e] : vs[i]) : 0.0;
            }
        }
        { const long long t = clock64(); c_mul += t - c_t; c_t = t; }
        // two passes of classical Gram-Schmidt against q_0 .. q_k
        // Chunks of four basis vectors, branch-free: the basis has n + 4 rows and the rows that are not written yet are zero, so
        // a chunk that runs past q_k computes zeros; the partial sums of at most kArnWarps warps are folded with a fixed
        // trip count (unused slots stay zero).  (A fully unrolled, predicated 33-wide body cost 25 k cycles per step in
        // instruction issue alone; predicated chunks serialised the four dots.)
        for (int pass = 0; pass < 2; ++pass) {
            double* mine = red + (size_t)rbuf * kArnWarps * kArnStride;
            rbuf ^= 1;
            for (int j0 = 0; j0 <= k; j0 += 4) {
                double d[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const double* Qj = Q + (size_t)(j0 + jj) * N;
#pragma unroll
                    for (int e = 0; e < EPT; e += 2) {
                        const int i0 = idx(e), i1 = idx(e + 1);
                        if (SQUARE || i0 < N) d[jj] = fma(Qj[i0], v[e], d[jj]);
                        if (SQUARE || i1 < N) d2[jj] = fma(Qj[i1], v[e + 1], d2[jj]);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) d[jj] += d2[jj];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) d[jj] += __shfl_xor_sync(0xffffffffu, d[jj], o);
                if (lane == 0)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) mine[warp * kArnStride + j0 + jj] = d[jj];
            }
            __syncthreads();
            for (int j0 = 0; j0 <= k; j0 += 4) {
                double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int w = 0; w < kArnWarps; ++w)       // same order in every thread
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) t[jj] += mine[w * kArnStride + j0 + jj];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const double* Qj = Q + (size_t)(j0 + jj) * N;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        const int i = idx(e);
                        if (SQUARE || i < N) v[e] = fma(-t[jj], Qj[i], v[e]);
                    }
                    if (threadIdx.x == 0 && j0 + jj <= k) hm[(size_t)pass * hstride + (size_t)(j0 + jj) * n + k] = t[jj];
                }
            }
        }
        one[0] = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) one[0] = fma(v[e], v[e], one[0]);
        arnoldi_block_sums<1>(one, 1, red1, rbuf1);
        const double nv = sqrt(one[0]);
        if (threadIdx.x == 0) hm[(size_t)(k + 1) * n + k] = nv;
        if (!(nv > 1e-12)) {
            l = k + 1;
            break;
        }
        double* Qn = Q + (size_t)(k + 1) * N;
        const double rnv = 1.0 / nv;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = idx(e);
            v[e] *= rnv;
            if (i < N) Qn[i] = v[e];      // read back by the same thread only
        }
        { const long long t = clock64(); c_gs += t - c_t; c_t = t; }
    }
    if (threadIdx.x == 0) hm[hstride - 1] = (double)l;
    if (threadIdx.x == 0 && P.prof) { P.prof[2 * blockIdx.x] = c_mul; P.prof[2 * blockIdx.x + 1] = c_gs; }
}

struct KpmParams {
    const cplx* __restrict__ in;   // [L][N] frequency-space input
    cplx* __restrict__ out;        // [L][N]
    const double* __restrict__ eVbar;
    const double2* __restrict__ csbar;
    const int2* __restrict__ bonds;
    const int* __restrict__ goff;
    const cplx* __restrict__ coeff;
    const int* __restrict__ order;
    const int* __restrict__ coeff_off;
    const int* __restrict__ schedule;
    const int* skip;
    int ngroups, N, L;
    double inv_mag, avg_over_mag;
};

__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

template <bool REVERSE>
__device__ __forceinline__ void sweep_cplx(cplx* __restrict__ sw, const KpmParams& P) {
    for (int gg = 0; gg < P.ngroups; ++gg) {
        const int g = REVERSE ? (P.ngroups - 1 - gg) : gg;
        const int lo = P.goff[g], hi = P.goff[g + 1];
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            const int2 ij = P.bonds[b];
            const double2 cs = P.csbar[b];
            const cplx t1 = sw[ij.x], t2 = sw[ij.y];
            sw[ij.x] = make_double2(cs.x * t1.x + cs.y * t2.x, cs.x * t1.y + cs.y * t2.y);
            sw[ij.y] = make_double2(cs.x * t2.x + cs.y * t1.x, cs.x * t2.y + cs.y * t1.y);
        }
        __syncthreads();
    }
}

// sum_m c_m T_m(A') v  (TRANSPOSED: A'^T and conjugated coefficients)  -- :625-676
template <int SPT, bool TRANSPOSED>
__device__ __forceinline__ void poly(cplx (&acc)[SPT], const cplx (&vin)[SPT], cplx* __restrict__ sw, const cplx* __restrict__ c,
                                     int order, const KpmParams& P, const double (&ev)[SPT]) {
    cplx uprev[SPT], un[SPT];
    cplx c0 = c[0];
    if (TRANSPOSED) c0.y = -c0.y;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        acc[k] = cmulc(c0, vin[k]);
        un[k] = vin[k];
        uprev[k] = make_double2(0.0, 0.0);
    }
    for (int n = 1; n < order; ++n) {
        // sw = A u_n  or  A^T u_n
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < P.N) sw[i] = TRANSPOSED ? un[k] : make_double2(ev[k] * un[k].x, ev[k] * un[k].y);
        }
        __syncthreads();
        sweep_cplx<TRANSPOSED>(sw, P);
        cplx cn = c[n];
        if (TRANSPOSED) cn.y = -cn.y;
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            cplx t = (i < P.N) ? sw[i] : make_double2(0.0, 0.0);
            if (TRANSPOSED) t = make_double2(ev[k] * t.x, ev[k] * t.y);
            // A' u = (1/mag) A u - (avg/mag) u
            cplx a = make_double2(P.inv_mag * t.x - P.avg_over_mag * un[k].x, P.inv_mag * t.y - P.avg_over_mag * un[k].y);
            if (n > 1) a = make_double2(2.0 * a.x - uprev[k].x, 2.0 * a.y - uprev[k].y);
            uprev[k] = un[k];
            un[k] = a;
            const cplx ca = cmulc(cn, a);
            acc[k].x += ca.x;
            acc[k].y += ca.y;
        }
        // the next iteration's write to sw[i] is by the owning thread; the barrier after it orders it against the sweep
    }
}

template <int SPT>
__global__ void __launch_bounds__(512) kpm_apply_kernel(KpmParams P) {
    extern __shared__ double smem_raw[];
    cplx* sw = reinterpret_cast<cplx*>(smem_raw);
    if (P.skip && *P.skip) return;
    const int w = P.schedule[blockIdx.x];
    const int order = P.order[w];
    const cplx* c = P.coeff + P.coeff_off[w];
    cplx v[SPT], t1[SPT], t2[SPT];
    double ev[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        v[k] = (i < P.N) ? P.in[(size_t)w * P.N + i] : make_double2(0.0, 0.0);
        ev[k] = (i < P.N) ? P.eVbar[i] : 0.0;
    }
    poly<SPT, true>(t1, v, sw, c, order, P, ev);   // M^-T[w,w]
    poly<SPT, false>(t2, t1, sw, c, order, P, ev); // M^-1[w,w]
    const int wm = P.L - 1 - w;  // mirror frequency, conj (:464-466; for odd L the middle one overwrites itself)
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        if (i < P.N) {
            if (wm != w) P.out[(size_t)w * P.N + i] = t2[k];
            P.out[(size_t)wm * P.N + i] = make_double2(t2[k].x, -t2[k].y);
        }
    }
}

template <int SPT>
void launch_apply(elph_handle* h, const KpmParams& P, int threads) {
    const size_t smem = (size_t)h->N * sizeof(cplx);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Nsites too large for the KPM shared-memory kernel");
    elph_enable_smem(h, kpm_apply_kernel<SPT>);
    kpm_apply_kernel<SPT><<<h->kpm.nsched, threads, smem, h->stream>>>(P);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace

void elph_kpm_init(elph_handle* h, int n, double buf, double c1, double c2) {
    KpmState& K = h->kpm;
    K.configured = true;
    K.n = std::min(n, h->N);
    K.buf = buf;
    K.c1 = c1;
    K.c2 = c2;
    K.Lo2 = (h->L + 1) / 2;
    K.phis.resize(K.Lo2);
    const double pi = 3.14159265358979323846;
    for (int w = 0; w < K.Lo2; ++w) K.phis[w] = 2 * pi / h->L * (w + 0.5);
    K.order.assign(K.Lo2, 1);
    K.coeff_off.resize(K.Lo2);
    std::iota(K.coeff_off.begin(), K.coeff_off.end(), 0);
    K.coeff.assign(K.Lo2, zc(0.0, 0.0));
    K.schedule.resize(K.Lo2);
    std::iota(K.schedule.begin(), K.schedule.end(), 0);
    K.nsched = K.Lo2;
    K.eVbar.assign(h->N, 0.0);
    K.cbar.assign(h->Nb, 0.0);
    K.sbar.assign(h->Nb, 0.0);
    K.d_eVbar = elph_dalloc<double>(h->N);
    K.d_csbar = elph_dalloc<double2>(h->Nb);
    K.d_order = elph_dalloc<int>(K.Lo2);
    K.d_coeff_off = elph_dalloc<int>(K.Lo2);
    K.d_schedule = elph_dalloc<int>(K.Lo2);
    K.d_nu = elph_dalloc<cplx>((size_t)h->L * h->N);
    ELPH_CUDA(cudaMemset(K.d_nu, 0, (size_t)h->L * h->N * sizeof(cplx)));
    ELPH_CUDA(cudaDeviceSynchronize());
    h->kpm_version++;
}

void elph_kpm_free(elph_handle* h) {
    KpmState& K = h->kpm;
    cudaFree(K.d_eVbar);
    cudaFree(K.d_csbar);
    cudaFree(K.d_coeff);
    cudaFree(K.d_order);
    cudaFree(K.d_coeff_off);
    cudaFree(K.d_schedule);
    cudaFree(K.d_nu);
    cudaFree(K.d_noise);
    cudaFree(K.d_csbar_tile);
    cudaFree(K.d_hm);
    cudaFree(K.d_Q);
    if (K.spec_stream) {
        cudaStreamDestroy(K.spec_stream);
        cudaEventDestroy(K.spec_fork);
        cudaEventDestroy(K.spec_done);
    }
    if (K.h_hm) cudaFreeHost(K.h_hm);
    K = KpmState();
}

// frequencies the chain kernels run, longest polynomial first; with an omega subset (omega-sharded apply of the tau-sharded
// lattice) only w = first, first + stride, ... of them
static void kpm_build_schedule(elph_handle* h) {
    KpmState& K = h->kpm;
    K.schedule.clear();
    for (int w = K.sub_first; w < K.Lo2; w += K.sub_stride) K.schedule.push_back(w);
    std::stable_sort(K.schedule.begin(), K.schedule.end(), [&](int a, int b) { return K.order[a] > K.order[b]; });
    K.nsched = (int)K.schedule.size();
    if (K.nsched)
        ELPH_CUDA(cudaMemcpyAsync(K.d_schedule, K.schedule.data(), K.nsched * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));   // the host vector may be rebuilt by the next call
    h->kpm_version++;
}

// Speculative set-up (dynamics.cu): the hysteresis keeps the polynomials of the previous set-up unless the spectral window moved
// by more than `buf` (src/KPMPreconditioners.jl:296-309), so the solve that follows setup!(P) can start as soon as update_A! is
// queued, while the Arnoldi kernel runs beside it on two SMs and the host reduces the two Hessenberg matrices.
bool elph_kpm_can_speculate(const elph_handle* h) {
    const KpmState& K = h->kpm;
    return h->kpm_speculate && K.configured && K.ever_setup && K.active && K.d_coeff && h->kpm_dev_arnoldi && K.n <= 32 && h->N <= 8192 &&
           !h->trace;
}
void elph_kpm_setup_begin(elph_handle* h, const double* noise) { elph_kpm_setup_impl(h, noise, nullptr, nullptr, 1); }
// true when the speculative solve has to be repeated (polynomials recomputed or the preconditioner switched itself off)
bool elph_kpm_setup_finish(elph_handle* h, elph_kpm_info* info) {
    elph_kpm_setup_impl(h, nullptr, info, nullptr, 2);
    return h->kpm.spec_stale;
}

void elph_kpm_set_omega_subset(elph_handle* h, int first, int stride) {
    KpmState& K = h->kpm;
    ELPH_REQUIRE(K.configured, ELPH_ERR_STATE, "KPM preconditioner not configured (kpm_n == 0 at elph_create)");
    ELPH_REQUIRE(stride >= 1 && first >= 0 && first < stride, ELPH_ERR_INVALID, "omega subset: need 0 <= first < stride");
    K.sub_first = first;
    K.sub_stride = stride;
    kpm_build_schedule(h);
}

// phase 0: the whole of setup!(P).  Phases 1 / 2 split it for the speculative solve of the force evaluation (dynamics.cu): phase 1
// queues update_A! on the handle's stream and the Arnoldi kernel + read-back on the side stream and returns without waiting;
// phase 2 waits for the read-back, finishes (bounds, hysteresis, coefficients) and tells through info->recomputed / the return
// of elph_kpm_setup_finish whether the polynomials the speculative solve used are still the right ones.
void elph_kpm_setup_impl(elph_handle* h, const double* noise, elph_kpm_info* info, const double* ext_eVbar, int phase) {
    KpmState& K = h->kpm;
    ELPH_REQUIRE(K.configured, ELPH_ERR_STATE, "KPM preconditioner not configured (kpm_n == 0 at elph_create)");
    ELPH_REQUIRE(noise != nullptr || phase == 2, ELPH_ERR_INVALID, "arnoldi_noise must provide 2*Nsites values");
    const int N = h->N, L = h->L, Nb = h->Nb, T = 256;
    // update_A!
    const bool dev = h->kpm_dev_arnoldi && K.n <= 32 && (N <= 8192);
    ELPH_REQUIRE(phase == 0 || (dev && K.ever_setup), ELPH_ERR_STATE, "split KPM set-up needs the device Arnoldi path and a previous set-up");
    cudaStream_t arn_stream = h->stream;
    if (phase == 1) {
        if (!K.spec_stream) {
            ELPH_CUDA(cudaStreamCreateWithFlags(&K.spec_stream, cudaStreamNonBlocking));
            ELPH_CUDA(cudaEventCreateWithFlags(&K.spec_fork, cudaEventDisableTiming));
            ELPH_CUDA(cudaEventCreateWithFlags(&K.spec_done, cudaEventDisableTiming));
        }
        arn_stream = K.spec_stream;
    }
    if (phase != 2) {
    if (h->model == ELPH_MODEL_HOLSTEIN) {
        if (ext_eVbar) {   // tau-sharded lattice: the mean over ALL slices was summed across the ranks by the caller
            ELPH_CUDA(cudaMemcpyAsync(K.d_eVbar, ext_eVbar, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        } else {
            taumean_kernel<<<(N + T - 1) / T, T, 0, h->stream>>>(h->d_D, K.d_eVbar, N, L);
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
        }
        if (!K.ever_setup) {  // static hoppings: cbar = cosht, sbar = sinht (:128-130)
            std::vector<double2> cs(Nb);
            ELPH_CUDA(cudaMemcpyAsync(cs.data(), h->d_cs, Nb * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            for (int b = 0; b < Nb; ++b) { K.cbar[b] = cs[b].x; K.sbar[b] = cs[b].y; }
            ELPH_CUDA(cudaMemcpyAsync(K.d_csbar, h->d_cs, Nb * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
        }
        if (!dev) {
            ELPH_CUDA(cudaMemcpyAsync(K.eVbar.data(), K.d_eVbar, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
        }
    } else {
        if (Nb > 0) {
            if (h->ssq.enabled && !K.d_csbar_tile) K.d_csbar_tile = elph_dalloc<double2>(Nb);
            if (ext_eVbar) {   // tau-sharded SSH lattice: ext = the all-reduced mean of the (cosh, sinh) pairs, [Ncolumns][2]
                ELPH_CUDA(cudaMemcpyAsync(K.d_csbar, ext_eVbar, Nb * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
                if (h->ssq.enabled)
                    csbar_tile_kernel<<<(Nb + T - 1) / T, T, 0, h->stream>>>(K.d_csbar, h->ssq.d_slot, K.d_csbar_tile, Nb);
            } else {
                taumean2_kernel<<<(Nb + T - 1) / T, T, 0, h->stream>>>(h->d_cs, K.d_csbar, Nb, L, h->ssq.enabled ? h->ssq.d_slot : nullptr,
                                                                       h->ssq.enabled ? K.d_csbar_tile : nullptr);
            }
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
        }
        ELPH_CUDA(cudaMemcpyAsync(K.d_eVbar, h->d_D, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        if (!dev) {
            std::vector<double2> cs(Nb);
            ELPH_CUDA(cudaMemcpyAsync(cs.data(), K.d_csbar, Nb * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaMemcpyAsync(K.eVbar.data(), K.d_eVbar, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            for (int b = 0; b < Nb; ++b) { K.cbar[b] = cs[b].x; K.sbar[b] = cs[b].y; }
        }
    }
    elph_trace_mark(h, "  kpm: tau-mean");
    }   // phase != 2
    // Arnoldi bounds
    double e_max, inv_max;
    if (dev) {
        // both Krylov runs in one launch (arnoldi_kernel): only the two small Hessenberg matrices come back
        const int n = K.n;
        const size_t hstride = (size_t)(n + 1) * n + 1;
        if (phase != 2) {
        if (phase == 1) {   // the side stream starts where update_A! ends on the handle's stream
            ELPH_CUDA(cudaEventRecord(K.spec_fork, h->stream));
            ELPH_CUDA(cudaStreamWaitEvent(arn_stream, K.spec_fork, 0));
        }
        if (!K.d_noise) {
            K.d_noise = elph_dalloc<double>(2 * (size_t)N);
            K.d_hm = elph_dalloc<double>(4 * hstride);
            ELPH_CUDA(cudaMallocHost(&K.h_hm, 4 * hstride * sizeof(double)));
        }
        ELPH_CUDA(cudaMemcpyAsync(K.d_noise, noise, 2 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, arn_stream));
        ArnoldiParams A;
        A.eVbar = K.d_eVbar; A.csbar = K.d_csbar; A.bonds = h->d_bonds; A.goff = h->d_goff; A.start = K.d_noise;
        A.hm = K.d_hm; A.ngroups = h->ngroups; A.N = N; A.n = n;
        A.prof = h->pipe_prof ? h->pipe_prof_buf : nullptr;
        // register tiles on periodic square lattices with one (cosh, sinh) per colour: 32 sites wide, 8 rows per warp
        const bool square = h->model == ELPH_MODEL_HOLSTEIN && h->sq.enabled && !h->sq_disable && h->sq.Lx == 32 &&
                            h->sq.Ly % 8 == 0 && h->sq.Ly / 8 <= 8;
        A.Ly = square ? h->sq.Ly : 0;
        for (int g = 0; g < 4; ++g) { A.c[g] = square ? h->sq.c[g] : 1.0; A.s[g] = square ? h->sq.s[g] : 0.0; }
        const int ept = square ? 8 : ((N <= 2048) ? 8 : ((N <= 4096) ? 16 : 32));
        const int threads = square ? 32 * (h->sq.Ly / 8) : ((N + ept - 1) / ept + 31) / 32 * 32;
        const size_t fixed = (2ull * kArnWarps * kArnStride + (square ? 2ull * (threads / 32) * 2 * 32 : (size_t)N)) * sizeof(double);
        size_t smem = fixed + (size_t)(n + 4) * N * sizeof(double);
        A.q_in_smem = smem <= h->smem_optin ? 1 : 0;
        if (!A.q_in_smem) {
            smem = fixed;
            ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Nsites too large for the Arnoldi kernel");
            if (!K.d_Q) K.d_Q = elph_dalloc<double>(2 * (size_t)(n + 4) * N);
        }
        A.Qg = K.d_Q;
#define ARN_LAUNCH(E, SQ)                                                                      \
        do {                                                                                   \
            if (A.q_in_smem) {                                                                 \
                elph_enable_smem(h, arnoldi_kernel<E, SQ, true>);                              \
                arnoldi_kernel<E, SQ, true><<<2, threads, smem, arn_stream>>>(A);              \
            } else {                                                                           \
                elph_enable_smem(h, arnoldi_kernel<E, SQ, false>);                             \
                arnoldi_kernel<E, SQ, false><<<2, threads, smem, arn_stream>>>(A);             \
            }                                                                                  \
        } while (0)
        if (square) ARN_LAUNCH(8, true);
        else if (ept == 8) ARN_LAUNCH(8, false);
        else if (ept == 16) ARN_LAUNCH(16, false);
        else ARN_LAUNCH(32, false);
#undef ARN_LAUNCH
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        ELPH_CUDA(cudaMemcpyAsync(K.h_hm, K.d_hm, 4 * hstride * sizeof(double), cudaMemcpyDeviceToHost, arn_stream));
        }   // phase != 2
        // h = first-pass + second-pass coefficients (the sub-diagonal norms sit in the first matrix only), then the two bounds
        auto reduce_bounds = [hm = K.h_hm, hstride, n]() {
            for (int run = 0; run < 2; ++run)
                for (size_t i = 0; i + 1 < hstride; ++i) hm[2 * run * hstride + i] += hm[(2 * run + 1) * hstride + i];
            const int l0 = (int)hm[hstride - 1], l1 = (int)hm[3 * hstride - 1];
            return std::make_pair(hessenberg_bound(hm, n, l0), hessenberg_bound(hm + 2 * hstride, n, l1));
        };
        if (phase == 1) {
            // a host thread waits for the read-back and reduces the Hessenberg matrices while the caller queues and waits for the solve
            ELPH_CUDA(cudaEventRecord(K.spec_done, arn_stream));
            K.spec_future = std::async(std::launch::async, [done = K.spec_done, dev_id = h->device, reduce_bounds]() {
                cudaSetDevice(dev_id);
                const cudaError_t e = cudaEventSynchronize(done);
                if (e != cudaSuccess) return std::make_pair((double)NAN, (double)NAN);
                return reduce_bounds();
            });
            return;
        }
        std::pair<double, double> eb;
        if (phase == 2) {
            eb = K.spec_future.get();
            ELPH_REQUIRE(eb.first == eb.first, ELPH_ERR_CUDA, "speculative KPM set-up: the Arnoldi read-back failed");
        } else {
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            elph_trace_mark(h, "  kpm: arnoldi kernel");
            eb = reduce_bounds();
        }
        e_max = eb.first;
        inv_max = eb.second;
    } else {
        // the two Krylov runs (on A for e_max, on A^-1 for e_min) are independent: second host thread for the inverse one
        auto inv_run = std::async(std::launch::async, [&]() { return host_arnoldi(h, K, noise + N, true); });
        e_max = host_arnoldi(h, K, noise, false);
        inv_max = inv_run.get();
    }
    elph_trace_mark(h, "  kpm: eigenvalues");
    const double e_min = std::isfinite(inv_max) ? 1.0 / inv_max : -INFINITY;
    const bool was_active = K.active;
    K.e_min = e_min;
    K.e_max = e_max;
    bool recomputed = false;
    if ((0.0 < e_min && e_min < 1.0) && (1.0 < e_max) && (e_max - e_min) < 2.0) {
        const double lam_lo = std::max(0.0, (1 - 2 * K.buf) * e_min);
        const double lam_hi = (1 + 2 * K.buf) * e_max;
        if (!isapprox_rtol(lam_lo, K.lam_lo, K.buf) || !isapprox_rtol(lam_hi, K.lam_hi, K.buf)) {
            K.lam_lo = lam_lo;
            K.lam_hi = lam_hi;
            K.lam_avg = (lam_hi + lam_lo) / 2;
            K.lam_mag = (lam_hi - lam_lo) / 2;
            int total = 0;
            for (int w = 0; w < K.Lo2; ++w) {
                int order = (int)std::floor((lam_hi - lam_lo) * (K.c1 / K.phis[w] + K.c2));
                order = std::max(1, order);
                K.order[w] = order;
                K.coeff_off[w] = total;
                total += order;
            }
            K.coeff.assign(total, zc(0.0, 0.0));
            for (int w = 0; w < K.Lo2; ++w) host_kpm_coefficients(&K.coeff[K.coeff_off[w]], K.order[w], lam_lo, lam_hi, K.phis[w]);
            if ((size_t)total > K.d_coeff_cap) {
                cudaFree(K.d_coeff);
                K.d_coeff_cap = (size_t)total * 2;
                K.d_coeff = elph_dalloc<cplx>(K.d_coeff_cap);
            }
            static_assert(sizeof(zc) == sizeof(cplx), "complex layout");
            ELPH_CUDA(cudaMemcpyAsync(K.d_coeff, K.coeff.data(), total * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
            ELPH_CUDA(cudaMemcpyAsync(K.d_order, K.order.data(), K.Lo2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            ELPH_CUDA(cudaMemcpyAsync(K.d_coeff_off, K.coeff_off.data(), K.Lo2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            kpm_build_schedule(h);   // synchronises: host vectors may be reallocated by the next setup; bumps kpm_version
            recomputed = true;     // (captured CG graphs carry the old polynomial orders / window)
        }
        K.active = true;
    } else {
        K.active = false;
    }
    K.ever_setup = true;
    K.spec_stale = recomputed || (K.active != was_active);
    if (info) {
        info->active = K.active ? 1 : 0;
        info->recomputed = recomputed ? 1 : 0;
        info->e_min = e_min;
        info->e_max = e_max;
        info->lambda_lo = K.lam_lo;
        info->lambda_hi = K.lam_hi;
        int64_t tot = 0, mx = 0;
        for (int w = 0; w < K.Lo2; ++w) { tot += K.order[w]; mx = std::max<int64_t>(mx, K.order[w]); }
        info->total_order = tot;
        info->max_order = mx;
    }
}

// the Chebyshev chains of every scheduled frequency: nu_out[w], nu_out[L-1-w] from nu_in[w]  (:606-679, :464-466)
static void kpm_launch_chains(elph_handle* h, const cplx* nu_in, cplx* nu_out, const int* skip) {
    KpmState& K = h->kpm;
    if (K.nsched == 0) return;
    KpmParams P;
    P.in = nu_in;
    P.out = nu_out;
    P.eVbar = K.d_eVbar;
    P.csbar = K.d_csbar;
    P.bonds = h->d_bonds;
    P.goff = h->d_goff;
    P.coeff = K.d_coeff;
    P.order = K.d_order;
    P.coeff_off = K.d_coeff_off;
    P.schedule = K.d_schedule;
    P.skip = skip;
    P.ngroups = h->ngroups;
    P.N = h->N;
    P.L = h->L;
    P.inv_mag = 1.0 / K.lam_mag;
    P.avg_over_mag = K.lam_avg / K.lam_mag;
    if (elph_launch_kpm_square(h, nu_in, nu_out, skip)) return;   // register/shuffle kernel (kpm_square.cu)
    int threads = 256;
    while (threads < 512 && threads * 4 < h->N) threads *= 2;   // <= 512 threads: __launch_bounds__(512) on the kernel
    const int spt = (h->N + threads - 1) / threads;
    if (spt <= 1) launch_apply<1>(h, P, threads);
    else if (spt <= 2) launch_apply<2>(h, P, threads);
    else if (spt <= 4) launch_apply<4>(h, P, threads);
    else if (spt <= 8) launch_apply<8>(h, P, threads);
    else if (spt <= 16) launch_apply<16>(h, P, threads);
    else ELPH_REQUIRE(false, ELPH_ERR_UNSUPPORTED, "Nsites too large for the KPM kernel");
}

// omega-sharded apply: only the chain phase, on frequency-space vectors [L][N] indexed by the GLOBAL frequency (the caller
// moved its frequencies' rows in through the all-to-all; rows of other frequencies are not touched)
void elph_kpm_chains_dev(elph_handle* h, const cplx* nu_in, cplx* nu_out) {
    KpmState& K = h->kpm;
    ELPH_REQUIRE(K.configured && K.ever_setup && K.active && K.d_coeff, ELPH_ERR_STATE, "elph_dev_kpm_chains needs an active preconditioner");
    ELPH_REQUIRE(nu_in && nu_out && nu_in != nu_out, ELPH_ERR_INVALID, "bad frequency-space buffers");
    kpm_launch_chains(h, nu_in, nu_out, nullptr);
}

void elph_kpm_apply_dev(elph_handle* h, const double* vin, double* vout) { elph_kpm_apply_dev_cg(h, vin, vout, nullptr); }

// cgf != nullptr (inside the preconditioned CG loop, preconditioner active): the forward FFT kernel first applies
// x += alpha p, r -= alpha Ap and the stop rule, and the inverse FFT kernel accumulates r.z -> beta (fft.cu);
// vin must then be the residual vector r.
void elph_kpm_apply_dev_cg(elph_handle* h, const double* vin, double* vout, const KpmCgFuse* cgf) {
    KpmState& K = h->kpm;
    ELPH_REQUIRE(K.configured && K.ever_setup, ELPH_ERR_STATE, "elph_kpm_apply before elph_kpm_setup");
    ELPH_REQUIRE(!cgf || K.active, ELPH_ERR_STATE, "fused KPM apply needs an active preconditioner");
    if (!K.active) {  // identity (:475-478)
        if (vout != vin) ELPH_CUDA(cudaMemcpyAsync(vout, vin, h->Ndim * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return;
    }
    ELPH_REQUIRE(K.d_coeff != nullptr, ELPH_ERR_STATE, "KPM coefficients missing");
    ELPH_REQUIRE(K.sub_stride == 1, ELPH_ERR_STATE, "this handle runs an omega subset (elph_kpm_set_omega_subset): use elph_dev_kpm_chains");
    const int* skip = &h->d_cg->done;
    if (!h->kpm_skip_enabled) skip = nullptr;
    // the recurrence reads nu_in and writes nu_out (the reference's v1 / v2)
    cplx* nu_in = K.d_nu;
    cplx* nu_out = h->d_nu2;
    auto inverse_fft = [&]() {
        if (cgf) elph_omega_to_tau_dev_cg(h, nu_out, vout, cgf->r);
        else elph_omega_to_tau_dev_skip(h, nu_out, vout, skip);
    };
    if (cgf) elph_tau_to_omega_dev_cg(h, cgf->x, cgf->r, cgf->p, cgf->ap, nu_in);
    else elph_tau_to_omega_dev_skip(h, vin, nu_in, skip);
    kpm_launch_chains(h, nu_in, nu_out, skip);
    inverse_fft();
}
