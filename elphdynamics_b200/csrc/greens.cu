// Green's-function convolutions of the stochastic estimator.
//
// Reference: src/GreensFunctions.jl  setup!(estimator, n1, n2) :239-296, convolve! :361-414, antiperiodic_copy! :420-433,
// periodic_product! :439-457.  For every pair (n1, n2) of random vectors the reference builds four pairs of complex arrays
// (2L, n_s, L1, L2, L3) from r1, r2, M^-1 r1, M^-1 r2, transforms both over (omega, k1, k2, k3) with FFTW, multiplies
// a'[w, s2, k] b'[-w, s1, -k] / V into (2L, n_s, n_s, L1, L2, L3) and transforms back: 12 multi-dimensional FFTs per
// pair, n_v (n_v - 1) / 2 pairs per measurement -- on the CPU that is seconds per measurement once the solves run on the
// GPU, so it moves too.
//
// Everything here works in the HOST layout (tau fastest), which is the reference's array layout.  The transforms are
// separable: the imaginary-time axis (length 2L, contiguous) runs one line per CTA in shared memory as a two-factor
// Cooley-Tukey step n = n1 n2 with direct sub-transforms (any length; n1 ~ sqrt(n), n2 = 1 for primes); the lattice axes
// (short, strided) are direct transforms on shared-memory tiles that are contiguous along omega, so global accesses stay
// coalesced.  Twiddles come from host-computed tables (exact to rounding for any length).  Bandwidth is irrelevant at
// these sizes (6.5 MB arrays at config B); what matters is that nothing returns to the host between the 44 launches of
// a pair.
#include "elph_internal.cuh"

#include <cmath>
#include <map>
#include <vector>

namespace {

constexpr int kT = 256;
constexpr int kTile = 32;   // omega values per tile of the lattice-axis transforms

struct GreensState {
    int nv = 0;
    int64_t ndim = 0;
    double* d_R = nullptr;      // [nv][Ndim] host layout
    double* d_MinvR = nullptr;
    cplx* a = nullptr;          // (2L, ns, cells)
    cplx* b = nullptr;
    cplx* ab = nullptr;         // (2L, ns, ns, cells)
    size_t cap_a = 0, cap_b = 0, cap_ab = 0;
    std::map<int, cplx*> tw;    // twiddle tables exp(-2 pi i k / n), k = 0 .. n-1
    double* d_stage = nullptr;  // 4 outputs (complex) for one D2H copy
    size_t cap_stage = 0;
};

GreensState* state(elph_handle* h) {
    if (!h->greens) h->greens = new GreensState();
    return static_cast<GreensState*>(h->greens);
}

__device__ __forceinline__ cplx cmulz(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cfmaz(cplx a, cplx b, cplx c) {
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// a, b of the four convolutions of setup! (:262-293) in one pass; line = (orbital, cell), 2L entries per line.
//   0: a = (ap(m1) + ap(m2)) / sqrt 2, b = (ap(r1) + ap(r2)) / sqrt 2      ap = antiperiodic copy
//   1: a = pp(m1, m2),  b = pp(r1, r2)                                      pp = periodic product
//   2: a = pp(m2, r2),  b = pp(m1, r1)
//   3: a = pp(m1, r2),  b = pp(m2, r1)
__global__ void build_ab_kernel(int conv, const double* __restrict__ r1, const double* __restrict__ m1, const double* __restrict__ r2,
                                const double* __restrict__ m2, cplx* __restrict__ a, cplx* __restrict__ b, int L, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long line = i / L;
        const int tau = (int)(i - line * L);
        const double R1 = r1[i], M1 = m1[i], R2 = r2[i], M2 = m2[i];
        double av, bv, sgn = 1.0;
        if (conv == 0) {
            const double s2 = sqrt(2.0);
            av = (M1 + M2) / s2;
            bv = (R1 + R2) / s2;
            sgn = -1.0;
        } else if (conv == 1) {
            av = M1 * M2;
            bv = R1 * R2;
        } else if (conv == 2) {
            av = M2 * R2;
            bv = M1 * R1;
        } else {
            av = M1 * R2;
            bv = M2 * R1;
        }
        const long long o = line * 2 * L + tau;
        a[o] = make_double2(av, 0.0);
        a[o + L] = make_double2(sgn * av, 0.0);
        b[o] = make_double2(bv, 0.0);
        b[o + L] = make_double2(sgn * bv, 0.0);
    }
}

// Transform of length n = n1 n2 along the contiguous axis, one line per CTA, in place.  tw[k] = exp(-2 pi i k / n);
// inverse: conjugated twiddles.  out[k1 + n1 k2] = sum_j2 W^(j2 k2 n1) [ W^(j2 k1) sum_j1 x[j1 n2 + j2] W^(j1 k1 n2) ].
__global__ void __launch_bounds__(kT) dft_contig_kernel(cplx* __restrict__ data, int n, int n1, int n2, const cplx* __restrict__ tw,
                                                        int inverse, double scale) {
    extern __shared__ __align__(16) unsigned char gsm[];
    cplx* x = reinterpret_cast<cplx*>(gsm);
    cplx* y = x + n;
    cplx* w = y + n;
    cplx* line = data + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        x[i] = line[i];
        cplx t = tw[i];
        if (inverse) t.y = -t.y;
        w[i] = t;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        const int k1 = o / n2, j2 = o - k1 * n2;
        cplx acc = make_double2(0.0, 0.0);
        int e = 0;                                   // (j1 k1 n2) mod n
        const int step = (int)(((long long)k1 * n2) % n);
        for (int j1 = 0; j1 < n1; ++j1) {
            acc = cfmaz(x[j1 * n2 + j2], w[e], acc);
            e += step;
            if (e >= n) e -= n;
        }
        y[o] = cmulz(acc, w[(int)(((long long)j2 * k1) % n)]);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        const int k2 = o / n1, k1 = o - k2 * n1;    // output index k = k1 + n1 k2 = o
        cplx acc = make_double2(0.0, 0.0);
        int e = 0;                                   // (j2 k2 n1) mod n
        const int step = (int)(((long long)k2 * n1) % n);
        for (int j2 = 0; j2 < n2; ++j2) {
            acc = cfmaz(y[k1 * n2 + j2], w[e], acc);
            e += step;
            if (e >= n) e -= n;
        }
        line[o] = make_double2(acc.x * scale, acc.y * scale);
    }
}

// Direct transform of length n along an axis of stride `stride` (elements), in place.  The array is viewed as
// [outer][n][stride]; a CTA owns one outer index and kTile consecutive inner indices: tile [n][kTile] in shared memory.
__global__ void __launch_bounds__(kT) dft_axis_kernel(cplx* __restrict__ data, int n, long long stride, const cplx* __restrict__ tw,
                                                      int inverse) {
    extern __shared__ __align__(16) unsigned char gsm[];
    cplx* tile = reinterpret_cast<cplx*>(gsm);      // [n][kTile]
    cplx* w = tile + (size_t)n * kTile;
    const long long tiles_per_outer = (stride + kTile - 1) / kTile;
    const long long outer = blockIdx.x / tiles_per_outer;
    const long long in0 = (blockIdx.x - outer * tiles_per_outer) * kTile;
    cplx* base = data + outer * (long long)n * stride + in0;
    const int tx = threadIdx.x % kTile, ty = threadIdx.x / kTile, ny = blockDim.x / kTile;
    const bool live = in0 + tx < stride;
    for (int j = ty; j < n; j += ny)
        if (live) tile[j * kTile + tx] = base[(long long)j * stride + tx];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        cplx t = tw[i];
        if (inverse) t.y = -t.y;
        w[i] = t;
    }
    __syncthreads();
    cplx outv[8];                                    // n <= 8 * ny outputs per thread (checked by the host)
    int cnt = 0;
    for (int k = ty; k < n; k += ny, ++cnt) {
        cplx acc = make_double2(0.0, 0.0);
        int e = 0;
        for (int j = 0; j < n; ++j) {
            acc = cfmaz(tile[j * kTile + tx], w[e], acc);
            e += k;
            if (e >= n) e -= n;
        }
        outv[cnt] = acc;
    }
    __syncthreads();
    cnt = 0;
    for (int k = ty; k < n; k += ny, ++cnt)
        if (live) base[(long long)k * stride + tx] = outv[cnt];
}

// ab'[w, s2, s1, k] = a'[w, s2, k] b'[-w, s1, -k] / V     (:381-399)
__global__ void product_kernel(const cplx* __restrict__ a, const cplx* __restrict__ b, cplx* __restrict__ ab, int n2L, int ns, int L1,
                               int L2, int L3, double invV, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int w = (int)(t % n2L); t /= n2L;
        const int s2 = (int)(t % ns); t /= ns;
        const int s1 = (int)(t % ns); t /= ns;
        const int k1 = (int)(t % L1); t /= L1;
        const int k2 = (int)(t % L2); t /= L2;
        const int k3 = (int)t;
        const int nw = w ? n2L - w : 0, nk1 = k1 ? L1 - k1 : 0, nk2 = k2 ? L2 - k2 : 0, nk3 = k3 ? L3 - k3 : 0;
        const long long cell = k1 + (long long)L1 * (k2 + (long long)L2 * k3);
        const long long ncell = nk1 + (long long)L1 * (nk2 + (long long)L2 * nk3);
        const cplx av = a[w + (long long)n2L * (s2 + ns * cell)];
        const cplx bv = b[nw + (long long)n2L * (s1 + ns * ncell)];
        const cplx p = cmulz(av, bv);
        ab[i] = make_double2(p.x * invV, p.y * invV);
    }
}

const cplx* twiddles(elph_handle* h, GreensState* G, int n) {
    auto it = G->tw.find(n);
    if (it != G->tw.end()) return it->second;
    std::vector<cplx> t(n);
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int k = 0; k < n; ++k) {
        const long double ang = two_pi * (long double)k / (long double)n;
        t[k] = make_double2((double)cosl(ang), (double)-sinl(ang));
    }
    cplx* d = elph_dalloc<cplx>(n);
    ELPH_CUDA(cudaMemcpy(d, t.data(), n * sizeof(cplx), cudaMemcpyHostToDevice));
    G->tw[n] = d;
    return d;
}

// n = n1 n2 with n1 the largest divisor <= sqrt(n)  (n1 = 1 for primes: one direct transform)
void split(int n, int& n1, int& n2) {
    n1 = 1;
    for (int d = 1; (long long)d * d <= n; ++d)
        if (n % d == 0) n1 = d;
    n2 = n / n1;
}

// transform of `data` viewed as (n2L, inner, L1, L2, L3) over omega and the three lattice axes
void fft4(elph_handle* h, GreensState* G, cplx* data, int n2L, long long inner, int L1, int L2, int L3, bool inverse) {
    const long long lines = inner * L1 * L2 * L3;
    int n1, n2;
    split(n2L, n1, n2);
    const double scale = inverse ? 1.0 / ((double)n2L * L1 * L2 * L3) : 1.0;
    const size_t smem = 3ull * n2L * sizeof(cplx);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Green's-function convolution: imaginary-time extent too long");
    elph_enable_smem(h, dft_contig_kernel);
    dft_contig_kernel<<<(unsigned)lines, kT, smem, h->stream>>>(data, n2L, n1, n2, twiddles(h, G, n2L), inverse ? 1 : 0, scale);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    const int dims[3] = {L1, L2, L3};
    long long stride = (long long)n2L * inner;
    for (int ax = 0; ax < 3; ++ax) {
        const int n = dims[ax];
        if (n > 1) {
            ELPH_REQUIRE(n <= 8 * (kT / kTile), ELPH_ERR_UNSUPPORTED, "Green's-function convolution: lattice extent above 64");
            const long long outer = (lines * n2L) / (stride * n);
            const long long tiles = (stride + kTile - 1) / kTile;
            const size_t sm = ((size_t)n * kTile + n) * sizeof(cplx);
            elph_enable_smem(h, dft_axis_kernel);
            dft_axis_kernel<<<(unsigned)(outer * tiles), kT, sm, h->stream>>>(data, n, stride, twiddles(h, G, n), inverse ? 1 : 0);
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
        }
        stride *= n;
    }
}

template <typename T>
void grow(T*& p, size_t& cap, size_t n) {
    if (cap >= n) return;
    if (p) cudaFree(p);
    p = elph_dalloc<T>(n);
    cap = n;
}

}  // namespace

void elph_greens_free(elph_handle* h) {
    if (!h->greens) return;
    GreensState* G = static_cast<GreensState*>(h->greens);
    cudaFree(G->d_R); cudaFree(G->d_MinvR); cudaFree(G->a); cudaFree(G->b); cudaFree(G->ab); cudaFree(G->d_stage);
    for (auto& kv : G->tw) cudaFree(kv.second);
    delete G;
    h->greens = nullptr;
}

// R, M^-1 R of update!(Gr, model, P) (:201-234), host layout, nv vectors of length Ndim each, kept on the device
void elph_greens_load_impl(elph_handle* h, int nv, const double* R, const double* MinvR) {
    GreensState* G = state(h);
    const size_t n = (size_t)nv * h->Ndim;
    if (G->nv != nv || G->ndim != h->Ndim) {
        cudaFree(G->d_R); cudaFree(G->d_MinvR);
        G->d_R = elph_dalloc<double>(n);
        G->d_MinvR = elph_dalloc<double>(n);
        G->nv = nv;
        G->ndim = h->Ndim;
    }
    ELPH_CUDA(cudaMemcpyAsync(G->d_R, R, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ELPH_CUDA(cudaMemcpyAsync(G->d_MinvR, MinvR, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
}

// setup!(estimator, n1, n2): the four convolutions for one pair; out[c] = complex array (2L, ns, ns, L1, L2, L3) in the
// reference's memory order, c = 0 G[D,0], 1 G[D,0] G[D,0], 2 G[D,D] G[0,0], 3 G[D,0] G[0,D]
void elph_greens_setup_impl(elph_handle* h, int n1, int n2, int L1, int L2, int L3, int ns, double* const out[4]) {
    GreensState* G = state(h);
    ELPH_REQUIRE(G->nv > 0, ELPH_ERR_STATE, "elph_greens_load has not been called");
    ELPH_REQUIRE(n1 >= 0 && n1 < G->nv && n2 >= 0 && n2 < G->nv, ELPH_ERR_INVALID, "vector index out of range");
    ELPH_REQUIRE(L1 >= 1 && L2 >= 1 && L3 >= 1 && ns >= 1 && (int64_t)ns * L1 * L2 * L3 == h->N, ELPH_ERR_INVALID,
                 "norbits * L1 * L2 * L3 must equal Nsites");
    const int L = h->L, n2L = 2 * L;
    const long long cells = (long long)L1 * L2 * L3;
    const size_t na = (size_t)n2L * ns * cells, nab = na * ns;
    grow(G->a, G->cap_a, na);
    grow(G->b, G->cap_b, na);
    grow(G->ab, G->cap_ab, nab);
    grow(G->d_stage, G->cap_stage, 4 * 2 * nab);
    const double* r1 = G->d_R + (size_t)n1 * h->Ndim;
    const double* m1 = G->d_MinvR + (size_t)n1 * h->Ndim;
    const double* r2 = G->d_R + (size_t)n2 * h->Ndim;
    const double* m2 = G->d_MinvR + (size_t)n2 * h->Ndim;
    const double invV = 1.0 / (2.0 * L * (double)h->N / ns);
    const int blocks = (int)std::min<long long>((h->Ndim + kT - 1) / kT, 8LL * h->sm_count);
    const int pblocks = (int)std::min<long long>(((long long)nab + kT - 1) / kT, 8LL * h->sm_count);
    for (int c = 0; c < 4; ++c) {
        build_ab_kernel<<<blocks, kT, 0, h->stream>>>(c, r1, m1, r2, m2, G->a, G->b, L, h->Ndim);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        fft4(h, G, G->a, n2L, ns, L1, L2, L3, false);
        fft4(h, G, G->b, n2L, ns, L1, L2, L3, false);
        product_kernel<<<pblocks, kT, 0, h->stream>>>(G->a, G->b, G->ab, n2L, ns, L1, L2, L3, invV, (long long)nab);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        fft4(h, G, G->ab, n2L, (long long)ns * ns, L1, L2, L3, true);
        ELPH_CUDA(cudaMemcpyAsync(G->d_stage + (size_t)c * 2 * nab, G->ab, nab * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    }
    for (int c = 0; c < 4; ++c)
        if (out[c])
            ELPH_CUDA(cudaMemcpyAsync(out[c], G->d_stage + (size_t)c * 2 * nab, nab * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
}
