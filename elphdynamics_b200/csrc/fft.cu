// K4: batched shared-memory FFT along imaginary time, fused with the reference's
// pre/post operations.
//
// Reference: src/TimeFreqFFTs.jl:32-45 (Theta twist + plan), :55-73 tau_to_omega!,
// :112-130 omega_to_tau! (real part); src/FourierAcceleration.jl:91-143
// fourier_accelerate! (plain FFT, multiply by Q^p or M^p, inverse FFT, real part).
// FFTW conventions (FFTW.jl 1.3.2): forward sum_j x_j exp(-2 pi i j k/L) unnormalised,
// inverse scaled by 1/L.
//
// Engine layout is [tau][site]; one CTA owns SB consecutive sites for ALL tau, so
// global accesses are coalesced along the site index and the transform runs entirely
// in shared memory ([tau][SB] complex, site fastest => conflict-free butterflies,
// twiddles warp-uniform).  Lengths are arbitrary: a mixed-radix Stockham (autosort,
// decimation in frequency) with one generic radix-r stage per prime factor
// (4 preferred over 2x2).  Ltau = 20, 40, 200, 400 factor into {4, 2, 5}.
#include "elph_internal.cuh"

#include <cmath>

namespace {

constexpr int kT = 256;
constexpr int kMaxRad = 24;

struct FftPlan {
    int L;
    int nrad;
    int rad[kMaxRad];
};

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// In-place-on-two-buffers Stockham FFT of SB interleaved sequences of length L.
// x, y: shared buffers [L][SB].  tw: exp(-2 pi i k/L) table in global memory.
// Returns the buffer holding the result.  All threads of the CTA must call it.
template <int SB>
__device__ cplx* fft_smem(cplx* x, cplx* y, const FftPlan& plan, const cplx* __restrict__ tw, bool inverse) {
    const int L = plan.L;
    const int site = threadIdx.x % SB;
    const int slot = threadIdx.x / SB;
    const int nslots = blockDim.x / SB;
    int n = L;  // current sub-transform length
    int s = 1;  // current stride
    for (int st = 0; st < plan.nrad; ++st) {
        const int r = plan.rad[st];
        const int m = n / r;
        const int wstep = L / r;  // omega_r = W_L^{L/r}
        // output element e = (p, j, q): index q + s*(r*p + j), inputs q + s*(p + m*k)
        for (int e = slot; e < L; e += nslots) {
            const int q = e % s;
            const int pj = e / s;
            const int j = pj % r;
            const int p = pj / r;
            cplx acc = make_double2(0.0, 0.0);
            const int base = q + s * p;
            int widx = 0;  // (j*k*wstep) mod L
            const int winc = (int)(((long long)j * wstep) % L);
            for (int k = 0; k < r; ++k) {
                cplx w = tw[widx];
                if (inverse) w.y = -w.y;
                acc = cfma(x[(size_t)(base + s * m * k) * SB + site], w, acc);
                widx += winc;
                if (widx >= L) widx -= L;
            }
            // twiddle wp^j = exp(-2 pi i p j / n) = W_L^{p*j*s}
            const int tidx = (int)(((long long)p * j * s) % L);
            cplx t = tw[tidx];
            if (inverse) t.y = -t.y;
            y[(size_t)e * SB + site] = cmul(acc, t);
        }
        __syncthreads();
        cplx* tmp = x;
        x = y;
        y = tmp;
        n = m;
        s *= r;
    }
    return x;
}

// mode 0: tau_to_omega  (real in, complex out):  out = FFT(theta .* in)
// mode 1: omega_to_tau  (complex in, real out):  out = Re(conj(theta) .* iFFT(in))
// mode 2: fourier accelerate (real in, real out): out = Re(iFFT(diag^power .* FFT(in)))
template <int SB>
__global__ void __launch_bounds__(kT) fft_kernel(int mode, const double* __restrict__ rin, const cplx* __restrict__ cin,
                                                 double* __restrict__ rout, cplx* __restrict__ cout, FftPlan plan, int N,
                                                 const cplx* __restrict__ tw, const cplx* __restrict__ theta,
                                                 const double* __restrict__ diag, double power, const int* skip) {
    extern __shared__ double smem_raw[];
    if (skip && *skip) return;
    const int L = plan.L;
    cplx* b0 = reinterpret_cast<cplx*>(smem_raw);
    cplx* b1 = b0 + (size_t)L * SB;
    const int site = threadIdx.x % SB;
    const int slot = threadIdx.x / SB;
    const int nslots = blockDim.x / SB;
    const int gsite = blockIdx.x * SB + site;
    const bool ok = gsite < N;
    const double invL = 1.0 / (double)L;

    for (int t = slot; t < L; t += nslots) {
        cplx v = make_double2(0.0, 0.0);
        if (ok) {
            if (mode == 0) {
                const double xr = rin[(size_t)t * N + gsite];
                const cplx th = theta[t];
                v = make_double2(th.x * xr, th.y * xr);
            } else if (mode == 1) {
                v = cin[(size_t)t * N + gsite];
            } else {
                v = make_double2(rin[(size_t)t * N + gsite], 0.0);
            }
        }
        b0[(size_t)t * SB + site] = v;
    }
    __syncthreads();
    cplx* res = fft_smem<SB>(b0, b1, plan, tw, mode == 1);
    if (mode == 0) {
        for (int t = slot; t < L; t += nslots)
            if (ok) cout[(size_t)t * N + gsite] = res[(size_t)t * SB + site];
        return;
    }
    if (mode == 1) {
        for (int t = slot; t < L; t += nslots) {
            if (ok) {
                const cplx v = res[(size_t)t * SB + site];
                const cplx th = theta[t];  // conj(theta) * v, real part
                rout[(size_t)t * N + gsite] = (th.x * v.x + th.y * v.y) * invL;
            }
        }
        return;
    }
    // mode 2: scale in frequency space, inverse transform, real part
    cplx* other = (res == b0) ? b1 : b0;
    for (int t = slot; t < L; t += nslots) {
        double d = ok ? diag[(size_t)t * N + gsite] : 1.0;
        double f;
        if (power == 1.0) f = d;
        else if (power == 0.5) f = sqrt(d);
        else if (power == -1.0) f = 1.0 / d;
        else if (power == -0.5) f = 1.0 / sqrt(d);
        else f = pow(d, power);
        cplx v = res[(size_t)t * SB + site];
        res[(size_t)t * SB + site] = make_double2(v.x * f, v.y * f);
    }
    __syncthreads();
    cplx* res2 = fft_smem<SB>(res, other, plan, tw, true);
    for (int t = slot; t < L; t += nslots)
        if (ok) rout[(size_t)t * N + gsite] = res2[(size_t)t * SB + site].x * invL;
}

FftPlan make_plan(const elph_handle* h) {
    FftPlan p;
    p.L = h->L;
    p.nrad = (int)h->fft_radices.size();
    for (int i = 0; i < p.nrad; ++i) p.rad[i] = h->fft_radices[i];
    return p;
}

template <int SB>
void launch_fft(elph_handle* h, int mode, int ncols, const double* rin, const cplx* cin, double* rout, cplx* cout,
                const double* diag, double power, const int* skip) {
    const size_t smem = 2ull * h->L * SB * sizeof(cplx);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Ltau too large for the shared-memory FFT");
    if (smem > 48 * 1024)
        ELPH_CUDA(cudaFuncSetAttribute(fft_kernel<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
    const int blocks = (ncols + SB - 1) / SB;
    fft_kernel<SB><<<blocks, kT, smem, h->stream>>>(mode, rin, cin, rout, cout, make_plan(h), ncols, h->d_twiddle, h->d_theta,
                                                     diag, power, skip);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void dispatch_fft(elph_handle* h, int mode, int ncols, const double* rin, const cplx* cin, double* rout, cplx* cout,
                  const double* diag, double power, const int* skip = nullptr) {
    // fewer sites per CTA -> more CTAs; keep at least ~one CTA per SM when possible, bounded by shared memory
    int sb = 32;
    while (sb > 8 && (ncols + sb - 1) / sb < h->sm_count) sb >>= 1;
    while (sb > 8 && 2ull * h->L * sb * sizeof(cplx) > h->smem_optin) sb >>= 1;
    switch (sb) {
        case 32: launch_fft<32>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
        case 16: launch_fft<16>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
        default: launch_fft<8>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
    }
}

}  // namespace

void elph_fft_init(elph_handle* h) {
    const int L = h->L;
    // factor L: 4 first, then 2, 3, 5, then remaining primes
    std::vector<int> rad;
    int n = L;
    while (n % 4 == 0) { rad.push_back(4); n /= 4; }
    for (int f = 2; (long long)f * f <= n; ++f)
        while (n % f == 0) { rad.push_back(f); n /= f; }
    if (n > 1) rad.push_back(n);
    ELPH_REQUIRE((int)rad.size() <= kMaxRad, ELPH_ERR_UNSUPPORTED, "Ltau has too many prime factors");
    h->fft_radices = rad;
    std::vector<cplx> tw(L), th(L);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < L; ++k) {
        // exact-argument reduction: angle = -2 pi k / L
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)L;
        tw[k] = make_double2((double)cosl(a), (double)sinl(a));
        const long double b = -3.14159265358979323846264338327950288L * (long double)k / (long double)L;
        th[k] = make_double2((double)cosl(b), (double)sinl(b));
    }
    (void)pi;
    h->d_twiddle = elph_dalloc<cplx>(L);
    h->d_theta = elph_dalloc<cplx>(L);
    ELPH_CUDA(cudaMemcpy(h->d_twiddle, tw.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
    ELPH_CUDA(cudaMemcpy(h->d_theta, th.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
}

void elph_tau_to_omega_dev_skip(elph_handle* h, const double* vin, cplx* vout, const int* skip) {
    dispatch_fft(h, 0, h->N, vin, nullptr, nullptr, vout, nullptr, 0.0, skip);
}
void elph_omega_to_tau_dev_skip(elph_handle* h, const cplx* vin, double* vout, const int* skip) {
    dispatch_fft(h, 1, h->N, nullptr, vin, vout, nullptr, nullptr, 0.0, skip);
}

void elph_tau_to_omega_dev(elph_handle* h, const double* vin, cplx* vout) {
    dispatch_fft(h, 0, h->N, vin, nullptr, nullptr, vout, nullptr, 0.0);
}

void elph_omega_to_tau_dev(elph_handle* h, const cplx* vin, double* vout) {
    dispatch_fft(h, 1, h->N, nullptr, vin, vout, nullptr, nullptr, 0.0);
}

void elph_fourier_accelerate_dev(elph_handle* h, const double* vin, double* vout, double power, bool use_mass) {
    const double* diag = use_mass ? h->d_Mass : h->d_Q;
    ELPH_REQUIRE(use_mass ? h->have_M : h->have_Q, ELPH_ERR_STATE, "fourier acceleration diagonal (fa_Q / fa_M) was not provided");
    dispatch_fft(h, 2, h->Nph, vin, nullptr, vout, nullptr, diag, power);
}
