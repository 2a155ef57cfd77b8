// K6: fermion force  <dM/dx> = u^T (dM/dx) v  and the bosonic action / gradient.
//
// Reference: muldMdx! Holstein src/HolsteinModels.jl:691-755, SSH src/SSHModels.jl:707-829;
// calc_dSbdx! src/PhononAction.jl:114-187 (Holstein), :189-233 (SSH); calc_Sb :11-66, :68-107;
// the Langevin force assembly dSdx = -2 <dM/dx> + dSb/dx(shifted) src/LangevinDynamics.jl:334-384.
//
// Holstein: one kernel.  A CTA stages a chunk of tau-slices of u in shared memory, applies the
// transposed checkerboard there, and multiplies by d(tau,i) = +-dtau (lam_i + 2 lam2_i x) expnV v(tau-1,i);
// the scale (-2) and the bosonic gradient are fused into the same pass.
//
// SSH: bonds inside a colour group commute, so the reference's bond-sequential recurrence
// (b <- Gamma_n b, c <- Gamma_n^-1 c, read off c_j b_i + c_i b_j) runs colour by colour in shared
// memory with b and c resident; every phonon owns exactly one bond, so the accumulation over
// `primary_field` is a gather over the phonons that share a primary (ordered, no atomics).
#include "elph_internal.cuh"

namespace {

constexpr int kT = 256;

struct FParams {
    const double* __restrict__ u;
    const double* __restrict__ v;
    double* __restrict__ out;
    const double* __restrict__ x;
    const double* __restrict__ D;
    const double* __restrict__ lam;
    const double* __restrict__ lam2;
    const double* __restrict__ omega;
    const double* __restrict__ omega4;
    const int2* __restrict__ bonds;
    const int* __restrict__ goff;
    const double2* __restrict__ cs;
    int ngroups, N, L, Nb, C;
    int open, tau0, Lglob;
    double dtau, scale;
    int add_dSb, shifted;
};

__global__ void __launch_bounds__(kT) holstein_force_kernel(FParams P) {
    extern __shared__ double smem[];
    const int N = P.N, L = P.L;
    const int a = blockIdx.x * P.C;
    const int nout = min(P.C, L - a);
    for (int k = 0; k < nout; ++k)
        for (int i = threadIdx.x; i < N; i += blockDim.x) smem[(size_t)k * N + i] = P.u[(size_t)(a + k) * N + i];
    __syncthreads();
    // y = K^T u  (src/HolsteinModels.jl:746-748)
    for (int g = P.ngroups - 1; g >= 0; --g) {
        const int lo = P.goff[g], hi = P.goff[g + 1];
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            const int2 ij = P.bonds[b];
            const double2 cs = P.cs[b];
            for (int k = 0; k < nout; ++k) {
                double* p = smem + (size_t)k * N;
                const double t1 = p[ij.x], t2 = p[ij.y];
                p[ij.x] = cs.x * t1 + cs.y * t2;
                p[ij.y] = cs.x * t2 + cs.y * t1;
            }
        }
        __syncthreads();
    }
    for (int k = 0; k < nout; ++k) {
        const int tau = a + k;
        // tau-sharded slab: v(tau-1) of the first own slice is the left halo (index -1); '-' sign on GLOBAL slice 0
        const long long taum = (tau == 0) ? (P.open ? -1 : L - 1) : tau - 1;
        const bool flip = ((P.tau0 + tau) % P.Lglob) == 0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const size_t idx = (size_t)tau * N + i;
            const double xt = P.x[idx];
            double d = P.dtau * (P.lam[i] + 2.0 * P.lam2[i] * xt) * P.D[idx] * P.v[taum * N + i];
            if (flip) d = -d;
            double r = P.scale * (smem[(size_t)k * N + i] * d);
            if (P.add_dSb)
                r += dSb_term(P.x, tau, i, N, L, P.dtau, P.omega[i], P.omega4[i], P.shifted ? P.dtau * P.lam[i] : 0.0);
            P.out[idx] = r;
        }
    }
}

// dSbdx += dSb/dx   (accumulating, like the reference)
__global__ void dSb_kernel(double* __restrict__ dS, const double* __restrict__ x, const double* __restrict__ omega,
                           const double* __restrict__ omega4, const double* __restrict__ lam, int ncols, int L, double dtau,
                           int shifted_holstein) {
    const long long n = (long long)ncols * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ncols);
        const int tau = (int)(idx / ncols);
        const double ls = shifted_holstein ? dtau * lam[i] : 0.0;
        dS[idx] += dSb_term(x, tau, i, ncols, L, dtau, omega[i], omega4[i], ls);
    }
}

// tau-sharded slab: x points at the first own slice of a halo'd field ([x(a-1)][own ...][x(b)]); the bosonic action is
// periodic in tau (no sign), so the ring closure is an ordinary halo
__global__ void dSb_open_kernel(double* __restrict__ dS, const double* __restrict__ x, const double* __restrict__ omega,
                                const double* __restrict__ omega4, const double* __restrict__ lam, int ncols, int L, double dtau,
                                int shifted_holstein) {
    const long long n = (long long)ncols * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ncols);
        const double w = omega[i], w4 = omega4[i];
        const double xt = x[idx];
        double d = dtau * w * w * xt - (shifted_holstein ? dtau * lam[i] : 0.0);
        d += dtau * 4.0 * w4 * xt * xt * xt;
        d -= (x[idx + ncols] + x[idx - ncols] - 2.0 * xt) / dtau;
        dS[idx] += d;
    }
}

// Sb partial sums (per CTA), folded on the host in index order
__global__ void __launch_bounds__(kT) Sb_kernel(const double* __restrict__ x, const double* __restrict__ omega,
                                                const double* __restrict__ omega4, const double* __restrict__ lam,
                                                const int* __restrict__ primary_ph, int ncols, int L, double dtau, int holstein,
                                                int shifted, double* __restrict__ partial) {
    __shared__ double red[32];
    const long long n = (long long)ncols * L;
    double s = 0.0;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ncols);
        const int tau = (int)(idx / ncols);
        const int tm = (tau == 0) ? L - 1 : tau - 1;
        const double xt = x[idx];
        const double xm = x[(size_t)tm * ncols + i];
        const double w = omega[i];
        if (holstein) {
            // dtau * [ w^2 x^2/2 + w4 x^4 - lam x shifted + (x-x-)^2/dtau^2/2 ]   (src/PhononAction.jl:23-38,63)
            double t = w * w * xt * xt / 2 + omega4[i] * xt * xt * xt * xt;
            if (shifted) t -= lam[i] * xt;
            t += (xt - xm) * (xt - xm) / (dtau * dtau) / 2;
            s += dtau * t;
        } else if (primary_ph[i] == i) {
            // only primary phonons (src/PhononAction.jl:79-103)
            s += dtau * w * w * xt * xt / 2 + dtau * omega4[i] * xt * xt * xt * xt + (xt - xm) * (xt - xm) / dtau / 2;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        partial[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------- SSH force
struct SParams {
    const double* __restrict__ u;
    const double* __restrict__ v;
    double* __restrict__ raw;       // [L][Nph] per-phonon <dM/dx> before the primary-field fold
    const double* __restrict__ x;
    const double* __restrict__ expmu;
    const double* __restrict__ alpha;
    const double* __restrict__ alpha2;
    const int2* __restrict__ bonds;
    const int* __restrict__ goff;
    const double2* __restrict__ cs;  // [L][Nb]
    const int* __restrict__ col_ph;  // column -> phonon or -1
    int ngroups, N, L, Nb, Nph, C;
    int open, tau0, Lglob;           // tau-sharded slab: v(tau-1) of the first own slice is the left halo, sign on GLOBAL slice 0
    double dtau;
};

__global__ void __launch_bounds__(kT) ssh_force_kernel(SParams P) {
    extern __shared__ double smem[];
    const int N = P.N, L = P.L;
    const int a = blockIdx.x * P.C;
    const int nout = min(P.C, L - a);
    double* B = smem;                          // b(tau) = expmu .* v(tau-1)
    double* Cc = smem + (size_t)P.C * N;       // c(tau) = K^T(tau) u(tau)
    for (int k = 0; k < nout; ++k) {
        const int tau = a + k;
        const long long taum = (tau == 0) ? (P.open ? -1 : L - 1) : tau - 1;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            B[(size_t)k * N + i] = P.expmu[i] * P.v[taum * N + i];
            Cc[(size_t)k * N + i] = P.u[(size_t)tau * N + i];
        }
    }
    __syncthreads();
    // c0 = K^T u : reverse colour order (src/SSHModels.jl:757-759)
    for (int g = P.ngroups - 1; g >= 0; --g) {
        const int lo = P.goff[g], hi = P.goff[g + 1];
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            const int2 ij = P.bonds[b];
            for (int k = 0; k < nout; ++k) {
                const double2 cs = P.cs[(size_t)(a + k) * P.Nb + b];
                double* p = Cc + (size_t)k * N;
                const double t1 = p[ij.x], t2 = p[ij.y];
                p[ij.x] = cs.x * t1 + cs.y * t2;
                p[ij.y] = cs.x * t2 + cs.y * t1;
            }
        }
        __syncthreads();
    }
    // forward over bonds: b <- Gamma_n b, c <- Gamma_n^-1 c, then read off the matrix element (:765-822)
    for (int g = 0; g < P.ngroups; ++g) {
        const int lo = P.goff[g], hi = P.goff[g + 1];
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            const int2 ij = P.bonds[b];
            const int ph = P.col_ph[b];
            for (int k = 0; k < nout; ++k) {
                const int tau = a + k;
                const double2 cs = P.cs[(size_t)tau * P.Nb + b];
                double* pb = B + (size_t)k * N;
                double* pc = Cc + (size_t)k * N;
                const double bi = pb[ij.x], bj = pb[ij.y];
                const double nbi = cs.x * bi + cs.y * bj;
                const double nbj = cs.x * bj + cs.y * bi;
                pb[ij.x] = nbi;
                pb[ij.y] = nbj;
                const double ci = pc[ij.x], cj = pc[ij.y];
                const double nci = cs.x * ci - cs.y * cj;
                const double ncj = cs.x * cj - cs.y * ci;
                pc[ij.x] = nci;
                pc[ij.y] = ncj;
                if (ph >= 0) {
                    const double xn = P.x[(size_t)tau * P.Nph + ph];
                    const double dK = P.alpha[ph] + 2.0 * P.alpha2[ph] * xn;
                    double dm = ncj * P.dtau * dK * nbi + (nci * P.dtau * dK) * nbj;
                    if (((P.tau0 + tau) % P.Lglob) == 0) dm = -dm;
                    P.raw[(size_t)tau * P.Nph + ph] = dm;
                }
            }
        }
        __syncthreads();
    }
}

// dMdx[field] = sum over phonons q with primary(q) == primary(field) of raw[q], in increasing q
// (the reference accumulates into dMdx[primary_field[field]] in bond order and then gathers, :820,:826)
__global__ void ssh_fold_kernel(const double* __restrict__ raw, double* __restrict__ out, const int* __restrict__ primary_ph,
                                const int* __restrict__ grp_start, const int* __restrict__ grp_members, int Nph, int L,
                                double scale, int add_dSb, const double* __restrict__ x, const double* __restrict__ omega,
                                const double* __restrict__ omega4, double dtau) {
    const long long n = (long long)Nph * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int ph = (int)(idx % Nph);
        const int tau = (int)(idx / Nph);
        const int pr = primary_ph[ph];
        double s = 0.0;
        for (int m = grp_start[pr]; m < grp_start[pr + 1]; ++m) s += raw[(size_t)tau * Nph + grp_members[m]];
        double r = scale * s;
        if (add_dSb) r += dSb_term(x, tau, ph, Nph, L, dtau, omega[ph], omega4[ph], 0.0);
        out[idx] = r;
    }
}

}  // namespace

void elph_muldMdx_dev(elph_handle* h, const double* u, const double* v, double* out, double scale, bool add_dSb, bool shifted) {
    const size_t slice = (size_t)h->N * sizeof(double);
    if (h->model == ELPH_MODEL_HOLSTEIN) {
        FParams P;
        P.u = u; P.v = v; P.out = out; P.x = h->d_x; P.D = h->d_D; P.lam = h->d_lam; P.lam2 = h->d_lam2;
        P.omega = h->d_omega; P.omega4 = h->d_omega4; P.bonds = h->d_bonds; P.goff = h->d_goff; P.cs = h->d_cs;
        P.ngroups = h->ngroups; P.N = h->N; P.L = h->L; P.Nb = h->Nb;
        P.open = h->sharded ? 1 : 0; P.tau0 = h->sharded ? h->shard_tau0 : 0; P.Lglob = h->sharded ? h->shard_Lglob : h->L;
        ELPH_REQUIRE(!(h->sharded && add_dSb), ELPH_ERR_UNSUPPORTED,
                     "tau-sharded force: add the bosonic gradient separately (it needs x halos)");
        int C = 1;
        for (int c : {4, 2}) if ((h->L + c - 1) / c >= 2 * h->sm_count && c * slice <= 96 * 1024) { C = c; break; }
        ELPH_REQUIRE(C * slice <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Nsites too large for the shared-memory force kernel");
        P.C = C; P.dtau = h->dtau; P.scale = scale; P.add_dSb = add_dSb ? 1 : 0; P.shifted = shifted ? 1 : 0;
        elph_enable_smem(h, holstein_force_kernel);
        holstein_force_kernel<<<(h->L + C - 1) / C, kT, C * slice, h->stream>>>(P);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        return;
    }
    // SSH
    SParams P;
    P.u = u; P.v = v; P.raw = h->d_tmp; P.x = h->d_x; P.expmu = h->d_D; P.alpha = h->d_alpha; P.alpha2 = h->d_alpha2;
    P.bonds = h->d_bonds; P.goff = h->d_goff; P.cs = h->d_cs; P.col_ph = h->d_col_ph;
    P.ngroups = h->ngroups; P.N = h->N; P.L = h->L; P.Nb = h->Nb; P.Nph = h->Nph; P.C = 1; P.dtau = h->dtau;
    P.open = h->sharded ? 1 : 0; P.tau0 = h->sharded ? h->shard_tau0 : 0; P.Lglob = h->sharded ? h->shard_Lglob : h->L;
    ELPH_REQUIRE(!(h->sharded && add_dSb), ELPH_ERR_UNSUPPORTED, "tau-sharded force: add the bosonic gradient separately (it needs x halos)");
    ELPH_REQUIRE(2 * slice <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Nsites too large for the shared-memory force kernel");
    ELPH_CUDA(cudaMemsetAsync(h->d_tmp, 0, h->Ndof * sizeof(double), h->stream));
    elph_enable_smem(h, ssh_force_kernel);
    ssh_force_kernel<<<h->L, kT, 2 * slice, h->stream>>>(P);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    const int blocks = (int)std::min<int64_t>((h->Ndof + kT - 1) / kT, 8LL * h->sm_count);
    ssh_fold_kernel<<<blocks, kT, 0, h->stream>>>(h->d_tmp, out, h->d_primary_ph, h->d_grp_start, h->d_grp_members, h->Nph, h->L,
                                                  scale, add_dSb ? 1 : 0, h->d_x, h->d_omega, h->d_omega4, h->dtau);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_dSbdx_dev(elph_handle* h, double* dSbdx, bool shifted) {
    const int blocks = (int)std::min<int64_t>((h->Ndof + kT - 1) / kT, 8LL * h->sm_count);
    const int sh = (shifted && h->model == ELPH_MODEL_HOLSTEIN) ? 1 : 0;
    dSb_kernel<<<blocks, kT, 0, h->stream>>>(dSbdx, h->d_x, h->d_omega, h->d_omega4, h->d_lam, h->Nph, h->L, h->dtau, sh);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_dSbdx_open_dev(elph_handle* h, double* dSbdx, const double* x_own, bool shifted) {
    const int blocks = (int)std::min<int64_t>((h->Ndof + kT - 1) / kT, 8LL * h->sm_count);
    const int sh = (shifted && h->model == ELPH_MODEL_HOLSTEIN) ? 1 : 0;
    dSb_open_kernel<<<blocks, kT, 0, h->stream>>>(dSbdx, x_own, h->d_omega, h->d_omega4, h->d_lam, h->Nph, h->L, h->dtau, sh);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_Sb_dev(elph_handle* h, bool shifted, double* host_out) {
    const int blocks = std::min(h->partial_cap, 2 * h->sm_count);
    Sb_kernel<<<blocks, kT, 0, h->stream>>>(h->d_x, h->d_omega, h->d_omega4, h->d_lam, h->d_primary_ph, h->Nph, h->L, h->dtau,
                                            h->model == ELPH_MODEL_HOLSTEIN ? 1 : 0, shifted ? 1 : 0, h->d_partial);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    std::vector<double> part(blocks);
    ELPH_CUDA(cudaMemcpyAsync(part.data(), h->d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    double s = 0.0;
    for (int k = 0; k < blocks; ++k) s += part[k];
    *host_out = s;
}
