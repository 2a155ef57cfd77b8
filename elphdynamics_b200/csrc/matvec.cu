// K1 + K2: field -> operator tables, and the fused fermion-matrix products
//   y = M v, y = M^T v, y = M^T M v
// with M the (N*Ltau)x(N*Ltau) space-time matrix of the reference:
//   (M v)(1)   = v(1)   + B(1) v(Ltau),   (M v)(tau)  = v(tau) - B(tau) v(tau-1)
//   (M^T v)(L) = v(L)   + B^T(1) v(1),    (M^T v)(tau)= v(tau) - B^T(tau+1) v(tau+1)
//   B(tau) = K(tau) diag(D(tau)),  K = ordered product of 2x2 bond rotations
// Reference: src/HolsteinModels.jl:526-549,569-684; src/SSHModels.jl:510-562,581-701;
//            src/Checkerboard.jl:57-230; src/Models.jl:215-224.
//
// Engine layout: vectors are [tau][site] (tau-slice-major), so one tau-slice is a
// contiguous N-vector.  A CTA owns a chunk of C output slices; it stages the
// slices it needs in shared memory, applies all colour groups of the checkerboard
// there (bonds inside a colour touch disjoint sites, so a colour is one parallel
// step; colours are separated by __syncthreads), and streams the result out.
// The fused M^T M kernel recomputes the one extra M-slice it needs at the chunk
// edge instead of round-tripping the intermediate M v through HBM: compulsory
// traffic is read v + read D + write y = 24 B per (site,tau) point.
#include "elph_internal.cuh"

namespace {

constexpr int kThreads = 256;

struct KParams {
    const double* __restrict__ v;
    double* __restrict__ y;
    const double* __restrict__ D;     // Holstein: expnV [L][N]; SSH: expmu [N]
    const int2* __restrict__ bonds;   // [Nb]
    const int* __restrict__ goff;     // [ngroups+1]
    const double2* __restrict__ cs;   // Holstein [Nb]; SSH [L][Nb]
    double* __restrict__ partial;     // optional per-CTA partial of dot(v,y)
    // CG fusion (MODE_MTM only): v := pr + beta*pold computed on the fly, written to pnew for the
    // CTA's own slices; the last CTA to finish folds the partials into S->pAp / S->alpha.
    const double* __restrict__ pr;
    const double* __restrict__ pold;
    double* __restrict__ pnew;
    CgScalars* S;
    unsigned int* ticket;
    int ngroups;
    int N, L, Nb, C;
    int open, tau0, Lglob;   // tau-sharded slab (halo slices at index -1 / L) instead of the periodic wrap
    long long v_stride, y_stride, D_stride;
};

__device__ __forceinline__ void rotate(double* __restrict__ p, int i, int j, double c, double s) {
    const double t1 = p[i];
    const double t2 = p[j];
    p[i] = c * t1 + s * t2;
    p[j] = c * t2 + s * t1;
}

// Apply the colour groups to `nsl` consecutive slices held in shared memory.
// tau0 = imaginary-time index of buf slice 0 (only used for the per-tau SSH tables).
template <bool SSH, bool REVERSE>
__device__ __forceinline__ void sweep_smem(double* __restrict__ buf, int nsl, int tau0, const KParams& P) {
    const int N = P.N;
    for (int gg = 0; gg < P.ngroups; ++gg) {
        const int g = REVERSE ? (P.ngroups - 1 - gg) : gg;
        const int lo = P.goff[g], hi = P.goff[g + 1];
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            const int2 ij = P.bonds[b];
            if (!SSH) {
                const double2 cs = P.cs[b];
#pragma unroll 4
                for (int k = 0; k < nsl; ++k) rotate(buf + (size_t)k * N, ij.x, ij.y, cs.x, cs.y);
            } else {
                for (int k = 0; k < nsl; ++k) {
                    int tau = tau0 + k;
                    if (tau >= P.L && !P.open) tau -= P.L;   // open slab: row L of the table is the right neighbour's first slice
                    const double2 cs = P.cs[(size_t)tau * P.Nb + b];
                    rotate(buf + (size_t)k * N, ij.x, ij.y, cs.x, cs.y);
                }
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double block_sum(double x, double* red) {
    // deterministic: fixed shuffle tree, then warp 0 sums the warp partials in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = x;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int k = 0; k < nw; ++k) t += red[k];
    }
    return t;  // valid on thread 0
}

template <int MODE, bool SSH, bool FUSEP>
__global__ void __launch_bounds__(kThreads) matvec_kernel(KParams P) {
    extern __shared__ double smem[];
    __shared__ double red[32];
    __shared__ bool is_last;
    const int N = P.N, L = P.L;
    double beta = 0.0;
    if (FUSEP) {
        if (P.S->done) return;  // convergence latch: later launches of the chunk are no-ops
        beta = P.S->beta;
    }
    const int a = blockIdx.x * P.C;
    const int nout = min(P.C, L - a);
    const double* __restrict__ v = P.v + (size_t)blockIdx.y * P.v_stride;
    double* __restrict__ y = P.y + (size_t)blockIdx.y * P.y_stride;
    const double* __restrict__ D = P.D + (size_t)blockIdx.y * P.D_stride;
    double acc = 0.0;
    auto loadv = [&](long long idx) -> double { return FUSEP ? fma(beta, P.pold[idx], P.pr[idx]) : v[idx]; };
    // memory slice index of logical local slice t in [-1, L]: periodic wrap, or (tau-sharded slab) halo slices
    auto midx = [&](int t) -> long long { return P.open ? (long long)t : (long long)(t < 0 ? t + L : (t >= L ? t - L : t)); };
    // the antiperiodic '+' sign belongs to GLOBAL time slice 0
    auto wrapsign = [&](int t) -> bool { return ((P.tau0 + t + P.Lglob) % P.Lglob) == 0; };

    if (MODE == MODE_M) {
        // A[k] = D(tau) .* v(tau-1), tau = a+k
        for (int k = 0; k < nout; ++k) {
            const int tau = a + k;
            const long long taum = midx(tau - 1);
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double d = SSH ? D[i] : D[(size_t)tau * N + i];
                smem[(size_t)k * N + i] = d * v[taum * N + i];
            }
        }
        __syncthreads();
        sweep_smem<SSH, false>(smem, nout, a, P);
        for (int k = 0; k < nout; ++k) {
            const int tau = a + k;
            const bool plus = wrapsign(tau);
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double vv = v[(size_t)tau * N + i];
                const double bv = smem[(size_t)k * N + i];
                const double r = plus ? (vv + bv) : (vv - bv);
                y[(size_t)tau * N + i] = r;
                acc += vv * r;
            }
        }
    } else if (MODE == MODE_MT) {
        // A[k] = v(tau'), u = K^T A ; y(tau'-1) = v(tau'-1) -/+ D(tau') u.  Periodic: tau' = a+k covers 0..L-1 (output
        // slice tau'-1 wraps); open slab: tau' = a+k+1 covers 1..L (slice L = right halo), outputs 0..L-1.
        const int shift = P.open ? 1 : 0;
        for (int k = 0; k < nout; ++k) {
            const int tau = a + k + shift;
            for (int i = threadIdx.x; i < N; i += blockDim.x) smem[(size_t)k * N + i] = v[(size_t)tau * N + i];
        }
        __syncthreads();
        sweep_smem<SSH, true>(smem, nout, a + shift, P);
        for (int k = 0; k < nout; ++k) {
            const int tau = a + k + shift;
            const long long taum = midx(tau - 1);
            const bool plus = wrapsign(tau);
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double d = SSH ? D[i] : D[(size_t)tau * N + i];
                const double vv = v[taum * N + i];
                const double bu = d * smem[(size_t)k * N + i];
                const double r = plus ? (vv + bu) : (vv - bu);
                y[taum * N + i] = r;
                acc += vv * r;
            }
        }
    } else {
        // fused M^T M.  w slices tau = a .. a+nout (mod L), outputs tau = a .. a+nout-1.
        const int nw = nout + 1;
        double* A = smem;                        // nw slices
        double* W = smem + (size_t)(P.C + 1) * N;  // copies of w(a+1 .. a+nout-1)
        for (int k = 0; k < nw; ++k) {
            const long long tau = midx(a + k);
            const long long taum = midx(a + k - 1);
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double d = SSH ? D[i] : D[tau * N + i];
                A[(size_t)k * N + i] = d * loadv(taum * N + i);
            }
        }
        __syncthreads();
        sweep_smem<SSH, false>(A, nw, a, P);
        // w = v -/+ B v(tau-1); same thread reads and writes element (k,i): no barrier needed before
        for (int k = 0; k < nw; ++k) {
            const long long tau = midx(a + k);
            const bool plus = wrapsign(a + k);
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double vv = loadv(tau * N + i);
                if (FUSEP && k < nout) P.pnew[tau * N + i] = vv;
                const double bv = A[(size_t)k * N + i];
                const double w = plus ? (vv + bv) : (vv - bv);
                A[(size_t)k * N + i] = w;
                if (k >= 1 && k < nout) W[(size_t)(k - 1) * N + i] = w;
            }
        }
        __syncthreads();
        sweep_smem<SSH, true>(A + N, nout, a + 1, P);  // u(tau) = K^T(tau) w(tau), tau = a+1 .. a+nout
        for (int k = 0; k < nout; ++k) {
            const int tau = a + k;
            const long long taup = midx(tau + 1);
            const bool plus = wrapsign(tau + 1);
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const double d = SSH ? D[i] : D[taup * N + i];
                const double w = (k == 0) ? A[i] : W[(size_t)(k - 1) * N + i];
                const double bu = d * A[(size_t)(k + 1) * N + i];
                const double r = plus ? (w + bu) : (w - bu);
                y[(size_t)tau * N + i] = r;
                if (P.partial) acc += loadv((size_t)tau * N + i) * r;
            }
        }
    }
    if (P.partial) {
        const double t = block_sum(acc, red);
        if (threadIdx.x == 0) P.partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
    if (FUSEP) {
        // last-CTA-done: fixed-order (index, not arrival) reduction of the per-CTA partials
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int n = atomicAdd(P.ticket, 1u);
            is_last = (n == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            double s = 0.0;
            for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += ((volatile double*)P.partial)[k];
            const double pAp = block_sum(s, red);
            if (threadIdx.x == 0) {
                P.S->pAp = pAp;
                P.S->alpha = P.S->rdotz / pAp;
                *P.ticket = 0u;
            }
        }
    }
}

// expnV[tau][i] = exp(-dtau*(lam_i x + lam2_i x^2 - mu_i))      src/HolsteinModels.jl:526-549
__global__ void holstein_update_kernel(const double* __restrict__ x, const double* __restrict__ lam,
                                       const double* __restrict__ lam2, const double* __restrict__ mu,
                                       double* __restrict__ expnV, int N, long long n, double dtau) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % N);
        const double xv = x[idx];
        expnV[idx] = exp(-dtau * (lam[i] * xv + lam2[i] * xv * xv + -mu[i]));
    }
}

// SSH: t' = t - (alpha x + sign(x) alpha2 x^2); cosh/sinh(dtau t') per (tau, column)   src/SSHModels.jl:510-540
__global__ void ssh_update_kernel(const double* __restrict__ x, const double* __restrict__ t, const double* __restrict__ alpha,
                                  const double* __restrict__ alpha2, const int* __restrict__ col_ph,
                                  const int* __restrict__ col_bond, double2* __restrict__ cs, double* __restrict__ tprime,
                                  const int* __restrict__ sq_slot, double2* __restrict__ sq_tab, int Nb, int Nph, long long n,
                                  double dtau, long long x_stride = 0, long long tab_stride = 0) {
    // blockIdx.y = replica (elph_dev_ssh_replica_tables): own phonon field, own tile-layout table, no column-layout copies
    x += (size_t)blockIdx.y * x_stride;
    if (sq_tab) sq_tab += (size_t)blockIdx.y * tab_stride;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(idx % Nb);
        const long long tau = idx / Nb;
        const int ph = col_ph[col];
        double tp = t[col_bond[col]];
        if (ph >= 0) {
            const double xv = x[tau * Nph + ph];
            const double sgn = (xv > 0.0) ? 1.0 : ((xv < 0.0) ? -1.0 : 0.0);
            tp -= alpha[ph] * xv + sgn * alpha2[ph] * xv * xv;
        }
        if (tprime) tprime[idx] = tp;
        const double2 v = make_double2(cosh(dtau * tp), sinh(dtau * tp));
        if (cs) cs[idx] = v;
        if (sq_tab) sq_tab[tau * Nb + sq_slot[col]] = v;   // tile layout [tau][dir][site] of ssh_square.cu
    }
}

__global__ void expmu_kernel(const double* __restrict__ mu, double* __restrict__ out, int N, double dtau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = exp(dtau * mu[i]);
}

// out[c][r] = in[r][c], tiled through shared memory (both sides coalesced)
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
    __shared__ T tile[32][33];
    const size_t boff = (size_t)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[dy][threadIdx.x] = in[boff + (size_t)r * cols + c];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int c = c0 + dy, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = tile[threadIdx.x][dy];
    }
}

template <int MODE, bool SSH, bool FUSEP = false>
void launch_one(elph_handle* h, const KParams& P, dim3 grid, size_t smem) {
    elph_enable_smem(h, matvec_kernel<MODE, SSH, FUSEP>);
    matvec_kernel<MODE, SSH, FUSEP><<<grid, kThreads, smem, h->stream>>>(P);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace

static int pick_chunk(elph_handle* h, MatvecMode mode, int64_t nbatch) {
    const size_t slice = (size_t)h->N * sizeof(double);
    auto smem_for = [&](int C) { return (mode == MODE_MTM) ? (size_t)(2 * C) * slice : (size_t)C * slice; };
    int C;
    if (h->chunk_override > 0) {
        C = h->chunk_override;
    } else {
        // enough CTAs to cover the machine twice, otherwise favour larger chunks (less halo recompute)
        C = 1;
        const int64_t want = 2LL * h->sm_count;
        for (int c : {8, 6, 4, 3, 2}) {
            const int64_t ctas = nbatch * ((h->L + c - 1) / c);
            if (ctas >= want && smem_for(c) <= 96 * 1024) { C = c; break; }
        }
    }
    if (C > h->L) C = h->L;
    while (C > 1 && smem_for(C) > h->smem_optin) --C;
    ELPH_REQUIRE(smem_for(C) <= h->smem_optin, ELPH_ERR_UNSUPPORTED,
                 "Nsites too large for the shared-memory slice kernels (one tau-slice must fit in shared memory)");
    return C;
}

void elph_launch_matvec(elph_handle* h, MatvecMode mode, const MatvecArgs& a) {
    if (mode == MODE_MTM && (elph_launch_mtm_square(h, a) || elph_launch_ssh_square(h, a))) return;
    ELPH_REQUIRE(a.v != a.y || a.cg_S, ELPH_ERR_INVALID, "matvec output must not alias its input");
    KParams P;
    P.v = a.v;
    P.y = a.y;
    P.D = a.D ? a.D : h->d_D;
    P.bonds = h->d_bonds;
    P.goff = h->d_goff;
    P.cs = h->d_cs;
    P.partial = a.partial_dot;
    P.pr = a.cg_pr;
    P.pold = a.cg_pold;
    P.pnew = a.cg_pnew;
    P.S = a.cg_S;
    P.ticket = a.cg_ticket;
    const bool fusep = (a.cg_S != nullptr);
    ELPH_REQUIRE(!fusep || (mode == MODE_MTM && a.nbatch == 1 && a.partial_dot), ELPH_ERR_INVALID,
                 "CG fusion is only available for the single-vector M^T M product");
    P.ngroups = h->ngroups;
    P.open = a.open ? 1 : 0;
    P.tau0 = a.open ? h->shard_tau0 : 0;
    P.Lglob = a.open ? h->shard_Lglob : h->L;
    P.N = h->N;
    P.L = h->L;
    P.Nb = h->Nb;
    P.C = pick_chunk(h, mode, a.nbatch);
    P.v_stride = a.v_stride;
    P.y_stride = a.y_stride;
    P.D_stride = a.D_stride;
    const int nchunks = (h->L + P.C - 1) / P.C;
    ELPH_REQUIRE(a.nbatch >= 1 && a.nbatch <= 65535, ELPH_ERR_INVALID, "batch count out of range");
    dim3 grid(nchunks, (unsigned)a.nbatch);
    if (a.partial_dot) {
        ELPH_REQUIRE((int64_t)nchunks * a.nbatch <= h->partial_cap, ELPH_ERR_INVALID, "partial buffer too small");
        if (a.npartial) *a.npartial = nchunks;
    }
    const size_t slice = (size_t)h->N * sizeof(double);
    const size_t smem = (mode == MODE_MTM) ? (size_t)(2 * P.C) * slice : (size_t)P.C * slice;
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    switch (mode) {
        case MODE_M:
            ssh ? launch_one<MODE_M, true>(h, P, grid, smem) : launch_one<MODE_M, false>(h, P, grid, smem);
            break;
        case MODE_MT:
            ssh ? launch_one<MODE_MT, true>(h, P, grid, smem) : launch_one<MODE_MT, false>(h, P, grid, smem);
            break;
        case MODE_MTM:
            if (fusep)
                ssh ? launch_one<MODE_MTM, true, true>(h, P, grid, smem) : launch_one<MODE_MTM, false, true>(h, P, grid, smem);
            else
                ssh ? launch_one<MODE_MTM, true>(h, P, grid, smem) : launch_one<MODE_MTM, false>(h, P, grid, smem);
            break;
    }
}

void elph_launch_update_model(elph_handle* h) {
    const int T = 256;
    if (h->model == ELPH_MODEL_HOLSTEIN) {
        const long long n = h->Ndim;
        const int blocks = (int)std::min<long long>((n + T - 1) / T, 8LL * h->sm_count);
        holstein_update_kernel<<<blocks, T, 0, h->stream>>>(h->d_x, h->d_lam, h->d_lam2, h->d_mu, h->d_D, h->N, n, h->dtau);
    } else {
        expmu_kernel<<<(h->N + T - 1) / T, T, 0, h->stream>>>(h->d_mu, h->d_D, h->N, h->dtau);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        const long long n = (long long)h->L * h->Nb;
        const int blocks = (int)std::min<long long>((n + T - 1) / T, 8LL * h->sm_count);
        ssh_update_kernel<<<blocks, T, 0, h->stream>>>(h->d_x, h->d_t, h->d_alpha, h->d_alpha2, h->d_col_ph,
                                                       h->d_col_bond, h->d_cs, h->d_tprime, h->ssq.enabled ? h->ssq.d_slot : nullptr,
                                                       h->ssq.enabled ? h->ssq.d_tab : nullptr, h->Nb, h->Nph, n, h->dtau);
    }
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

// update_model! (src/SSHModels.jl:510-540) for nrep independent phonon fields at once, written straight into per-replica tables
// in the tile layout [tau][dir][site] that ssh_square_kernel stages with TMA
void elph_launch_ssh_replica_tables(elph_handle* h, int64_t nrep, const double* x_dev, int64_t x_stride, double2* tab_dev,
                                    int64_t tab_stride) {
    ELPH_REQUIRE(h->model == ELPH_MODEL_SSH && h->ssq.enabled, ELPH_ERR_UNSUPPORTED,
                 "replica tables need the SSH model on a periodic square lattice (register-tile kernel)");
    const int T = 256;
    const long long n = (long long)h->L * h->Nb;
    dim3 grid((unsigned)std::min<long long>((n + T - 1) / T, 8LL * h->sm_count), (unsigned)nrep);
    ssh_update_kernel<<<grid, T, 0, h->stream>>>(x_dev, h->d_t, h->d_alpha, h->d_alpha2, h->d_col_ph, h->d_col_bond, nullptr, nullptr,
                                                 h->ssq.d_slot, tab_dev, h->Nb, h->Nph, n, h->dtau, x_stride, tab_stride);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_launch_transpose(elph_handle* h, const double* in, double* out, int rows, int cols, int64_t nbatch) {
    dim3 block(32, 8);
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, (unsigned)nbatch);
    transpose_kernel<double><<<grid, block, 0, h->stream>>>(in, out, rows, cols);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_launch_transpose_c(elph_handle* h, const cplx* in, cplx* out, int rows, int cols) {
    dim3 block(32, 8);
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, 1);
    transpose_kernel<cplx><<<grid, block, 0, h->stream>>>(in, out, rows, cols);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}
