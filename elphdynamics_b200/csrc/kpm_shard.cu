// KPM preconditioner of a tau-sharded lattice with the transposes THROUGH PEER MEMORY (no NCCL call per application).
//
// Reference: ldiv!(v', P::SymmetricKPMPreconditioner, v), src/KPMPreconditioners.jl:426-481 -- tau_to_omega! (src/TimeFreqFFTs.jl:
// 55-73), the per-frequency polynomials (:606-679), the mirror frequencies (:464-466), omega_to_tau! (:112-130).  On one process
// these are three loops over one array; on P GPUs they live on three shardings of the lattice:
//     tau-sharded r [Lloc][N]  ->  site-sharded [L][N/P]  ->  omega-sharded [L/2P][N]  ->  site-sharded  ->  tau-sharded z.
// Every rank keeps one arena (cudaMalloc + CUDA IPC, mapped by all ranks) with the OUTPUT of each of its stages; the kernel of
// the next stage PULLS what it needs straight out of the producers' arenas over NVLink, inside its own load phase:
//     A  r slab           (copied in)          read by  the forward FFT of every rank (its site block, all slices)
//     B  nu [L/2][N/P]    forward FFT          read by  the gather of every rank (its frequencies, all site blocks)
//     C  nu' [L][N]       Chebyshev chains     read by  the inverse FFT of every rank (its site block, all frequencies)
//     D  z [L][N/P]       inverse FFT          read by  the final gather of every rank (its slices, all site blocks)
// Between two stages one cross-GPU barrier: a 32-thread kernel in which lane q stores the sequence number into rank q's flag
// word and polls its own word from rank q.  Data written by a finished kernel sits in the owner's L2, which is where NVLink
// reads are served from; the readers bypass their L1 (ld.cv).  Write-after-read hazards are excluded by the barrier order (a
// buffer is overwritten one full application after its readers signalled the barrier that follows their reads).
// Timeouts (2 s on %globaltimer) raise a pinned failure flag instead of hanging.
#include "elph_internal.cuh"
#include "fft_smem.cuh"

#include <cstring>

namespace {

using namespace fftsm;
constexpr int kT = 256;
constexpr int kKsMaxWorld = 16;

struct KsPeers {
    char* base[kKsMaxWorld];      // arena of every rank (own included)
    int tau0[kKsMaxWorld + 1];    // first global slice of every rank's slab; tau0[world] = L
    int s0[kKsMaxWorld + 1];      // first site of every rank's site block; s0[world] = N
    int world, me;
    size_t offA, offB, offC, offD;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// cross-GPU barrier number `seq`: lane q tells rank q and waits for rank q
__global__ void ks_sync_kernel(KsPeers P, unsigned long long seq, unsigned int* fail) {
    const int q = threadIdx.x;
    if (q >= P.world) return;
    __threadfence_system();
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(P.base[q]) + P.me;
    *dst = seq;
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(P.base[P.me]) + q;
    const unsigned long long t0 = globaltimer_ns();
    while (*mine < seq) {
        if (globaltimer_ns() - t0 > 2000000000ull) {
            *fail = 1u;
            break;
        }
    }
    __threadfence_system();
}

// stage 1: twisted forward FFT of this rank's site block; the slices come from the slabs (A) of all ranks
template <int SB>
__global__ void __launch_bounds__(kT) ks_fft_fwd_kernel(KsPeers P, FftPlan plan, int N, int Lo2, const cplx* __restrict__ tw_g,
                                                        const cplx* __restrict__ theta) {
    extern __shared__ __align__(16) double smem_raw[];
    const int L = plan.L;
    cplx* b0 = reinterpret_cast<cplx*>(smem_raw);
    cplx* b1 = b0 + (size_t)L * SB;
    cplx* tw = b1 + (size_t)L * SB;
    const int site = threadIdx.x % SB, slot = threadIdx.x / SB, nslots = blockDim.x / SB;
    const int s0 = P.s0[P.me], nloc = P.s0[P.me + 1] - s0;
    const int gsite = blockIdx.x * SB + site;
    const bool ok = gsite < nloc;
    for (int k = threadIdx.x; k < L; k += blockDim.x) tw[k] = tw_g[k];
    int q = 0;
    for (int t = slot; t < L; t += nslots) {
        while (t >= P.tau0[q + 1]) ++q;
        cplx v = make_double2(0.0, 0.0);
        if (ok) {
            const double* A = reinterpret_cast<const double*>(P.base[q] + P.offA);
            const double xr = __ldcv(A + (size_t)(t - P.tau0[q]) * N + s0 + gsite);
            const cplx th = theta[t];
            v = make_double2(th.x * xr, th.y * xr);
        }
        b0[(size_t)t * SB + site] = v;
    }
    __syncthreads();
    cplx* res = fft_smem_auto<SB>(b0, b1, plan, tw, false);
    cplx* B = reinterpret_cast<cplx*>(P.base[P.me] + P.offB);
    for (int t = slot; t < Lo2; t += nslots)
        if (ok) B[(size_t)t * nloc + gsite] = res[(size_t)t * SB + site];
}

// stage 2a: rows of this rank's frequencies (w = me, me + world, ...) for ALL sites, from the B of every site block's owner
__global__ void ks_gather_nu_kernel(KsPeers P, cplx* __restrict__ nu_in, int N) {
    const int w = P.me + blockIdx.y * P.world;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    int p = 0;
    while (e >= P.s0[p + 1]) ++p;
    const int nl = P.s0[p + 1] - P.s0[p];
    const cplx* B = reinterpret_cast<const cplx*>(P.base[p] + P.offB);
    nu_in[(size_t)w * N + e] = __ldcv(B + (size_t)w * nl + (e - P.s0[p]));
}

// stage 3: inverse FFT of this rank's site block; row t of the frequency-space vector comes from the C of the rank that ran
// the chain of frequency t (t in the lower half) or of its mirror L-1-t (the chain kernels write both rows)
template <int SB>
__global__ void __launch_bounds__(kT) ks_fft_inv_kernel(KsPeers P, FftPlan plan, int N, int Lo2, const cplx* __restrict__ tw_g,
                                                        const cplx* __restrict__ theta) {
    extern __shared__ __align__(16) double smem_raw[];
    const int L = plan.L;
    cplx* b0 = reinterpret_cast<cplx*>(smem_raw);
    cplx* b1 = b0 + (size_t)L * SB;
    cplx* tw = b1 + (size_t)L * SB;
    const int site = threadIdx.x % SB, slot = threadIdx.x / SB, nslots = blockDim.x / SB;
    const int s0 = P.s0[P.me], nloc = P.s0[P.me + 1] - s0;
    const int gsite = blockIdx.x * SB + site;
    const bool ok = gsite < nloc;
    for (int k = threadIdx.x; k < L; k += blockDim.x) tw[k] = tw_g[k];
    for (int t = slot; t < L; t += nslots) {
        const int w = (t >= L - Lo2) ? (L - 1 - t) : t;   // for odd L the middle frequency counts as a mirror (:464-466)
        cplx v = make_double2(0.0, 0.0);
        if (ok) {
            const cplx* Cq = reinterpret_cast<const cplx*>(P.base[w % P.world] + P.offC);
            v = __ldcv(Cq + (size_t)t * N + s0 + gsite);
        }
        b0[(size_t)t * SB + site] = v;
    }
    __syncthreads();
    cplx* res = fft_smem_auto<SB>(b0, b1, plan, tw, true);
    double* D = reinterpret_cast<double*>(P.base[P.me] + P.offD);
    const double invL = 1.0 / (double)L;
    for (int t = slot; t < L; t += nslots)
        if (ok) {
            const cplx v = res[(size_t)t * SB + site];
            const cplx th = theta[t];
            D[(size_t)t * nloc + gsite] = (th.x * v.x + th.y * v.y) * invL;
        }
}

// stage 4: this rank's slices of z for ALL sites, from the D of every site block's owner
__global__ void ks_gather_z_kernel(KsPeers P, double* __restrict__ z_own, int N) {
    const int tau = blockIdx.y;                       // local slice
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    int p = 0;
    while (e >= P.s0[p + 1]) ++p;
    const int nl = P.s0[p + 1] - P.s0[p];
    const double* D = reinterpret_cast<const double*>(P.base[p] + P.offD);
    z_own[(size_t)tau * N + e] = __ldcv(D + (size_t)(P.tau0[P.me] + tau) * nl + (e - P.s0[p]));
}

FftPlan ks_plan(const elph_handle* h) {
    FftPlan p;
    p.L = h->L;
    p.nrad = (int)h->fft_radices.size();
    for (int i = 0; i < p.nrad; ++i) p.rad[i] = h->fft_radices[i];
    return p;
}

KsPeers ks_peers(const elph_handle* h) {
    const auto& S = h->kshard;
    KsPeers P;
    for (int q = 0; q < kKsMaxWorld; ++q) P.base[q] = (q < S.world) ? static_cast<char*>(S.peer[q]) : nullptr;
    for (int q = 0; q <= kKsMaxWorld; ++q) {
        P.tau0[q] = S.tau0s[std::min(q, S.world)];
        P.s0[q] = S.s0s[std::min(q, S.world)];
    }
    P.world = S.world;
    P.me = S.rank;
    P.offA = S.offA; P.offB = S.offB; P.offC = S.offC; P.offD = S.offD;
    return P;
}

int ks_pick_sb(const elph_handle* h, int ncols) {
    int sb = 32;
    while (sb > 4 && (ncols + sb - 1) / sb < 2 * h->sm_count) sb >>= 1;
    while (sb > 4 && (2ull * h->L * sb + h->L) * sizeof(cplx) > h->smem_optin) sb >>= 1;
    return sb;
}

template <int SB>
void ks_launch_fft(elph_handle* h, bool inverse, const KsPeers& P, int nloc, int Lo2) {
    const size_t smem = (2ull * h->L * SB + h->L) * sizeof(cplx);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Ltau too large for the shared-memory FFT");
    const int blocks = (nloc + SB - 1) / SB;
    if (inverse) {
        elph_enable_smem(h, ks_fft_inv_kernel<SB>);
        ks_fft_inv_kernel<SB><<<blocks, kT, smem, h->stream>>>(P, ks_plan(h), h->N, Lo2, h->d_twiddle, h->d_theta);
    } else {
        elph_enable_smem(h, ks_fft_fwd_kernel<SB>);
        ks_fft_fwd_kernel<SB><<<blocks, kT, smem, h->stream>>>(P, ks_plan(h), h->N, Lo2, h->d_twiddle, h->d_theta);
    }
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void ks_fft(elph_handle* h, bool inverse, const KsPeers& P, int nloc, int Lo2) {
    switch (ks_pick_sb(h, nloc)) {
        case 32: ks_launch_fft<32>(h, inverse, P, nloc, Lo2); break;
        case 16: ks_launch_fft<16>(h, inverse, P, nloc, Lo2); break;
        case 8: ks_launch_fft<8>(h, inverse, P, nloc, Lo2); break;
        default: ks_launch_fft<4>(h, inverse, P, nloc, Lo2); break;
    }
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

// Allocate this rank's arena for the slab [tau0, tau0 + lloc) of the aux handle's (global) lattice and export its IPC handle.
void elph_kpm_shard_export_impl(elph_handle* h, int rank, int world, int tau0, int lloc, unsigned char* handle_out) {
    ELPH_REQUIRE(h->kpm.configured, ELPH_ERR_STATE, "KPM preconditioner not configured (kpm_n == 0 at elph_create)");
    ELPH_REQUIRE(world >= 1 && world <= kKsMaxWorld && rank >= 0 && rank < world, ELPH_ERR_INVALID, "bad rank / world");
    ELPH_REQUIRE(lloc >= 0 && tau0 >= 0 && tau0 + lloc <= h->L, ELPH_ERR_INVALID, "bad slab bounds");
    auto& S = h->kshard;
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    if (S.arena) elph_kpm_shard_close_impl(h);
    S.rank = rank; S.world = world; S.tau0 = tau0; S.lloc = lloc;
    const int N = h->N, L = h->L, Lo2 = (L + 1) / 2;
    const int lmax = (L + world - 1) / world, nmax = (N + world - 1) / world;
    S.offA = 256;
    S.offB = S.offA + align256((size_t)lmax * N * sizeof(double));
    S.offC = S.offB + align256((size_t)Lo2 * nmax * sizeof(cplx));
    S.offD = S.offC + align256((size_t)L * N * sizeof(cplx));
    S.bytes = S.offD + align256((size_t)L * nmax * sizeof(double));
    ELPH_CUDA(cudaMalloc(&S.arena, S.bytes));
    ELPH_CUDA(cudaMemset(S.arena, 0, S.bytes));
    if (!S.nu_in) S.nu_in = elph_dalloc<cplx>((size_t)L * N);
    ELPH_CUDA(cudaMemset(S.nu_in, 0, (size_t)L * N * sizeof(cplx)));
    if (!S.h_fail) {
        ELPH_CUDA(cudaHostAlloc(&S.h_fail, sizeof(unsigned int), cudaHostAllocMapped));
        ELPH_CUDA(cudaHostGetDevicePointer(&S.d_fail, S.h_fail, 0));
    }
    *S.h_fail = 0u;
    S.seq = 0;
    ELPH_CUDA(cudaDeviceSynchronize());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t ipc;
    ELPH_CUDA(cudaIpcGetMemHandle(&ipc, S.arena));
    memcpy(handle_out, &ipc, sizeof(ipc));
}

// Open the arenas of all ranks (handles: world x 64 bytes in rank order; tau0s: first slice of every slab).
void elph_kpm_shard_open_impl(elph_handle* h, const unsigned char* handles, const int64_t* tau0s) {
    auto& S = h->kshard;
    ELPH_REQUIRE(S.arena, ELPH_ERR_STATE, "elph_kpm_shard_export must be called first");
    ELPH_REQUIRE(handles && tau0s, ELPH_ERR_INVALID, "null argument");
    S.tau0s.assign(S.world + 1, h->L);
    for (int q = 0; q < S.world; ++q) S.tau0s[q] = (int)tau0s[q];
    ELPH_REQUIRE(S.tau0s[0] == 0 && S.tau0s[S.rank] == S.tau0 && S.tau0s[S.rank + 1] == S.tau0 + S.lloc, ELPH_ERR_INVALID,
                 "slab starts do not match this rank's slab");
    for (int q = 0; q < S.world; ++q)
        ELPH_REQUIRE(S.tau0s[q] <= S.tau0s[q + 1] && S.tau0s[q + 1] - S.tau0s[q] <= (h->L + S.world - 1) / S.world, ELPH_ERR_INVALID,
                     "slabs must be contiguous and near-equal");
    S.s0s.assign(S.world + 1, h->N);       // near-equal contiguous site blocks, as sharded.py: slab_bounds(N, world, q)
    const int base = h->N / S.world, rem = h->N % S.world;
    for (int q = 0; q < S.world; ++q) S.s0s[q] = q * base + std::min(q, rem);
    S.peer.assign(S.world, nullptr);
    for (int q = 0; q < S.world; ++q) {
        if (q == S.rank) {
            S.peer[q] = S.arena;
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + (size_t)q * sizeof(ipc), sizeof(ipc));
        ELPH_CUDA(cudaIpcOpenMemHandle(&S.peer[q], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    S.opened = true;
}

void elph_kpm_shard_close_impl(elph_handle* h) {
    auto& S = h->kshard;
    for (int q = 0; q < (int)S.peer.size(); ++q)
        if (S.peer[q] && q != S.rank) cudaIpcCloseMemHandle(S.peer[q]);
    S.peer.clear();
    if (S.arena) cudaFree(S.arena);
    S.arena = nullptr;
    S.opened = false;
}

void elph_kpm_shard_free(elph_handle* h) {
    auto& S = h->kshard;
    elph_kpm_shard_close_impl(h);
    cudaFree(S.nu_in);
    S.nu_in = nullptr;
    if (S.h_fail) cudaFreeHost(S.h_fail);
    S.h_fail = nullptr;
}

// z_own = P^-1 r_own on the slab owned by this rank ([lloc][N] each, device); every rank makes the same call.
void elph_kpm_shard_apply_impl(elph_handle* h, const double* r_own, double* z_own) {
    auto& S = h->kshard;
    KpmState& K = h->kpm;
    ELPH_REQUIRE(S.opened, ELPH_ERR_STATE, "elph_kpm_shard_open has not been called");
    ELPH_REQUIRE(K.ever_setup && K.active && K.d_coeff, ELPH_ERR_STATE, "the fused sharded apply needs an active preconditioner");
    ELPH_REQUIRE(K.sub_first == S.rank && K.sub_stride == S.world, ELPH_ERR_STATE,
                 "elph_kpm_set_omega_subset(rank, world) must select this rank's frequencies");
    ELPH_REQUIRE(*S.h_fail == 0u, ELPH_ERR_STATE, "sharded KPM apply: a peer GPU did not reach a barrier (timeout); re-open the arenas");
    ELPH_REQUIRE(r_own && z_own, ELPH_ERR_INVALID, "null device pointer");
    const int N = h->N, L = h->L, Lo2 = K.Lo2;
    const KsPeers P = ks_peers(h);
    const int nloc = S.s0s[S.rank + 1] - S.s0s[S.rank];
    cudaStream_t st = h->stream;
    auto barrier = [&]() {
        ks_sync_kernel<<<1, 32, 0, st>>>(P, ++S.seq, S.d_fail);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
    };
    if (S.lloc > 0)
        ELPH_CUDA(cudaMemcpyAsync(static_cast<char*>(S.arena) + S.offA, r_own, (size_t)S.lloc * N * sizeof(double),
                                  cudaMemcpyDeviceToDevice, st));
    barrier();
    if (nloc > 0) ks_fft(h, false, P, nloc, Lo2);
    barrier();
    if (K.nsched > 0) {
        ks_gather_nu_kernel<<<dim3((N + kT - 1) / kT, K.nsched), kT, 0, st>>>(P, S.nu_in, N);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        elph_kpm_chains_dev(h, S.nu_in, reinterpret_cast<cplx*>(static_cast<char*>(S.arena) + S.offC));
    }
    barrier();
    if (nloc > 0) ks_fft(h, true, P, nloc, Lo2);
    barrier();
    if (S.lloc > 0) {
        ks_gather_z_kernel<<<dim3((N + kT - 1) / kT, S.lloc), kT, 0, st>>>(P, z_own, N);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
    }
}

// Synchronise and report whether every barrier of the applications so far completed.
bool elph_kpm_shard_ok(elph_handle* h) {
    auto& S = h->kshard;
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    return S.h_fail == nullptr || *S.h_fail == 0u;
}
