// Register-resident fused M^T M kernel for the SSH model on periodic square lattices (config C: 32x32xL200).
//
// Operator: src/SSHModels.jl:581-701 (mulM!/mulMT!), src/Checkerboard.jl:86-121,177-210 (per-tau tables),
// src/Models.jl:215-224 (M^T M).  Same tiling as mtm_square.cu -- lane = x, PY rows per warp in registers, the four
// colours as butterfly shuffle / rotate shuffle / register pairs / edge-row exchange -- but every bond carries its
// own (cosh, sinh) per time slice: B(tau) = K(tau) diag(exp(dtau mu)), K(tau) = prod_bonds Gamma(cosh, sinh)(tau, bond).
//
// Tables.  update_model writes a second copy of the (cosh, sinh) table in the tile layout [tau][dir][site]
// (dir 0: the +x bond leaving `site`, dir 1: the +y bond), so that the table tile of a warp is contiguous like its v
// tile.  Per step the warp's pipeline stage receives by TMA bulk copy: v(tau) [PY][LX], the x-table [PY][LX] double2,
// the y-table [PY][LX] double2 plus the row above the tile (the y-odd bonds entering the tile from the neighbour warp).
// Step tau uses K(tau) twice (K t for w(tau), K^T w(tau) for y(tau-1)), both from the same staged tile, so the
// compulsory traffic is v + y + 2 x 16 B of tables = 48 B/pt.
#include "bulk_copy.cuh"
#include "elph_internal.cuh"
#include "square_tiles.cuh"

namespace {

using namespace tma;
using namespace sqt;

struct SshParams {
    const double* __restrict__ v;
    double* __restrict__ y;
    const double* __restrict__ Dmu;      // exp(dtau mu) [N]
    const double2* __restrict__ tab;     // [L][2][N]
    const double* __restrict__ pr;
    const double* __restrict__ pold;
    double* __restrict__ pnew;
    double* __restrict__ partial;
    CgScalars* S;
    unsigned int* ticket;
    long long v_stride, y_stride, tab_stride;   // tab_stride != 0: one table per blockIdx.y (independent replicas)
    int L, Ly, C;
};

template <int NSEG, int PY, bool FUSEP, int STAGES, int MAXT>
__global__ void __launch_bounds__(MAXT) ssh_square_kernel(SshParams P) {
    constexpr int LX = 32 * NSEG;
    constexpr int TILE = PY * LX;                          // sites per tile
    constexpr int NV = FUSEP ? 2 : 1;                      // v (or pr and pold)
    constexpr int STAGE_DBL = NV * TILE + 2 * TILE + 2 * (PY + 1) * LX;
    constexpr uint32_t STAGE_BYTES = STAGE_DBL * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L;
    const int N = LX * P.Ly;
    const int a = blockIdx.x * P.C;
    const int nout = min(P.C, L - a);
    const int nsteps = nout + 1;

    double beta = 0.0;
    if (FUSEP) {
        if (P.S->done) return;
        beta = P.S->beta;
    }
    const double* __restrict__ vin = FUSEP ? P.pr : P.v + (size_t)blockIdx.y * P.v_stride;
    double* __restrict__ y = P.y + (size_t)blockIdx.y * P.y_stride;
    const double2* __restrict__ tab = P.tab + (size_t)blockIdx.y * P.tab_stride;

    double* stage_base = reinterpret_cast<double*>(smem_raw) + (size_t)warp * STAGES * STAGE_DBL;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwarps * STAGES * STAGE_BYTES) + warp * STAGES;
    double* strips = reinterpret_cast<double*>(smem_raw + (size_t)nwarps * STAGES * STAGE_BYTES + (size_t)nwarps * STAGES * 8);
    const int y0 = warp * PY;
    const size_t tile_off = (size_t)y0 * LX;
    const size_t halo_off = (size_t)((y0 + P.Ly - 1) % P.Ly) * LX;

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < STAGES; ++k) mbar_init(&bars[k], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto issue = [&](int j) {
        if (lane == 0) {
            int tau = a + j;
            if (tau >= L) tau -= L;
            const int st = j % STAGES;
            double* dst = stage_base + (size_t)st * STAGE_DBL;
            fence_proxy_async();
            mbar_expect_tx(&bars[st], STAGE_BYTES);
            const size_t g = (size_t)tau * N + tile_off;
            bulk_g2s(dst, vin + g, TILE * sizeof(double), &bars[st]);
            if (FUSEP) bulk_g2s(dst + TILE, P.pold + g, TILE * sizeof(double), &bars[st]);
            const double2* tx = tab + (size_t)tau * 2 * N;
            const double2* ty = tx + N;
            double* dtx = dst + NV * TILE;
            double* dty = dtx + 2 * TILE;
            bulk_g2s(dtx, tx + tile_off, TILE * sizeof(double2), &bars[st]);
            bulk_g2s(dty, ty + halo_off, LX * sizeof(double2), &bars[st]);
            bulk_g2s(dty + 2 * LX, ty + tile_off, TILE * sizeof(double2), &bars[st]);
        }
    };
#pragma unroll
    for (int j = 0; j < STAGES; ++j)
        if (j < nsteps) issue(j);

    Tile<NSEG, PY> vprev, wprev, t, u, dm;
    {   // v(a-1) and exp(dtau mu), straight from global (coalesced, once per chunk)
        const int taum = (a == 0) ? L - 1 : a - 1;
        const size_t g = (size_t)taum * N + tile_off;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = (size_t)r * LX + 32 * q + lane;
                vprev.a[r][q] = FUSEP ? fma(beta, P.pold[g + e], P.pr[g + e]) : vin[g + e];
                wprev.a[r][q] = 0.0;
                dm.a[r][q] = P.Dmu[tile_off + e];
            }
    }

    double acc = 0.0;
    int xbuf = 0;
    double above[NSEG], below[NSEG];
    for (int j = 0; j < nsteps; ++j) {
        int tau = a + j;
        if (tau >= L) tau -= L;
        const bool wrap = (tau == 0);   // antiperiodic boundary
        const int st = j % STAGES;
        const double* sv = stage_base + (size_t)st * STAGE_DBL;
        const double2* tx = reinterpret_cast<const double2*>(sv + NV * TILE);
        const double2* ty_halo = tx + TILE;            // [1][LX] row above the tile, then the tile rows
        const double2* ty = ty_halo + LX;
        mbar_wait(&bars[st], (uint32_t)((j / STAGES) & 1));
        // t = exp(dtau mu) .* v(tau-1);  t = K(tau) t
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) t.a[r][q] = dm.a[r][q] * vprev.a[r][q];
        g0_tab(t, tx, lane);
        g1_tab(t, tx, lane);
        g2_tab(t, ty, lane);
        exchange_edges1(t, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
        xbuf ^= 1;
        g3_tab(t, ty, ty_halo, lane, above, below);
        // w(tau) = v(tau) -/+ t
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const int e = r * LX + 32 * q + lane;
                const double vc = FUSEP ? fma(beta, sv[TILE + e], sv[e]) : sv[e];
                if (FUSEP && j < nout) P.pnew[(size_t)tau * N + tile_off + e] = vc;
                const double w = wrap ? (vc + t.a[r][q]) : (vc - t.a[r][q]);
                t.a[r][q] = w;
                vprev.a[r][q] = vc;
                if (FUSEP && j < nout) acc = fma(w, w, acc);
            }
        if (j >= 1) {
            // u = K^T(tau) w(tau): colours in reverse order;  y(tau-1) = w(tau-1) -/+ exp(dtau mu) .* u
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) u.a[r][q] = t.a[r][q];
            exchange_edges1(u, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
            xbuf ^= 1;
            g3_tab(u, ty, ty_halo, lane, above, below);
            g2_tab(u, ty, lane);
            g1_tab(u, tx, lane);
            g0_tab(u, tx, lane);
            const size_t g = (size_t)(a + j - 1) * N + tile_off;   // always one of the CTA's own output slices
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const int e = r * LX + 32 * q + lane;
                    const double du = dm.a[r][q] * u.a[r][q];
                    y[g + e] = wrap ? (wprev.a[r][q] + du) : (wprev.a[r][q] - du);
                }
        }
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) wprev.a[r][q] = t.a[r][q];
        __syncwarp();  // every lane is done reading this stage
        if (j + STAGES < nsteps) issue(j + STAGES);
    }

    if (FUSEP) {
        // per-CTA partial of p.Ap = |M p|^2 over the CTA's own slices, then last-CTA fold in index order (cg.cu)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int k = 0; k < nwarps; ++k) s += red[k];
            P.partial[blockIdx.x] = s;
            __threadfence();
            const unsigned int n = atomicAdd(P.ticket, 1u);
            is_last = (n == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            double s = 0.0;
            for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += ((volatile double*)P.partial)[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            __syncthreads();
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                double pAp = 0.0;
                for (int k = 0; k < nwarps; ++k) pAp += red[k];
                P.S->pAp = pAp;
                P.S->alpha = P.S->rdotz / pAp;
                *P.ticket = 0u;
            }
        }
    }
}

template <int NSEG, int PY, bool FUSEP>
size_t smem_need(int nwarps, int stages) {
    constexpr int LX = 32 * NSEG;
    constexpr int NV = FUSEP ? 2 : 1;
    const size_t stage = (size_t)(NV * PY * LX + 2 * PY * LX + 2 * (PY + 1) * LX) * sizeof(double);
    return (size_t)nwarps * stages * stage + (size_t)nwarps * stages * 8 + 2ull * nwarps * 2 * LX * sizeof(double);
}

template <int NSEG, int PY, bool FUSEP, int MAXT>
bool launch_ssh(elph_handle* h, const SshParams& P, dim3 grid, int nwarps) {
    if (smem_need<NSEG, PY, FUSEP>(nwarps, 2) <= h->smem_optin) {
        auto k = ssh_square_kernel<NSEG, PY, FUSEP, 2, MAXT>;
        elph_enable_smem(h, k);
        k<<<grid, nwarps * 32, smem_need<NSEG, PY, FUSEP>(nwarps, 2), h->stream>>>(P);
    } else if (smem_need<NSEG, PY, FUSEP>(nwarps, 1) <= h->smem_optin) {
        auto k = ssh_square_kernel<NSEG, PY, FUSEP, 1, MAXT>;   // large slices: single-stage (no intra-CTA prefetch)
        elph_enable_smem(h, k);
        k<<<grid, nwarps * 32, smem_need<NSEG, PY, FUSEP>(nwarps, 1), h->stream>>>(P);
    } else {
        return false;
    }
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    return true;
}

}  // namespace

void elph_detect_ssh_square(elph_handle* h) {
    h->ssq.enabled = false;
    if (h->model != ELPH_MODEL_SSH) return;
    int Lx, Ly;
    std::vector<int> slot;
    if (!elph_match_square(h, &Lx, &Ly, &slot)) return;
    h->ssq.Lx = Lx;
    h->ssq.Ly = Ly;
    h->ssq.d_slot = elph_dalloc<int>(slot.size());
    ELPH_CUDA(cudaMemcpy(h->ssq.d_slot, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->ssq.d_tab = elph_dalloc<double2>((size_t)h->L * h->Nb);
    h->ssq.enabled = true;
}

bool elph_launch_ssh_square(elph_handle* h, const MatvecArgs& a) {
    if (!h->ssq.enabled || h->sq_disable || a.open || a.D) return false;
    if (a.partial_dot && !a.cg_S) return false;
    const int Lx = h->ssq.Lx, Ly = h->ssq.Ly;
    const bool fusep = (a.cg_S != nullptr);
    const int PY = (Lx == 32 && h->sq_py != 4) ? 8 : 4;
    if (Ly % PY) return false;
    const int nwarps = Ly / PY;
    if (nwarps > 32 || nwarps < 2) return false;
    SshParams P;
    P.v = a.v; P.y = a.y; P.Dmu = h->d_D; P.tab = a.ssh_tab ? a.ssh_tab : h->ssq.d_tab;
    P.tab_stride = a.ssh_tab ? a.ssh_tab_stride : 0;
    P.pr = a.cg_pr; P.pold = a.cg_pold; P.pnew = a.cg_pnew; P.partial = a.partial_dot; P.S = a.cg_S; P.ticket = a.cg_ticket;
    P.v_stride = a.v_stride; P.y_stride = a.y_stride;
    P.L = h->L; P.Ly = Ly;
    int C = h->chunk_override;
    if (C <= 0) {
        // one wave with about one CTA per SM for a single lattice; longer chunks once the grid covers the machine
        C = (int)std::min<int64_t>(8, std::max<int64_t>(1, (a.nbatch * (int64_t)h->L + h->sm_count - 1) / h->sm_count));
        if (a.nbatch * (int64_t)h->L >= 6LL * h->sm_count) C = 4;
    }
    if (C > h->L) C = h->L;
    P.C = C;
    const int nchunks = (h->L + C - 1) / C;
    dim3 grid(nchunks, (unsigned)a.nbatch);
    if (fusep && nchunks > h->partial_cap) return false;
    if (a.npartial) *a.npartial = nchunks;
#define SSH_CASE(NS, PYV, MAXT)                                                        \
    if (Lx == 32 * NS && PY == PYV && nwarps * 32 <= MAXT)                             \
        return fusep ? launch_ssh<NS, PYV, true, MAXT>(h, P, grid, nwarps)             \
                     : launch_ssh<NS, PYV, false, MAXT>(h, P, grid, nwarps);
    SSH_CASE(1, 8, 256)
    SSH_CASE(1, 4, 512)
    SSH_CASE(2, 4, 512)
    SSH_CASE(3, 4, 512)
    SSH_CASE(4, 4, 512)
#undef SSH_CASE
    return false;
}
