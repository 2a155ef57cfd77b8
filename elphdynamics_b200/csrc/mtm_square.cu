// Register-resident fused M^T M kernel for periodic square lattices (Holstein, uniform hopping per colour).
//
// Same operator as matvec.cu (reference: src/HolsteinModels.jl:569-684, src/Checkerboard.jl:57-175,
// src/Models.jl:215-224); this is the roofline path for configs B (32x32xL200) and E (64x64xL400).
//
// Mapping.  site = x + Lx*y.  A warp owns PY consecutive rows of one tau-slice; lane = x mod 32, and for
// Lx = 32*NSEG each lane carries NSEG x-segments, so the warp's tile (PY x Lx sites) lives in registers
// (R = PY*NSEG doubles per array).  The four colour groups of the square lattice
//     g0: (x,x+1) x even     g1: (x,x+1) x odd (incl. the periodic wrap)
//     g2: (y,y+1) y even     g3: (y,y+1) y odd (incl. the periodic wrap)
// become:  g0 = one butterfly shuffle (lane^1), g1 = one rotate shuffle (lane+-1, wrap lanes forward the
// neighbouring x-segment), g2 = register-to-register, g3 = register-to-register except the two tile-edge rows,
// which are exchanged with the neighbouring warps of the CTA through a double-buffered shared-memory strip
// (one __syncthreads per sweep).  No shared-memory traffic for the sweeps themselves.
//
// Streaming.  A CTA (all warps of a slice) walks a chunk of C+1 consecutive slices:
//     t = D(tau) .* v(tau-1);  t = K t;  w(tau) = v(tau) -/+ t;  u = K^T w(tau);
//     y(tau-1) = w(tau-1) -/+ D(tau) .* u
// keeping v(tau-1), w(tau-1) in registers, so each slice of v and D is read once per chunk (+1 halo slice)
// and y written once: 24 B/pt compulsory traffic.  The v/D tiles of the next slices are prefetched by TMA
// bulk copies (cp.async.bulk + mbarrier, one private pipeline per warp: a tile is contiguous in the
// [tau][y][x] layout), so HBM latency overlaps the sweeps.
//
// CG fusion (FUSEP): v := pr + beta*pold is formed on the fly and written to pnew; p.Ap = |M p|^2 is
// accumulated as sum w^2 (identical in exact arithmetic to dot(p, M^T M p)); the last CTA folds the per-CTA
// partials in index order and publishes alpha (see cg.cu).
#include "bulk_copy.cuh"
#include "elph_internal.cuh"
#include "ll_words.cuh"
#include "square_tiles.cuh"

namespace {

constexpr int kStages = 2;

struct SqParams {
    const double* __restrict__ v;
    double* __restrict__ y;
    const double* __restrict__ D;
    const double* __restrict__ pr;
    const double* __restrict__ pold;
    double* __restrict__ pnew;
    double* __restrict__ partial;
    CgScalars* S;
    unsigned int* ticket;
    long long v_stride, y_stride, D_stride;
    int L, Ly, C;
    int open, tau0, Lglob;   // tau-sharded slab: halo slices at index -1 / L instead of the periodic wrap
    double c0, s0, c1, s1, c2, s2, c3, s3;
    double t0, t1, t2, t3, cprod;   // tanh form: t_g = s_g / c_g, cprod = c0 c1 c2 c3
    // HALO: the halo exchange of the sharded product inside this kernel (see HaloArgs in elph_internal.cuh)
    unsigned long long* hx_mine;
    unsigned long long* hx_left;
    unsigned long long* hx_right;
    unsigned int* hx_fail;
    double* v_halo_out;      // the vector again, writable: the received slices are also stored in its halo rows
    unsigned int hx_tag;
};

using namespace tma;   // mbarrier + cp.async.bulk helpers (bulk_copy.cuh)

// register tiles and colour groups: square_tiles.cuh (shared with the CG, KPM and SSH kernels)
using namespace sqt;

template <int NSEG, int PY>
__device__ __forceinline__ void exchange_edges(const Tile<NSEG, PY>& t, double* strip, int warp, int nwarps, int lane,
                                               double (&above)[NSEG], double (&below)[NSEG]) {
    exchange_edges1(t, strip, warp, nwarps, lane, above, below);
}

// HC: honeycomb lattice 32 cells wide (NSEG = 2 = the orbitals of a cell; element (r, q) of lane l sits at r * 64 + 2 l + q of
// the tile), three colours instead of four (hc tiles of square_tiles.cuh)
// HALO (tau-sharded slab): the one-slice halo exchange of the product happens inside the kernel.  CTA 0 pushes
// the first and the last own slice of v into the neighbour GPUs' arenas as self-validating words (ll_words.cuh) before anything
// else; only the CTA of the first chunk (left halo, v[-1]) and of the last chunk (right halo, v[L]) wait for their slice, every
// thread polling the elements of its own tile, while all other chunks stream as usual -- the NVLink round trip hides behind the
// interior of the slab, and a sharded product is ONE launch instead of exchange kernel + product kernel.
// TANH (square lattices): the sweeps in tanh form (square_tiles.cuh) -- K = (c0 c1 c2 c3) prod_g (1 + t_g X_g), the constant folded
// into D: 12 fp64 operations per lattice point and product instead of 20.  The kernel is HBM bound at full clocks but loses ~13 %
// when the SM clock drops under the power cap, i.e. it is then limited by instruction issue; the result differs by rounding only.
template <int NSEG, int PY, bool FUSEP, int MAXT, bool HC = false, bool HALO = false, bool TANH = false>
__global__ void __launch_bounds__(MAXT) mtm_square_kernel(SqParams P) {
    constexpr int LX = 32 * NSEG;
    static_assert(!HC || NSEG == 2, "honeycomb tiles hold the two orbitals of a cell");
    auto eoff = [&](int r, int q, int lane_) -> int { return HC ? r * LX + 2 * lane_ + q : r * LX + 32 * q + lane_; };
    constexpr int TILE = PY * LX;                      // doubles per tile
    constexpr int NT = FUSEP ? 3 : 2;                  // tiles per stage: v (or pr, pold) and D
    constexpr uint32_t STAGE_BYTES = NT * TILE * sizeof(double);
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
    const double* __restrict__ D = P.D + (size_t)blockIdx.y * P.D_stride;
    double* __restrict__ y = P.y + (size_t)blockIdx.y * P.y_stride;

    double* stage_base = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kStages * NT * TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwarps * kStages * STAGE_BYTES) + warp * kStages;
    double* strips = reinterpret_cast<double*>(smem_raw + (size_t)nwarps * kStages * STAGE_BYTES + (size_t)nwarps * kStages * 8);
    const size_t tile_off = (size_t)warp * TILE;  // offset of this warp's rows inside a slice

    const size_t hx_side = (size_t)N * 2, hx_par = (size_t)(P.hx_tag & 1u) * 2 * hx_side;
    if (HALO) {
        // the first CTAs of the grid (all in the first wave) share the push of the two boundary slices: one element per thread
        // and slice, so the loads and the posted stores of a slice are all in flight at once
        const int npush = min((int)gridDim.x, (N + (int)blockDim.x - 1) / (int)blockDim.x);
        if ((int)blockIdx.x < npush) {
            for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N; e += npush * blockDim.x) {
                const double first = vin[e], last = vin[(size_t)(L - 1) * N + e];
                ll::push(P.hx_left + hx_par + hx_side + 2 * (size_t)e, first, P.hx_tag);     // first slice -> left GPU's hi row
                ll::push(P.hx_right + hx_par + 2 * (size_t)e, last, P.hx_tag);               // last slice -> right GPU's lo row
            }
        }
    }
    // wait for this thread's tile of the lo (side 0) / hi (side 1) halo row of this exchange: all elements are polled together
    auto hx_wait_tile = [&](int side, Tile<NSEG, PY>& out) {
        unsigned long long w0[PY][NSEG], w1[PY][NSEG];
        const unsigned long long* row = P.hx_mine + hx_par + (size_t)side * hx_side + 2 * tile_off;
        unsigned int spins = 0;
        bool ok;
        do {
            ok = true;
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) ll::ld2(row + 2 * (size_t)eoff(r, q, lane), w0[r][q], w1[r][q]);
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) ok = ok && ll::tag_ok(w0[r][q], w1[r][q], P.hx_tag);
        } while (!ok && ++spins < (1u << 24));
        if (!ok) *reinterpret_cast<volatile unsigned int*>(P.hx_fail) = 1u;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) out.a[r][q] = ll::unpack(w0[r][q], w1[r][q]);
    };

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kStages; ++k) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](int j) {
        if (lane == 0) {
            int tau = a + j;                       // memory slice index; in an open slab index L is the right halo
            if (!P.open && tau >= L) tau -= L;
            const int st = j % kStages;
            double* dst = stage_base + (size_t)st * NT * TILE;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bars[st], STAGE_BYTES);
            const size_t g = (size_t)tau * N + tile_off;
            bulk_g2s(dst, vin + g, TILE * sizeof(double), &bars[st]);
            if (FUSEP) bulk_g2s(dst + TILE, P.pold + g, TILE * sizeof(double), &bars[st]);
            bulk_g2s(dst + (NT - 1) * TILE, D + g, TILE * sizeof(double), &bars[st]);
        }
    };
#pragma unroll
    for (int j = 0; j < kStages; ++j)
        if (j < nsteps) issue(j);

    Tile<NSEG, PY> vprev, wprev, t, u, dsc;
    {   // v(a-1), straight from global (coalesced, once per chunk)
        const int taum = (a == 0) ? (P.open ? -1 : L - 1) : a - 1;   // open slab: index -1 is the left halo slice
        const long long g = (long long)taum * N + (long long)tile_off;
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const long long e = g + eoff(r, q, lane);
                if (!(HALO && a == 0)) vprev.a[r][q] = FUSEP ? fma(beta, P.pold[e], P.pr[e]) : vin[e];
                wprev.a[r][q] = 0.0;
            }
    }

    if (HALO && a == 0) {     // left halo slice v[-1]: from the arena (also stored into the halo row of v for later readers)
        hx_wait_tile(0, vprev);
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) P.v_halo_out[-(long long)N + (long long)tile_off + eoff(r, q, lane)] = vprev.a[r][q];
    }

    double acc = 0.0;
    int xbuf = 0;
    double above[NSEG], below[NSEG];
    for (int j = 0; j < nsteps; ++j) {
        int tau = a + j;
        if (!P.open && tau >= L) tau -= L;
        // antiperiodic boundary: the '+' sign belongs to GLOBAL time slice 0 (tau0 = global index of local slice 0)
        const bool wrap = ((P.tau0 + a + j) % P.Lglob) == 0;
        const int st = j % kStages;
        const double* sv = stage_base + (size_t)st * NT * TILE;
        const double* sD = sv + (NT - 1) * TILE;
        mbar_wait(&bars[st], (uint32_t)((j / kStages) & 1));
        // t = D(tau) .* v(tau-1)
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                if constexpr (TANH) {
                    dsc.a[r][q] = P.cprod * sD[eoff(r, q, lane)];
                    t.a[r][q] = dsc.a[r][q] * vprev.a[r][q];
                } else {
                    t.a[r][q] = sD[eoff(r, q, lane)] * vprev.a[r][q];
                }
            }
        // t = K t : the colour groups in order
        if constexpr (TANH) {
            g0_x_even_t(t, P.t0);
            g1_x_odd_t(t, P.t1, lane);
            g2_y_even_t(t, P.t2);
            exchange_edges(t, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
            xbuf ^= 1;
            g3_y_odd_t(t, P.t3, above, below);
        } else if constexpr (HC) {
            hc0_cell(t, P.c0, P.s0);
            hc1_lane(t, P.c1, P.s1, lane);
            double aB, bA;
            exchange_hc1(t, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, aB, bA);
            xbuf ^= 1;
            hc2_row(t, P.c2, P.s2, aB, bA);
        } else {
            g0_x_even(t, P.c0, P.s0);
            g1_x_odd(t, P.c1, P.s1, lane);
            g2_y_even(t, P.c2, P.s2);
            exchange_edges(t, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
            xbuf ^= 1;
            g3_y_odd(t, P.c3, P.s3, above, below);
        }
        // w(tau) = v(tau) -/+ t
        Tile<NSEG, PY> hright;
        if (HALO && a + j == L) hx_wait_tile(1, hright);
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const int e = eoff(r, q, lane);
                double vc = FUSEP ? fma(beta, sv[TILE + e], sv[e]) : sv[e];
                if (HALO && a + j == L) {          // the right halo slice: from the arena, not from the (stale) halo row of v
                    vc = hright.a[r][q];
                    P.v_halo_out[(size_t)L * N + tile_off + e] = vc;
                }
                if (FUSEP && j < nout) P.pnew[(size_t)tau * N + tile_off + e] = vc;
                const double w = wrap ? (vc + t.a[r][q]) : (vc - t.a[r][q]);
                t.a[r][q] = w;
                vprev.a[r][q] = vc;
                if (FUSEP && j < nout) acc = fma(w, w, acc);
            }
        if (j >= 1) {
            // u = K^T w(tau) : g3, g2, g1, g0
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) u.a[r][q] = t.a[r][q];
            if constexpr (TANH) {
                exchange_edges(u, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
                xbuf ^= 1;
                g3_y_odd_t(u, P.t3, above, below);
                g2_y_even_t(u, P.t2);
                g1_x_odd_t(u, P.t1, lane);
                g0_x_even_t(u, P.t0);
            } else if constexpr (HC) {
                double aB, bA;
                exchange_hc1(u, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, aB, bA);
                xbuf ^= 1;
                hc2_row(u, P.c2, P.s2, aB, bA);
                hc1_lane(u, P.c1, P.s1, lane);
                hc0_cell(u, P.c0, P.s0);
            } else {
                exchange_edges(u, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
                xbuf ^= 1;
                g3_y_odd(u, P.c3, P.s3, above, below);
                g2_y_even(u, P.c2, P.s2);
                g1_x_odd(u, P.c1, P.s1, lane);
                g0_x_even(u, P.c0, P.s0);
            }
            // y(tau-1) = w(tau-1) -/+ D(tau) .* u     ('+' on the antiperiodic wrap tau = 0)
            const int taum = a + j - 1;   // always one of the CTA's own output slices
            const size_t g = (size_t)taum * N + tile_off;
#pragma unroll
            for (int r = 0; r < PY; ++r)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const int e = eoff(r, q, lane);
                    if constexpr (TANH) {
                        y[g + e] = wrap ? fma(dsc.a[r][q], u.a[r][q], wprev.a[r][q]) : fma(-dsc.a[r][q], u.a[r][q], wprev.a[r][q]);
                    } else {
                        const double du = sD[e] * u.a[r][q];
                        y[g + e] = wrap ? (wprev.a[r][q] + du) : (wprev.a[r][q] - du);
                    }
                }
        }
#pragma unroll
        for (int r = 0; r < PY; ++r)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) wprev.a[r][q] = t.a[r][q];
        __syncwarp();  // every lane is done reading this stage
        if (j + kStages < nsteps) issue(j + kStages);
    }

    if (FUSEP) {
        // per-CTA partial of p.Ap = |M p|^2 over the CTA's own slices, then last-CTA fold (index order)
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

template <int NSEG, int PY, bool FUSEP, int MAXT, bool HC = false, bool HALO = false, bool TANH = false>
void launch_sq(elph_handle* h, const SqParams& P, dim3 grid, int nwarps) {
    constexpr int LX = 32 * NSEG;
    constexpr int NT = FUSEP ? 3 : 2;
    const size_t smem = (size_t)nwarps * kStages * NT * PY * LX * sizeof(double) + (size_t)nwarps * kStages * 8 +
                        2ull * nwarps * 2 * LX * sizeof(double);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "square kernel: tile pipeline does not fit in shared memory");
    elph_enable_smem(h, mtm_square_kernel<NSEG, PY, FUSEP, MAXT, HC, HALO, TANH>);
    mtm_square_kernel<NSEG, PY, FUSEP, MAXT, HC, HALO, TANH><<<grid, nwarps * 32, smem, h->stream>>>(P);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace

// Recognise the periodic square lattice with the reference's colouring: 4 groups = (x-bonds, x even), (x-bonds, x odd
// incl. the wrap), (y-bonds, y even), (y-bonds, y odd incl. the wrap).  slot[b] = dir*N + origin site of column b
// (origin = the site the bond leaves in +x / +y direction).
bool elph_match_square(const elph_handle* h, int* Lx_out, int* Ly_out, std::vector<int>* slot) {
    if (h->ngroups != 4) return false;
    const int N = h->N;
    if (h->Nb != 2 * N) return false;
    for (int Lx = 32; Lx <= 128; Lx += 32) {
        if (N % Lx) continue;
        const int Ly = N / Lx;
        if (Ly < 4 || (Ly & 1)) continue;
        bool ok = true;
        if (slot) slot->assign(h->Nb, -1);
        for (int g = 0; g < 4 && ok; ++g) {
            const int lo = h->goff_host[g], hi = h->goff_host[g + 1];
            if (hi - lo != N / 2) { ok = false; break; }
            std::vector<char> seen(N, 0);
            for (int b = lo; b < hi && ok; ++b) {
                int i = h->bonds_host[b].x, j = h->bonds_host[b].y;
                const int xi = i % Lx, yi = i / Lx, xj = j % Lx, yj = j / Lx;
                bool match = false;
                if (g < 2) {   // x-bond: same row, x' = x+1 mod Lx with x parity = g
                    if (yi == yj) {
                        if ((xi + 1) % Lx == xj && (xi & 1) == g) match = true;
                        else if ((xj + 1) % Lx == xi && (xj & 1) == g) { match = true; std::swap(i, j); }
                    }
                } else {       // y-bond: same column, y' = y+1 mod Ly with y parity = g-2
                    if (xi == xj) {
                        if ((yi + 1) % Ly == yj && (yi & 1) == g - 2) match = true;
                        else if ((yj + 1) % Ly == yi && (yj & 1) == g - 2) { match = true; std::swap(i, j); }
                    }
                }
                if (!match || seen[i]) ok = false;
                seen[i] = 1;
                if (ok && slot) (*slot)[b] = (g / 2) * N + i;
            }
        }
        if (!ok) continue;
        *Lx_out = Lx;
        *Ly_out = Ly;
        return true;
    }
    return false;
}

// Holstein fast path: additionally the (cosh, sinh) pair must be uniform per colour.
void elph_detect_square(elph_handle* h, const std::vector<double2>& cs) {
    h->sq.enabled = false;
    if (h->model != ELPH_MODEL_HOLSTEIN) return;
    int Lx, Ly;
    if (!elph_match_square(h, &Lx, &Ly, nullptr)) return;
    for (int g = 0; g < 4; ++g)
        for (int b = h->goff_host[g]; b < h->goff_host[g + 1]; ++b)
            if (cs[b].x != cs[h->goff_host[g]].x || cs[b].y != cs[h->goff_host[g]].y) return;
    h->sq.enabled = true;
    h->sq.Lx = Lx;
    h->sq.Ly = Ly;
    for (int g = 0; g < 4; ++g) {
        h->sq.c[g] = cs[h->goff_host[g]].x;
        h->sq.s[g] = cs[h->goff_host[g]].y;
    }
}

// Honeycomb lattice, 32 unit cells wide, uniform hopping per bond type: the three colour groups must be exactly the three bond
// types in the order cell / lane / row (this is what the reference's greedy colouring of the sorted neighbour table yields).
void elph_detect_honeycomb(elph_handle* h, const std::vector<double2>& cs) {
    h->hc.enabled = false;
    if (h->model != ELPH_MODEL_HOLSTEIN || h->ngroups != 3) return;
    const int N = h->N, L1 = 32;
    if (N % (2 * L1) || 2 * h->Nb != 3 * N) return;
    const int L2 = N / (2 * L1);
    if (L2 < 4 || L2 % 4) return;
    for (int g = 0; g < 3; ++g) {
        const int lo = h->goff_host[g], hi = h->goff_host[g + 1];
        if (hi - lo != N / 2) return;
        std::vector<char> seen(N / 2, 0);
        for (int b = lo; b < hi; ++b) {
            int i = h->bonds_host[b].x, j = h->bonds_host[b].y;
            if (i & 1) std::swap(i, j);           // i = the A site
            if ((i & 1) || !(j & 1)) return;
            const int ci = i / 2, cj = j / 2;
            const int l1 = ci % L1, l2 = ci / L1, m1 = cj % L1, m2 = cj / L1;
            bool ok;
            if (g == 0) ok = (ci == cj);
            else if (g == 1) ok = (m2 == l2 && m1 == (l1 + L1 - 1) % L1);
            else ok = (m1 == l1 && m2 == (l2 + L2 - 1) % L2);
            if (!ok || seen[ci]) return;
            seen[ci] = 1;
            if (cs[b].x != cs[lo].x || cs[b].y != cs[lo].y) return;
        }
    }
    h->hc.enabled = true;
    h->hc.L1 = L1;
    h->hc.L2 = L2;
    for (int g = 0; g < 3; ++g) {
        h->hc.c[g] = cs[h->goff_host[g]].x;
        h->hc.s[g] = cs[h->goff_host[g]].y;
    }
}

bool elph_launch_mtm_square(elph_handle* h, const MatvecArgs& a) {
    const bool hc = h->hc.enabled && h->hc_tiles && !h->sq_disable && !a.open;
    if (!hc && (!h->sq.enabled || h->sq_disable)) return false;
    if (a.partial_dot && !a.cg_S) return false;
    const int Lx = hc ? 64 : h->sq.Lx, Ly = hc ? h->hc.L2 : h->sq.Ly;
    const bool fusep = (a.cg_S != nullptr);
    // tile shape: PY rows per warp
    int PY;
    // latency regime (one L2-resident lattice, fewer slices than the machine can hold): shorter per-warp chains win
    // (measured at 32x32xL200: PY=8, C=2 -> 5.5 us; PY=16, C=1 -> 7.6 us); throughput regime: PY=16, C=4
    const bool latency_regime = (a.nbatch * (int64_t)h->L < 6LL * h->sm_count);
    if (Lx == 32) PY = (Ly % 16 == 0 && h->sq_py != 8 && !(latency_regime && h->sq_py == 0)) ? 16 : 8;
    else if (Lx == 128) PY = 4;
    else PY = 8;
    if (h->sq_py == 4 && Lx == 64) PY = 4;
    if (hc) PY = (Ly % 8 == 0 && !latency_regime) ? 8 : 4;
    if (Ly % PY) return false;
    const int nwarps = Ly / PY;
    if (nwarps > 32 || nwarps < 2) return false;
    SqParams P;
    P.v = a.v; P.y = a.y; P.D = a.D ? a.D : h->d_D;
    P.pr = a.cg_pr; P.pold = a.cg_pold; P.pnew = a.cg_pnew; P.partial = a.partial_dot; P.S = a.cg_S; P.ticket = a.cg_ticket;
    P.v_stride = a.v_stride; P.y_stride = a.y_stride; P.D_stride = a.D_stride;
    P.L = h->L; P.Ly = Ly;
    P.open = a.open ? 1 : 0; P.tau0 = a.open ? h->shard_tau0 : 0; P.Lglob = a.open ? h->shard_Lglob : h->L;
    const bool halo = a.halo.enabled && a.open && !fusep && !hc && a.nbatch == 1;
    if (a.halo.enabled && !halo) return false;     // the caller falls back to exchange kernel + product
    P.hx_mine = a.halo.mine; P.hx_left = a.halo.left; P.hx_right = a.halo.right; P.hx_fail = a.halo.fail; P.hx_tag = a.halo.tag;
    P.v_halo_out = a.halo.v_out;
    int C = h->chunk_override;
    if (C <= 0) {
        // halo overhead is one extra K-sweep and one extra v slice per chunk: favour long chunks once the
        // machine is covered ~4x, otherwise split finer
        // measured on B200 (32x32xL200, 16..128 replicas): C = 4 is the optimum once the grid covers the
        // machine several times (5.7-6.2 TB/s); longer chunks lose more to the tail than they save in halo
        // latency regime: one wave with about one CTA per SM (measured: 32x32xL200 -> C=2: 5.5 us vs 7.4 us at C=1;
        // 64x64xL400 -> C=3: 12.3 us vs 16.4 us at C=1 and 21.5 us at C=8)
        C = 1;
        if (latency_regime) C = (int)std::min<int64_t>(8, (a.nbatch * (int64_t)h->L + h->sm_count - 1) / h->sm_count);
        const int64_t want = 6LL * h->sm_count;
        for (int c : {4, 2}) {
            if (a.nbatch * ((h->L + c - 1) / c) >= want) { C = c; break; }
        }
    }
    if (C > h->L) C = h->L;
    P.C = C;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    P.t0 = P.s0 / P.c0; P.t1 = P.s1 / P.c1; P.t2 = P.s2 / P.c2; P.t3 = P.s3 / P.c3;
    P.cprod = P.c0 * P.c1 * P.c2 * P.c3;
    const bool tanh_form = h->mtm_tanh && !hc;
    if (hc) {
        P.c0 = h->hc.c[0]; P.s0 = h->hc.s[0]; P.c1 = h->hc.c[1]; P.s1 = h->hc.s[1];
        P.c2 = h->hc.c[2]; P.s2 = h->hc.s[2]; P.c3 = 1.0; P.s3 = 0.0;
    }
    const int nchunks = (h->L + C - 1) / C;
    dim3 grid(nchunks, (unsigned)a.nbatch);
    if (fusep) ELPH_REQUIRE(nchunks <= h->partial_cap, ELPH_ERR_INVALID, "partial buffer too small");
    if (a.npartial) *a.npartial = nchunks;
#define SQ_CASE(NS, PYV, MAXT)                                                      \
    if (Lx == 32 * NS && PY == PYV && nwarps * 32 <= MAXT) {                        \
        if (fusep) launch_sq<NS, PYV, true, MAXT>(h, P, grid, nwarps);              \
        else if (halo && tanh_form) launch_sq<NS, PYV, false, MAXT, false, true, true>(h, P, grid, nwarps); \
        else if (halo) launch_sq<NS, PYV, false, MAXT, false, true>(h, P, grid, nwarps); \
        else if (tanh_form) launch_sq<NS, PYV, false, MAXT, false, false, true>(h, P, grid, nwarps); \
        else launch_sq<NS, PYV, false, MAXT>(h, P, grid, nwarps);                   \
        return true;                                                                \
    }
    if (hc) {
        if (PY == 8 && nwarps * 32 <= 256) {
            if (fusep) launch_sq<2, 8, true, 256, true>(h, P, grid, nwarps);
            else launch_sq<2, 8, false, 256, true>(h, P, grid, nwarps);
            return true;
        }
        if (PY == 4 && nwarps * 32 <= 512) {
            if (fusep) launch_sq<2, 4, true, 512, true>(h, P, grid, nwarps);
            else launch_sq<2, 4, false, 512, true>(h, P, grid, nwarps);
            return true;
        }
        return false;
    }
    SQ_CASE(1, 16, 256)
    SQ_CASE(1, 8, 256)
    SQ_CASE(2, 8, 256)
    SQ_CASE(2, 4, 512)
    SQ_CASE(3, 8, 256)
    SQ_CASE(4, 4, 512)
#undef SQ_CASE
    return false;
}
