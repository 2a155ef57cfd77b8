// Register tiles for periodic square lattices: the four checkerboard colours as warp shuffles and
// register-to-register rotations (see the header comment of mtm_square.cu for the mapping).
// Shared by kpm_square.cu (complex tiles = a pair of real tiles).
#pragma once

#include "elph_internal.cuh"

namespace sqt {

template <int NSEG, int PY>
struct Tile {
    double a[PY][NSEG];
};

template <int NSEG, int PY>
__device__ __forceinline__ void g0_x_even(Tile<NSEG, PY>& t, double c, double s) {
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double o = __shfl_xor_sync(0xffffffffu, t.a[r][q], 1);
            t.a[r][q] = c * t.a[r][q] + s * o;
        }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g1_x_odd(Tile<NSEG, PY>& t, double c, double s, int lane) {
    const int partner = (lane & 1) ? ((lane + 1) & 31) : ((lane + 31) & 31);
#pragma unroll
    for (int r = 0; r < PY; ++r) {
        double o[NSEG];
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            double send = t.a[r][q];
            if (NSEG > 1) {
                const double nxt = t.a[r][(q + 1) % NSEG], prv = t.a[r][(q + NSEG - 1) % NSEG];
                send = (lane == 0) ? nxt : ((lane == 31) ? prv : send);
            }
            o[q] = __shfl_sync(0xffffffffu, send, partner);
        }
#pragma unroll
        for (int q = 0; q < NSEG; ++q) t.a[r][q] = c * t.a[r][q] + s * o[q];
    }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g2_y_even(Tile<NSEG, PY>& t, double c, double s) {
#pragma unroll
    for (int r = 0; r < PY; r += 2)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double t1 = t.a[r][q], t2 = t.a[r + 1][q];
            t.a[r][q] = c * t1 + s * t2;
            t.a[r + 1][q] = c * t2 + s * t1;
        }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g3_y_odd(Tile<NSEG, PY>& t, double c, double s, const double (&above)[NSEG],
                                         const double (&below)[NSEG]) {
#pragma unroll
    for (int r = 1; r + 1 < PY; r += 2)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double t1 = t.a[r][q], t2 = t.a[r + 1][q];
            t.a[r][q] = c * t1 + s * t2;
            t.a[r + 1][q] = c * t2 + s * t1;
        }
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        t.a[0][q] = c * t.a[0][q] + s * above[q];
        t.a[PY - 1][q] = c * t.a[PY - 1][q] + s * below[q];
    }
}

// ---- "tanh form" of the same colour groups --------------------------------------------------------------------------------
// A bond rotation (a, b) -> (c a + s b, s a + c b) equals c * (a + t b, b + t a) with t = s / c.  With one (c, s) per colour
// the four factors c commute with everything: K = (c0 c1 c2 c3) * prod_g (1 + t_g X_g), X_g = the pair swap of colour g.
// Latency-bound callers (the Chebyshev chains of the KPM preconditioner) fold the constant into a diagonal they multiply
// with anyway, and a colour costs ONE fused multiply-add per site instead of a multiply and a multiply-add.
template <int NSEG, int PY>
__device__ __forceinline__ void g0_x_even_t(Tile<NSEG, PY>& t, double th) {
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double o = __shfl_xor_sync(0xffffffffu, t.a[r][q], 1);
            t.a[r][q] = fma(th, o, t.a[r][q]);
        }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g1_x_odd_t(Tile<NSEG, PY>& t, double th, int lane) {
    const int partner = (lane & 1) ? ((lane + 1) & 31) : ((lane + 31) & 31);
#pragma unroll
    for (int r = 0; r < PY; ++r) {
        double o[NSEG];
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            double send = t.a[r][q];
            if (NSEG > 1) {
                const double nxt = t.a[r][(q + 1) % NSEG], prv = t.a[r][(q + NSEG - 1) % NSEG];
                send = (lane == 0) ? nxt : ((lane == 31) ? prv : send);
            }
            o[q] = __shfl_sync(0xffffffffu, send, partner);
        }
#pragma unroll
        for (int q = 0; q < NSEG; ++q) t.a[r][q] = fma(th, o[q], t.a[r][q]);
    }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g2_y_even_t(Tile<NSEG, PY>& t, double th) {
#pragma unroll
    for (int r = 0; r < PY; r += 2)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double t1 = t.a[r][q], t2 = t.a[r + 1][q];
            t.a[r][q] = fma(th, t2, t1);
            t.a[r + 1][q] = fma(th, t1, t2);
        }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g3_y_odd_t(Tile<NSEG, PY>& t, double th, const double (&above)[NSEG], const double (&below)[NSEG]) {
#pragma unroll
    for (int r = 1; r + 1 < PY; r += 2)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double t1 = t.a[r][q], t2 = t.a[r + 1][q];
            t.a[r][q] = fma(th, t2, t1);
            t.a[r + 1][q] = fma(th, t1, t2);
        }
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        t.a[0][q] = fma(th, above[q], t.a[0][q]);
        t.a[PY - 1][q] = fma(th, below[q], t.a[PY - 1][q]);
    }
}

// real tile: publish the edge rows, one barrier, fetch the neighbours' edge rows.  strip: [nwarps][2][LX].
template <int NSEG, int PY>
__device__ __forceinline__ void exchange_edges1(const Tile<NSEG, PY>& t, double* strip, int warp, int nwarps, int lane,
                                                double (&above)[NSEG], double (&below)[NSEG]) {
    constexpr int LX = 32 * NSEG;
    double* mine = strip + (size_t)warp * 2 * LX;
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        mine[32 * q + lane] = t.a[0][q];
        mine[LX + 32 * q + lane] = t.a[PY - 1][q];
    }
    __syncthreads();
    const int up = (warp == 0) ? nwarps - 1 : warp - 1;
    const int dn = (warp + 1 == nwarps) ? 0 : warp + 1;
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        above[q] = strip[(size_t)up * 2 * LX + LX + 32 * q + lane];
        below[q] = strip[(size_t)dn * 2 * LX + 32 * q + lane];
    }
}

// The same for a slice that is split into row strips over the CTAs of a thread-block cluster (kpm_square_wide_kernel): the first
// warp also sends its first row to the CTA holding the strip above and the last warp its last row to the CTA below, through
// distributed shared memory, as self-validating 16-byte words {value, sequence number} -- the receiving lane polls the word it
// needs, so the strips synchronise pairwise and per row instead of through the cluster barrier (measured ~800 cycles per sweep
// on an 8-CTA cluster, more than the sweep itself).  Two halo rows of LX words per buffer half, zeroed at kernel start:
//     halo[0][x] = last row of the strip above,  halo[1][x] = first row of the strip below.
// A neighbour can be at most one exchange ahead (it needs this CTA's row of the next exchange to go further), and consecutive
// exchanges alternate between the buffer halves, so a word is never overwritten before it has been read.
struct WideCtx {
    uint32_t up_rank, dn_rank;   // cluster ranks of the CTAs holding the strips above / below (periodic)
    unsigned long long seq;      // number of the current exchange (tags start at 1)
    ulonglong2* halo;            // [2 buffer halves][2][LX] in this CTA's shared memory
};

template <int NSEG, int PY>
__device__ __forceinline__ void exchange_edges1_wide(const Tile<NSEG, PY>& t, double* strip, int xbuf, int warp, int nwarps, int lane,
                                                     double (&above)[NSEG], double (&below)[NSEG], WideCtx& wc) {
    constexpr int LX = 32 * NSEG;
    double* mine = strip + (size_t)warp * 2 * LX;
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        mine[32 * q + lane] = t.a[0][q];
        mine[LX + 32 * q + lane] = t.a[PY - 1][q];
    }
    const unsigned long long seq = ++wc.seq;
    ulonglong2* halo = wc.halo + (size_t)xbuf * 2 * LX;
    if (warp == 0) {            // my first row is the row BELOW the last row of the strip above: its halo[1]
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(halo + LX);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(wc.up_rank));
#pragma unroll
        for (int q = 0; q < NSEG; ++q)
            asm volatile("st.shared::cluster.v2.b64 [%0], {%1, %2};" ::"r"(remote + (uint32_t)((32 * q + lane) * sizeof(ulonglong2))),
                         "l"(__double_as_longlong(t.a[0][q])), "l"(seq)
                         : "memory");
    }
    if (warp == nwarps - 1) {   // my last row is the row ABOVE the first row of the strip below: its halo[0]
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(halo);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(wc.dn_rank));
#pragma unroll
        for (int q = 0; q < NSEG; ++q)
            asm volatile("st.shared::cluster.v2.b64 [%0], {%1, %2};" ::"r"(remote + (uint32_t)((32 * q + lane) * sizeof(ulonglong2))),
                         "l"(__double_as_longlong(t.a[PY - 1][q])), "l"(seq)
                         : "memory");
    }
    __syncthreads();
    auto poll = [&](const ulonglong2* src) -> double {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(src);
        unsigned long long v, tag;
        unsigned int spins = 0;
        do {
            asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(tag) : "r"(a) : "memory");
        } while (tag != seq && ++spins < (1u << 24));   // bounded: a lost neighbour gives a wrong result, not a hang
        return __longlong_as_double((long long)v);
    };
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        above[q] = (warp == 0) ? poll(halo + 32 * q + lane) : strip[(size_t)(warp - 1) * 2 * LX + LX + 32 * q + lane];
        below[q] = (warp + 1 == nwarps) ? poll(halo + LX + 32 * q + lane) : strip[(size_t)(warp + 1) * 2 * LX + 32 * q + lane];
    }
}

// publish the edge rows of a complex tile (re, im), one barrier, fetch the neighbours' edge rows.
// strip: [nwarps][4][LX] doubles = (first row re, first row im, last row re, last row im) per warp.
template <int NSEG, int PY>
__device__ __forceinline__ void exchange_edges2(const Tile<NSEG, PY>& re, const Tile<NSEG, PY>& im, double* strip, int warp,
                                                int nwarps, int lane, double (&ab_re)[NSEG], double (&ab_im)[NSEG],
                                                double (&be_re)[NSEG], double (&be_im)[NSEG]) {
    constexpr int LX = 32 * NSEG;
    double* mine = strip + (size_t)warp * 4 * LX;
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        mine[0 * LX + 32 * q + lane] = re.a[0][q];
        mine[1 * LX + 32 * q + lane] = im.a[0][q];
        mine[2 * LX + 32 * q + lane] = re.a[PY - 1][q];
        mine[3 * LX + 32 * q + lane] = im.a[PY - 1][q];
    }
    __syncthreads();
    const int up = (warp == 0) ? nwarps - 1 : warp - 1;
    const int dn = (warp + 1 == nwarps) ? 0 : warp + 1;
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        ab_re[q] = strip[(size_t)up * 4 * LX + 2 * LX + 32 * q + lane];
        ab_im[q] = strip[(size_t)up * 4 * LX + 3 * LX + 32 * q + lane];
        be_re[q] = strip[(size_t)dn * 4 * LX + 0 * LX + 32 * q + lane];
        be_im[q] = strip[(size_t)dn * 4 * LX + 1 * LX + 32 * q + lane];
    }
}

// ---- honeycomb lattice, 32 unit cells wide -----------------------------------------------------------------------------------
// Sites are numbered 2 * (l1 + L1 * l2) + orbit (src/Lattices.jl:52-107), so a row l2 is 64 consecutive values and lane l1 holds
// the cell (A, B) = a[r][0], a[r][1] of PY rows.  The three checkerboard colours are the three bond types of
// examples/holstein_hmc_honeycomb.toml:46-64:  A(l1,l2)-B(l1,l2) = register pair;  A(l1,l2)-B(l1-1,l2) = lane rotation;
// A(l1,l2)-B(l1,l2-1) = row pair (registers inside a warp's rows, shared-memory strip between warps).
template <int PY>
__device__ __forceinline__ void hc0_cell(Tile<2, PY>& t, double c, double s) {
#pragma unroll
    for (int r = 0; r < PY; ++r) {
        const double A = t.a[r][0], B = t.a[r][1];
        t.a[r][0] = c * A + s * B;
        t.a[r][1] = c * B + s * A;
    }
}

template <int PY>
__device__ __forceinline__ void hc1_lane(Tile<2, PY>& t, double c, double s, int lane) {
#pragma unroll
    for (int r = 0; r < PY; ++r) {
        const double oA = __shfl_sync(0xffffffffu, t.a[r][1], (lane + 31) & 31);   // B of the cell to the left
        const double oB = __shfl_sync(0xffffffffu, t.a[r][0], (lane + 1) & 31);    // A of the cell to the right
        t.a[r][0] = c * t.a[r][0] + s * oA;
        t.a[r][1] = c * t.a[r][1] + s * oB;
    }
}

template <int PY>
__device__ __forceinline__ void hc2_row(Tile<2, PY>& t, double c, double s, double aboveB, double belowA) {
#pragma unroll
    for (int r = 1; r < PY; ++r) {
        const double A = t.a[r][0], B = t.a[r - 1][1];
        t.a[r][0] = c * A + s * B;
        t.a[r - 1][1] = c * B + s * A;
    }
    t.a[0][0] = c * t.a[0][0] + s * aboveB;
    t.a[PY - 1][1] = c * t.a[PY - 1][1] + s * belowA;
}

// publish the first row's A and the last row's B, one barrier, fetch the B of the row above and the A of the row below.
// strip: [nwarps][2][32]
template <int PY>
__device__ __forceinline__ void exchange_hc1(const Tile<2, PY>& t, double* strip, int warp, int nwarps, int lane, double& aboveB,
                                             double& belowA) {
    double* mine = strip + (size_t)warp * 64;
    mine[lane] = t.a[0][0];
    mine[32 + lane] = t.a[PY - 1][1];
    __syncthreads();
    const int up = (warp == 0) ? nwarps - 1 : warp - 1;
    const int dn = (warp + 1 == nwarps) ? 0 : warp + 1;
    aboveB = strip[(size_t)up * 64 + 32 + lane];
    belowA = strip[(size_t)dn * 64 + lane];
}

// two tiles, one barrier.  strip: [nwarps][4][32]
template <int PY>
__device__ __forceinline__ void exchange_hc2(const Tile<2, PY>& t, const Tile<2, PY>& u, double* strip, int warp, int nwarps, int lane,
                                             double& aboveB_t, double& belowA_t, double& aboveB_u, double& belowA_u) {
    double* mine = strip + (size_t)warp * 128;
    mine[lane] = t.a[0][0];
    mine[32 + lane] = t.a[PY - 1][1];
    mine[64 + lane] = u.a[0][0];
    mine[96 + lane] = u.a[PY - 1][1];
    __syncthreads();
    const int up = (warp == 0) ? nwarps - 1 : warp - 1;
    const int dn = (warp + 1 == nwarps) ? 0 : warp + 1;
    aboveB_t = strip[(size_t)up * 128 + 32 + lane];
    belowA_t = strip[(size_t)dn * 128 + lane];
    aboveB_u = strip[(size_t)up * 128 + 96 + lane];
    belowA_u = strip[(size_t)dn * 128 + 64 + lane];
}

// ---- colour groups with per-bond coefficients (SSH): (cosh, sinh) read from shared-memory tables -------------------
// tx: [PY][LX] double2 of the tile rows, entry (r, x) = bond (x,y)-(x+1,y);  ty: same for the bond (x,y)-(x,y+1);
// ty_halo: [LX] double2 = the y-table row above the tile (its y-odd bonds enter tile row 0).
template <int NSEG, int PY>
__device__ __forceinline__ void g0_tab(Tile<NSEG, PY>& t, const double2* __restrict__ tx, int lane) {
    constexpr int LX = 32 * NSEG;
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double2 cs = tx[r * LX + 32 * q + (lane & ~1)];
            const double o = __shfl_xor_sync(0xffffffffu, t.a[r][q], 1);
            t.a[r][q] = cs.x * t.a[r][q] + cs.y * o;
        }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g1_tab(Tile<NSEG, PY>& t, const double2* __restrict__ tx, int lane) {
    constexpr int LX = 32 * NSEG;
    const int partner = (lane & 1) ? ((lane + 1) & 31) : ((lane + 31) & 31);
#pragma unroll
    for (int r = 0; r < PY; ++r) {
        double o[NSEG];
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            double send = t.a[r][q];
            if (NSEG > 1) {
                const double nxt = t.a[r][(q + 1) % NSEG], prv = t.a[r][(q + NSEG - 1) % NSEG];
                send = (lane == 0) ? nxt : ((lane == 31) ? prv : send);
            }
            o[q] = __shfl_sync(0xffffffffu, send, partner);
        }
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            // the bond leaves the odd site: own column for odd lanes, the column to the left (periodic) for even lanes
            const int x = 32 * q + lane;
            const int xo = (lane & 1) ? x : ((x + LX - 1) % LX);
            const double2 cs = tx[r * LX + xo];
            t.a[r][q] = cs.x * t.a[r][q] + cs.y * o[q];
        }
    }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g2_tab(Tile<NSEG, PY>& t, const double2* __restrict__ ty, int lane) {
    constexpr int LX = 32 * NSEG;
#pragma unroll
    for (int r = 0; r < PY; r += 2)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double2 cs = ty[r * LX + 32 * q + lane];
            const double t1 = t.a[r][q], t2 = t.a[r + 1][q];
            t.a[r][q] = cs.x * t1 + cs.y * t2;
            t.a[r + 1][q] = cs.x * t2 + cs.y * t1;
        }
}

template <int NSEG, int PY>
__device__ __forceinline__ void g3_tab(Tile<NSEG, PY>& t, const double2* __restrict__ ty, const double2* __restrict__ ty_halo,
                                       int lane, const double (&above)[NSEG], const double (&below)[NSEG]) {
    constexpr int LX = 32 * NSEG;
#pragma unroll
    for (int r = 1; r + 1 < PY; r += 2)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const double2 cs = ty[r * LX + 32 * q + lane];
            const double t1 = t.a[r][q], t2 = t.a[r + 1][q];
            t.a[r][q] = cs.x * t1 + cs.y * t2;
            t.a[r + 1][q] = cs.x * t2 + cs.y * t1;
        }
#pragma unroll
    for (int q = 0; q < NSEG; ++q) {
        const double2 ca = ty_halo[32 * q + lane];                 // bond from the row above into tile row 0
        const double2 cb = ty[(PY - 1) * LX + 32 * q + lane];      // bond from tile row PY-1 into the row below
        t.a[0][q] = ca.x * t.a[0][q] + ca.y * above[q];
        t.a[PY - 1][q] = cb.x * t.a[PY - 1][q] + cb.y * below[q];
    }
}

}  // namespace sqt
