// K3, cell-binned variant for r_max << L, warp-autonomous: every WARP owns one tile of 128 consecutive molecules of a
// z-sorted cell column (4 per lane, in registers) and walks its own neighbour runs.  Same arithmetic, bracket and exact FP64
// re-evaluation as k3_rdf (b2a_kernels_rdf.cuh: pair_bins2, recheck, exact_bin; reference src/general/calc_rdf_region.cpp:73-95),
// different plumbing:
//   * no block barrier between the first and the last pair: j-molecules are staged 32 at a time in a per-warp shared slot (one
//     coalesced LDG.128 per lane, the next chunk in flight while this one is evaluated, LDS.128 broadcast per j), candidates go to a
//     per-warp queue through a ballot (no atomic), and a warp drains its own queue in full rounds of 32;
//   * a 128-molecule tile spans ~3.5 cells of the column instead of the ~7 of a 256-molecule block tile (the round-1 design: one
//     block per tile, j staged per block), so with a reach of three cells it searches ~9.5 cells of every neighbour column instead
//     of 13 -- and unlike per-warp sub-ranges inside a block tile (measured: the block's barriers keep the critical path at 13
//     cells, +3 % only) the saving is not waited away;
//   * several j-iterations share one candidate test (B2A_CW_J2): independent dependency chains in flight, one vote and one branch;
//   * a block is one warp by default (B2A_CW_NW; several warps would share the histogram -- zeroed at the start, flushed with the
//     block's weight at the end -- and wait for each other there).
// Symmetric evaluation in cell mode (same molecules on both sides, every molecule of the frame in ONE slab): the j side is the i
// side's own sorted array, and a pair of sorted positions t1 < t2 is evaluated once, by the tile that holds t1, with weight 2 (the
// neighbourhood relation of cells is symmetric, so t1 lies in the j runs of t2's tile exactly when t2 lies in the j runs of t1's);
// the tiles against themselves are separate block rows with weight 1 (both orders of every pair, self pairs rejected by the exact path).
#pragma once

#include <type_traits>

#include "b2a_kernels_rdf.cuh"

namespace b2a
{

#ifndef B2A_CW_IPT
#define B2A_CW_IPT 4
#endif
constexpr int kCwIPT   = B2A_CW_IPT;     // i-molecules per lane
constexpr int kCwTile  = 32 * kCwIPT;
#ifndef B2A_CW_J2
#define B2A_CW_J2 4             // j-iterations per candidate test (measured on config 5: 1 -> 96.9 frames/s, 2 -> 106.0, 3 -> 105.5, 4 -> 107.4)
#endif
constexpr int kCwQueue = 96 + 32 * (B2A_CW_J2 > 1 ? B2A_CW_J2 : 1);   // candidate entries per warp; drained from 96 up, a test adds at most 32 per j-iteration
#ifndef B2A_CW_NW
#define B2A_CW_NW 1              // warps per block (measured on config 5 with four j per test: 1 -> 126.8 frames/s, 2 -> 121.1: a warp whose
#endif                           // tile is done -- a column's short last tile above all -- frees its slot at once instead of waiting for its sibling)
#ifndef B2A_CW_WARPS_PER_SM
#define B2A_CW_WARPS_PER_SM 20   // launch bound: 20 warps per SM = 96 registers
#endif
constexpr int kCwWarps = B2A_CW_NW;   // warps (= tiles in flight) per block

template <int NW, bool CUBIC, bool PARANOID>
__global__ void __launch_bounds__(32 * NW, B2A_CW_WARPS_PER_SM / NW)
    k3_rdf_cellw(const int4* __restrict__ sortedA, const double* __restrict__ cenA, const double* __restrict__ cenB,
                 const unsigned* __restrict__ slab_cnt, const RdfFrame* __restrict__ frames, RdfDev p, CellDev cd,
                 unsigned long long* __restrict__ ghist, unsigned long long* __restrict__ gstats)
{
    constexpr int IPT = kCwIPT, TW = kCwTile;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4*     wsj_all = reinterpret_cast<int4*>(smem_raw);                              // [NW][32]
    unsigned* wq_all  = reinterpret_cast<unsigned*>(smem_raw + NW * 32 * sizeof(int4)); // [NW][kCwQueue]
    unsigned* hist    = wq_all + NW * kCwQueue + 4;                                     // hist[-1] = trash
    __shared__ RdfBlock wblk[NW];

    const int      f = blockIdx.z, warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    const unsigned ltmask = (1u << lane) - 1u;
    int4* const     wsj = wsj_all + warp * 32;
    unsigned* const wq  = wq_all + warp * kCwQueue;
    RdfBlock* const blk = &wblk[warp];

    // block row -> column of i-cells (fixed slab, cx, cy): the sorted array holds a column's molecules contiguously, in order of cz, so
    // full tiles of TW molecules are cut from it regardless of how few molecules one cell holds; tiles are dealt to warps
    const unsigned  colA = blockIdx.x;
    const unsigned* sa   = cd.startA + (size_t)f * (cd.nkeysA + 1);
    const int       cell_a0 = (int)sa[colA << cd.bits], cell_a1 = (int)sa[(colA << cd.bits) + (1u << cd.bits)];
    const unsigned  cmask = (1u << cd.bits) - 1u;
    const int       slab = (int)(colA >> (2 * cd.bits)), ccy = (int)(colA & cmask), ccx = (int)((colA >> cd.bits) & cmask);
    bool csym = false;
    // block rows: [0, split * nrows) = tile rows x parts of the neighbour-column list; symmetric evaluation appends nrows rows
    // for the tiles against themselves (light: they fill the tail of the launch)
    const int split = cd.split;
    int       crole = 0, row = (int)blockIdx.y, part = 0, weight = 1;
    int       nrows = p.sym_enabled ? (int)gridDim.y / (split + 1) : (int)gridDim.y / split;
    if (row >= split * nrows) { crole = 1; row -= split * nrows; }
    else { part = row % split; row /= split; }
    if (p.sym_enabled)
    {
        csym = slab_cnt[(size_t)f * p.nslab + slab] == (unsigned)p.n0;
        if (crole == 1 && !csym) return;
        if (csym) weight = crole == 0 ? 2 : 1;
    }
    if (cell_a0 + row * NW * TW >= cell_a1) return;   // no tile for warp 0 of this row: none for the others either (block-uniform)

    const RdfFrame fr = frames[f];
    const int4*    A0 = sortedA + (size_t)f * p.n0;
    const int4*    B  = csym ? A0 : cd.sortedB + (size_t)f * p.n1;
    for (unsigned k = threadIdx.x; k < p.kstride; k += 32 * NW) hist[(int)k - 1] = 0;
    if (lane == 0)
    {
        blk->A = A0; blk->Bsorted = B;
        blk->cA = cenA + (size_t)f * 3 * p.n0; blk->cB = cenB + (size_t)f * 3 * p.n1;
        blk->hist = hist; blk->tcount = 0; blk->n0 = p.n0; blk->n1 = p.n1; blk->Rbin = p.Rbin;
        blk->rbin_bits = p.rbin_bits; blk->lowk_bits = fr.lowk_bits;
        blk->s_lo = fr.s_lo; blk->s_hi = fr.s_hi; blk->ay = fr.ay; blk->az = fr.az; blk->r2max = fr.r2max;
        blk->L0 = fr.L[0]; blk->L1 = fr.L[1]; blk->L2 = fr.L[2];
        blk->rm2 = p.rm2; blk->Rbin_d = p.Rbin_d; blk->Rm = p.Rm;
        blk->stats[0] = blk->stats[1] = blk->stats[2] = blk->stats[3] = 0;
    }
    __syncthreads();

    const unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist);
    const unsigned cb_ok   = hist_sa - kMagicBits * 4u, trash = hist_sa - 4u;
    const float    r2max   = fr.r2max;
    const unsigned lowk    = fr.lowk_bits;
    const unsigned long long s_lo2 = f2pack(fr.s_lo, fr.s_lo), s_hi2 = f2pack(fr.s_hi, fr.s_hi), magic2 = f2pack(kMagic, kMagic),
                             ay2 = f2pack(fr.ay, fr.ay), az2 = f2pack(fr.az, fr.az);

    int                xi[IPT], yi[IPT], zi[IPT];
    bool               okq[IPT];
    unsigned           qn = 0;          // entries in this warp's queue (warp-uniform)
    unsigned long long npairs = 0, nsel = 0;
    int                tstart = 0, tcount = 0;

    // re-examine the queued (j, lane) candidates: full rounds of 32 only, unless `all` (a partly filled round costs the same
    // latency as a full one); the remainder moves to the front of the queue
    auto drain = [&](bool all) {
        __syncwarp();
        const unsigned nfull = all ? qn : (qn & ~31u);
        for (unsigned e = (unsigned)lane; e < nfull; e += 32u)
        {
            const unsigned ent = wq[e];
            const int      jg  = (int)(ent >> 5);
            recheck<IPT, 32, true, CUBIC, PARANOID>(blk, (int)(ent & 31u), B[jg], jg);
        }
        __syncwarp();
        const unsigned rem = qn - nfull;
        unsigned       keep = 0;
        if ((unsigned)lane < rem) keep = wq[nfull + lane];
        __syncwarp();
        if ((unsigned)lane < rem) wq[lane] = keep;
        qn = rem;
        __syncwarp();
    };

    // `cnt` staged j-molecules (global positions from jc) against the first QN register slots; PART: slots beyond tcount exist
    auto chunk = [&](auto qn_c, auto part_c, int jc, int cnt) {
        constexpr int  QN   = decltype(qn_c)::value;
        constexpr bool PART = decltype(part_c)::value;
        auto eval = [&](const int4& pj, unsigned& dacc, unsigned& kmin) {
#pragma unroll
            for (int q = 0; q < QN; q += 2)
            {
                unsigned k1a, k2a, k1b, k2b;
                pair_bins2<true, CUBIC>(xi[q], yi[q], zi[q], xi[q + 1], yi[q + 1], zi[q + 1], pj, s_lo2, s_hi2, magic2, ay2, az2, r2max,
                                        k1a, k2a, k1b, k2b);
                if (!PART)
                {
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(k1a * 4u + cb_ok) : "memory");
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(k1b * 4u + cb_ok) : "memory");
                    dacc += (k2a - k1a) + (k2b - k1b);
                    kmin = min(kmin, min(k1a, k1b));
                }
                else
                {
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(okq[q] ? (k1a * 4u + cb_ok) : trash) : "memory");
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(okq[q + 1] ? (k1b * 4u + cb_ok) : trash) : "memory");
                    dacc += (okq[q] ? (k2a - k1a) : 0u) + (okq[q + 1] ? (k2b - k1b) : 0u);
                    kmin = min(kmin, min(okq[q] ? k1a : 0xffffffffu, okq[q + 1] ? k1b : 0xffffffffu));
                }
            }
        };
        int j = 0;
#if B2A_CW_J2 > 1
        // several j per candidate test: independent dependency chains in flight, one vote and one branch for all of them
        constexpr int JB = B2A_CW_J2;
#pragma unroll 1
        for (; j + JB <= cnt; j += JB)
        {
            unsigned d[JB], m[JB];
            bool     cj[JB], any = false;
#pragma unroll
            for (int u = 0; u < JB; ++u)
            {
                const int4 pj = wsj[j + u];
                d[u] = 0; m[u] = 0xffffffffu;
                eval(pj, d[u], m[u]);
            }
#pragma unroll
            for (int u = 0; u < JB; ++u) { cj[u] = PARANOID || d[u] != 0 || m[u] <= lowk; any = any || cj[u]; }
            if (__ballot_sync(0xffffffffu, any))   // warp-uniform
            {
#pragma unroll
                for (int u = 0; u < JB; ++u)
                {
                    const unsigned b = __ballot_sync(0xffffffffu, cj[u]);
                    if (cj[u]) wq[qn + __popc(b & ltmask)] = ((unsigned)(jc + j + u) << 5) | (unsigned)lane;
                    qn += __popc(b);
                }
                if (qn >= (unsigned)(kCwQueue - 32 * JB)) drain(false);
            }
        }
#else
#pragma unroll 2
#endif
        for (; j < cnt; ++j)
        {
            const int4 pj   = wsj[j];
            unsigned   dacc = 0, kmin = 0xffffffffu;
            eval(pj, dacc, kmin);
            const bool     cand = PARANOID || dacc != 0 || kmin <= lowk;
            const unsigned m    = __ballot_sync(0xffffffffu, cand);
            if (m)   // warp-uniform
            {
                if (cand) wq[qn + __popc(m & ltmask)] = ((unsigned)(jc + j) << 5) | (unsigned)lane;
                qn += __popc(m);
                if (qn >= (unsigned)(kCwQueue - 32 * (B2A_CW_J2 > 1 ? B2A_CW_J2 : 1))) drain(false);
            }
        }
    };

    // stream the j-molecules [ja, jb) of B against this warp's tile
    auto range = [&](int ja, int jb) {
        if (ja >= jb) return;
        npairs += (unsigned long long)tcount * (unsigned long long)(jb - ja);
        int4 nxt = make_int4(0, 0, 0, 0);
        if (ja + lane < jb) nxt = B[ja + lane];
        for (int jc = ja; jc < jb; jc += 32)
        {
            __syncwarp();                       // everybody has left the previous chunk
            wsj[lane] = nxt;
            __syncwarp();
            if (jc + 32 + lane < jb) nxt = B[jc + 32 + lane];   // in flight while this chunk is evaluated
            const int cnt = min(32, jb - jc);
            if (tcount == TW) chunk(std::integral_constant<int, IPT>{}, std::false_type{}, jc, cnt);
            else if (IPT < 4 || tcount > TW / 2) chunk(std::integral_constant<int, IPT>{}, std::true_type{}, jc, cnt);
            else chunk(std::integral_constant<int, (IPT >= 4 ? IPT / 2 : IPT)>{}, std::true_type{}, jc, cnt);   // at most half a tile: half the slots
        }
    };

    const int       nc  = 1 << cd.bits;
    const unsigned* sb  = csym ? cd.startA + (size_t)f * (cd.nkeysA + 1) + ((size_t)slab << (3 * cd.bits)) : cd.startB + (size_t)f * (cd.nkeysB + 1);
    const int       nx  = min(2 * cd.rx + 1, nc), ny = min(2 * cd.ry + 1, nc);
    const unsigned  zsh = 32u - (unsigned)cd.bits;
    const bool      diag = csym && crole == 1;       // the tile against itself, both orders (self pairs fall to the exact path)
    for (tstart = cell_a0 + (row * NW + warp) * TW; tstart < cell_a1; tstart += nrows * NW * TW)
    {
        drain(true);                                 // candidates of the previous tile name ITS molecules
        tcount = min(TW, cell_a1 - tstart);
        const int4* At = A0 + tstart;
        if (lane == 0) { blk->A = At; blk->tcount = tcount; }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < IPT; ++q)
        {
            const int il = lane + q * 32;
            okq[q]       = il < tcount;
            const int4 a = okq[q] ? At[il] : make_int4(0, 0, 0, 0);
            xi[q] = a.x; yi[q] = a.y; zi[q] = a.z;
        }
        if (!diag && part == 0) nsel += (unsigned long long)tcount;
        const int jmin = (csym && !diag) ? tstart + tcount : 0;   // symmetric: only sorted positions beyond the tile
        const int cza  = (int)((unsigned)At[0].z >> zsh), czb = (int)((unsigned)At[tcount - 1].z >> zsh);
        const int ncol = diag ? 1 : nx * ny;
        for (int ci = diag ? 0 : part; ci < ncol; ci += diag ? 1 : split)
        {
            int pa[2], pb[2], np = 1;
            if (diag) { pa[0] = tstart; pb[0] = tstart + tcount; }
            else
            {
                const int ix = ci / ny, iy = ci - ix * ny;
                const int rzc = cd.use_rzc ? (int)cd.rzc[ix * 15 + iy] : cd.rz;
                if (rzc < 0) continue;                    // the whole column lies outside the cut-off sphere
                const int      z0  = cza - rzc, z1 = czb + rzc;   // inclusive, may leave [0, nc)
                const int      cx  = (ccx - (nx >> 1) + ix + nc) & (nc - 1);
                const int      cy  = (ccy - (ny >> 1) + iy + nc) & (nc - 1);
                const unsigned col = ((unsigned)cx << cd.bits | (unsigned)cy) << cd.bits;
                if (z1 - z0 + 1 >= nc) { pa[0] = (int)sb[col]; pb[0] = (int)sb[col + nc]; }
                else if (z0 < 0)
                {
                    pa[0] = (int)sb[col + z0 + nc]; pb[0] = (int)sb[col + nc];
                    pa[1] = (int)sb[col]; pb[1] = (int)sb[col + z1 + 1]; np = 2;
                }
                else if (z1 >= nc)
                {
                    pa[0] = (int)sb[col + z0]; pb[0] = (int)sb[col + nc];
                    pa[1] = (int)sb[col]; pb[1] = (int)sb[col + z1 - nc + 1]; np = 2;
                }
                else { pa[0] = (int)sb[col + z0]; pb[0] = (int)sb[col + z1 + 1]; }
            }
            for (int k = 0; k < np; ++k) range(max(jmin, pa[k]), pb[k]);
        }
    }
    drain(true);
    __syncthreads();
    unsigned long long* gh = ghist + (size_t)slab * p.Rbin;
    for (int k = (int)threadIdx.x; k < p.Rbin; k += 32 * NW)
    {
        const unsigned v = hist[k];
        if (v) atomicAdd(&gh[k], (unsigned long long)v * (unsigned long long)weight);
    }
    if (lane == 0)
    {
        if (blk->stats[0]) atomicAdd(&gstats[2], (unsigned long long)blk->stats[0]);
        if (blk->stats[1]) atomicAdd(&gstats[3], (unsigned long long)blk->stats[1] * (unsigned long long)weight);
        if (blk->stats[2]) atomicAdd(&gstats[1], (unsigned long long)blk->stats[2] * (unsigned long long)weight);
        if (npairs) atomicAdd(&gstats[5], npairs);
        if (nsel) atomicAdd(&gstats[4], nsel);
    }
}

} // namespace b2a
