// K3: all-pairs minimum-image RDF histogram (reference src/general/calc_rdf_region.cpp:73-95) for sm_100a.
//
// Design (measured on B200, see profiles/r01_*.md):
//  * Issue-bound, not HBM- or atomic-bound: the kernel is built to minimise instructions per pair
//    (~16 in the hot loop) and to keep the hot loop small enough for the instruction cache.
//  * Coordinates are 32-bit box fractions (x_fix = rint(x * 2^32/L) mod 2^32, written by K1).  The integer
//    difference wraps mod 2^32, which IS the minimum image: no floor/round per dimension.
//  * i-molecules live in registers (IPT per thread), j-molecules are streamed through shared memory and
//    read as one broadcast LDS.128 per j per warp.
//  * bin = mantissa of fma(r, scale, 2^23 - 0.5): float->int conversion and floor inside the FMA.
//    Two FMAs with scale*(1-eps) and scale*(1+eps) bracket the FP32 error; if they agree the bin is proven
//    (DESIGN.md "K3 error bound").  Otherwise the (j, thread) pair is queued in shared memory; every
//    kSubJ j-iterations the block drains the queue: it recomputes the FP32 bins of the queued candidates and
//    re-evaluates each undecided pair in FP64 exactly as the reference does (double periodicity loops,
//    un-fused multiply-adds, IEEE sqrt and divide).  Counts are bit-exact with no exceptions.
//  * The shared histogram covers every reachable bin (r up to the half box diagonal), so the atomic is
//    unconditional: no range test, no predicate, no divergence.  Bins >= Rbin are simply never flushed.
//    Undecided pairs are counted first (in the low-estimate bin) and CORRECTED by the exact path.
//  * Slabs (`meZ`) and the region filter depend on i only (calc_rdf_region.cpp:75-77).  A prep pass drops
//    filtered-out i and groups the rest by slab, so a block sees one slab and needs one histogram.
//  * SYM: when both groups are the same molecules and every molecule passes the filter into one slab, only
//    tile pairs with j-tile > i-tile are evaluated and flushed with weight 2 (periodicity(a-b) and
//    periodicity(b-a) give bit-identical squares because the inequalities in gmx_handle.ipp:412-415 are
//    strict); diagonal tiles are evaluated in full with weight 1.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "b2a_packed.cuh"

namespace b2a
{

constexpr unsigned kMagicBits = 0x4B000000u;   // bits of 2^23 (float)
constexpr float    kMagic     = 8388607.5f;    // 2^23 - 0.5: fma(r, s, kMagic) rounds to floor(r*s) in the mantissa
constexpr int      kQueueCap  = 1024;          // (j, thread) candidates per block between drains
#ifndef B2A_K3_TILEJ
#define B2A_K3_TILEJ 512
#endif
#ifndef B2A_K3_SUBJ
#define B2A_K3_SUBJ 256
#endif
constexpr int      kTileJ     = B2A_K3_TILEJ;  // j-molecules staged per shared-memory tile
constexpr int      kSubJ      = B2A_K3_SUBJ;   // j-iterations between queue drains
#ifndef B2A_K3_PACKED
#define B2A_K3_PACKED 1
#endif
#ifndef B2A_EXP_NOKMIN
#define B2A_EXP_NOKMIN 0   // experiment only (WRONG counts in the lowest bins): upper bound of what removing the low-bin test could save
#endif
constexpr bool     kPacked    = B2A_K3_PACKED != 0;   // issue the FP32 chain of two pairs as packed f32x2 instructions

struct RdfFrame   // per frame (the box may change frame to frame); host-computed
{
    float    s_lo, s_hi;   // bins per x-fixed-unit, times (1-eps) rounded down / (1+eps) rounded up
    float    ay, az;       // Ly/Lx, Lz/Lx (exactly 1 when cubic)
    float    r2max;        // CLAMP: squared distance (x fixed units) whose bracket bins are both Rbin
    unsigned lowk_bits;    // kMagicBits + LOWK: bins below LOWK are always re-evaluated (absolute error term)
    double   L[3];
};

struct RdfDev
{
    int      n0, n1, nbin, nslab, Rbin;   // nslab = nbin + 1 (slab nbin collects i whose meZ >= nbin: reference UB)
    int      dim, dim2;
    double   low, up, low2, up2, width;   // float bounds widened; width = double(float(up - low)) (:77)
    double   rm2;                         // double(float(Rm * Rm)) (:88)
    double   Rm, Rbin_d;
    unsigned rbin_bits;                   // kMagicBits + Rbin
    unsigned kstride;                     // words per shared histogram incl. the leading trash word
    int      jchunk, jsplit;              // full mode: j range per block; sym mode: jchunk = j per work item
    int      sym_enabled;                 // host: groups identical and symmetric evaluation requested
};

// ---- prep: slab id per i, counts per slab -----------------------------------------------------------------
// Counters are aggregated per block in shared memory first: with the program defaults every molecule of a frame
// lands in the same slab, and 72 000 global atomics on one address cost 0.35 ms (measured, profiles/r01_k3_tuning.md).
__global__ void __launch_bounds__(256) k3_slab_count(const double* __restrict__ cenA, int nframes, RdfDev p,
                                                     int* __restrict__ slab_id, unsigned* __restrict__ slab_cnt)
{
    extern __shared__ unsigned scnt[];
    const int f = blockIdx.y;
    for (int s = threadIdx.x; s < p.nslab; s += blockDim.x) scnt[s] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n0 && f < nframes)
    {
        const double* c   = cenA + (size_t)f * 3 * p.n0;
        const double  pd  = c[(size_t)p.dim * p.n0 + i];
        const double  pd2 = c[(size_t)p.dim2 * p.n0 + i];
        int           s   = -1;
        if (p.low < pd && pd < p.up && p.low2 < pd2 && pd2 < p.up2)   // strict, :75
        {
            // meZ = nbin * (posc - lowPos) / (upPos - lowPos), left to right in double, truncation (:77)
            const int meZ = __double2int_rz(__ddiv_rn(__dmul_rn((double)p.nbin, __dsub_rn(pd, p.low)), p.width));
            s             = (meZ < 0 || meZ >= p.nbin) ? p.nbin : meZ;
        }
        slab_id[(size_t)f * p.n0 + i] = s;
        if (s >= 0) atomicAdd(&scnt[s], 1u);
    }
    __syncthreads();
    if (f < nframes)
        for (int s = threadIdx.x; s < p.nslab; s += blockDim.x)
            if (scnt[s]) atomicAdd(&slab_cnt[(size_t)f * p.nslab + s], scnt[s]);
}

// scatter: sortedA[f][off[slab] + k] = {fix.xyz, i}; order inside a slab is arbitrary (integer sums commute)
__global__ void __launch_bounds__(256) k3_slab_scatter(const int4* __restrict__ fixA, const int* __restrict__ slab_id,
                                                       int nframes, int n0, int nslab,
                                                       const unsigned* __restrict__ slab_cnt,
                                                       unsigned* __restrict__ cursor, int4* __restrict__ sortedA)
{
    extern __shared__ unsigned ssm[];
    unsigned* soff  = ssm;               // slab start in the sorted array
    unsigned* lcnt  = ssm + nslab;       // this block's count per slab
    unsigned* lbase = ssm + 2 * nslab;   // this block's reserved range per slab
    const int f = blockIdx.y;
    if (f >= nframes) return;
    for (int s = threadIdx.x; s < nslab; s += blockDim.x) lcnt[s] = 0;
    if (threadIdx.x == 0)
    {
        unsigned acc = 0;
        for (int s = 0; s < nslab; ++s) { soff[s] = acc; acc += slab_cnt[(size_t)f * nslab + s]; }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int       s = -1;
    unsigned  rank = 0;
    if (i < n0)
    {
        s = slab_id[(size_t)f * n0 + i];
        if (s >= 0) rank = atomicAdd(&lcnt[s], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nslab; t += blockDim.x)
        if (lcnt[t]) lbase[t] = soff[t] + atomicAdd(&cursor[(size_t)f * nslab + t], lcnt[t]);
    __syncthreads();
    if (s < 0) return;
    int4 v = fixA[(size_t)f * n0 + i];
    v.w    = i;
    sortedA[(size_t)f * n0 + lbase[s] + rank] = v;
}

// ---- exact FP64 evaluation of one pair, operation for operation as calc_rdf_region.cpp:81-92 -----------------
__device__ __forceinline__ double periodicity_d(double dx, double box)
{
    // ipp:410-417; box / 2.0 is exact, so hoisting it changes nothing
    const double h = box / 2.0;
    while (dx > h) dx = __dsub_rn(dx, box);
    while (dx < -h) dx = __dadd_rn(dx, box);
    return dx;
}

// returns the reference bin (>= 0), -1 if the pair is not counted, -2 if the reference would write out of bounds
__device__ __noinline__ int exact_bin(const double* __restrict__ cA, const double* __restrict__ cB, int n0, int n1,
                                      int i, int j, double L0, double L1, double L2, double rm2, double Rbin_d,
                                      double Rm, int Rbin)
{
    const double L[3] = { L0, L1, L2 };
    double       R    = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const double r = periodicity_d(__dsub_rn(cA[(size_t)k * n0 + i], cB[(size_t)k * n1 + j]), L[k]);
        R              = __dadd_rn(R, __dmul_rn(r, r));
    }
    if (!(1e-6 < R && R < rm2)) return -1;
    const int meR = __double2int_rz(__ddiv_rn(__dmul_rn(Rbin_d, __dsqrt_rn(R)), Rm));
    return (meR >= Rbin) ? -2 : meR;
}

// ---- FP32 fast path for one pair: identical instruction sequence in the hot loop and in the drain ------------
template <bool CLAMP, bool CUBIC>
__device__ __forceinline__ void pair_bins(int xi, int yi, int zi, const int4& pj, float s_lo, float s_hi, float ay,
                                          float az, float r2max, unsigned& k1, unsigned& k2)
{
    float fx = __int2float_rn(xi - pj.x);   // wraps mod 2^32 = minimum image
    float fy = __int2float_rn(yi - pj.y);
    float fz = __int2float_rn(zi - pj.z);
    if (!CUBIC) { fy = __fmul_rn(fy, ay); fz = __fmul_rn(fz, az); }
    float r2 = __fmaf_rn(fz, fz, __fmaf_rn(fy, fy, __fmul_rn(fx, fx)));
    // CLAMP: every pair beyond the cut-off collapses onto r2max, chosen by the host so that both bracket bins of
    // sqrt(r2max) equal Rbin (the trash bin): one FMNMX, one shared address for ~half of the lanes, never undecided
    if (CLAMP) r2 = fminf(r2, r2max);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
    k1 = __float_as_uint(__fmaf_rn(r, s_lo, kMagic));
    k2 = __float_as_uint(__fmaf_rn(r, s_hi, kMagic));
}

// ---- the same, two pairs per instruction ---------------------------------------------------------------------------
// Blackwell has packed FP32 arithmetic (PTX fma.rn.f32x2 / mul.rn.f32x2, SASS FFMA2 / FMUL2): one issue slot, two IEEE
// round-to-nearest results on a 64-bit register pair.  The kernel is issue-bound, so the r^2 chain and the two bin FMAs of
// TWO i-molecules against the same j are issued together: 5 packed instructions replace 10.  Every lane result is
// bit-identical to pair_bins() (same operations, same rounding), which the out-of-line recheck path keeps using.
template <bool CLAMP, bool CUBIC>
__device__ __forceinline__ void pair_bins2(int xa, int ya, int za, int xb, int yb, int zb, const int4& pj, unsigned long long s_lo2,
                                           unsigned long long s_hi2, unsigned long long magic2, unsigned long long ay2,
                                           unsigned long long az2, float r2max, unsigned& k1a, unsigned& k2a, unsigned& k1b, unsigned& k2b)
{
    unsigned long long fx = f2pack(__int2float_rn(xa - pj.x), __int2float_rn(xb - pj.x));
    unsigned long long fy = f2pack(__int2float_rn(ya - pj.y), __int2float_rn(yb - pj.y));
    unsigned long long fz = f2pack(__int2float_rn(za - pj.z), __int2float_rn(zb - pj.z));
    if (!CUBIC) { fy = f2mul(fy, ay2); fz = f2mul(fz, az2); }
    const unsigned long long r2 = f2fma(fz, fz, f2fma(fy, fy, f2mul(fx, fx)));
    float r2a, r2b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r2a), "=f"(r2b) : "l"(r2));
    if (CLAMP) { r2a = fminf(r2a, r2max); r2b = fminf(r2b, r2max); }
    float ra, rb;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(r2a));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(r2b));
    const unsigned long long r  = f2pack(ra, rb);
    const unsigned long long k1 = f2fma(r, s_lo2, magic2), k2 = f2fma(r, s_hi2, magic2);
    asm("mov.b64 {%0, %1}, %2;" : "=r"(k1a), "=r"(k1b) : "l"(k1));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(k2a), "=r"(k2b) : "l"(k2));
}

struct RdfBlock   // block-uniform state shared by the hot loop and the cold paths (lives in shared memory)
{
    const int4*   A;        // this tile's i-molecules (sorted array + tile start)
    const int4*   Bsorted;  // SYM: j side indexes the sorted array; full mode: nullptr
    const double* cA;
    const double* cB;
    unsigned*     hist;     // shared histogram, hist[-1] = trash
    int           tcount, n0, n1, Rbin;
    unsigned      rbin_bits, lowk_bits;
    float         s_lo, s_hi, ay, az, r2max;
    double        L0, L1, L2, rm2, Rbin_d, Rm;
    unsigned      stats[4];   // exact, moved, dropped, overflow
};

// Re-examine the IPT pairs (thread t's molecules) x (one j): recompute the FP32 bins, and for every pair the
// FP32 bracket could not decide run the exact FP64 evaluation and move the count if needed.
template <int IPT, int THREADS, bool CLAMP, bool CUBIC, bool PARANOID>
__device__ __noinline__ void recheck(RdfBlock* b, int t, int4 pj, int jglobal)
{
    // phase 1 (converged across the warp): FP32 bins of the thread's IPT pairs, remember the undecided ones
    unsigned mask = 0;
    unsigned k1v[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q)
    {
        const int  il = t + q * THREADS;
        const bool ok = il < b->tcount;
        const int4 a  = ok ? b->A[il] : make_int4(0, 0, 0, 0);
        unsigned   k1, k2;
        pair_bins<CLAMP, CUBIC>(a.x, a.y, a.z, pj, b->s_lo, b->s_hi, b->ay, b->az, b->r2max, k1, k2);
        const bool amb = ok && (PARANOID || (k1 != k2 && k1 < b->rbin_bits) || k1 <= b->lowk_bits);
        k1v[q] = k1;
        mask |= amb ? (1u << q) : 0u;
    }
    // phase 2: one exact FP64 evaluation per undecided pair; lanes walk their own masks in lock step
    const int jorig = b->Bsorted ? b->Bsorted[jglobal].w : jglobal;
    while (mask)
    {
        const int q = __ffs(mask) - 1;
        mask &= mask - 1;
        unsigned k1 = 0;
#pragma unroll
        for (int u = 0; u < IPT; ++u) k1 = (u == q) ? k1v[u] : k1;
        const int k1idx = (int)(k1 - kMagicBits);
        const int iorig = b->A[t + q * THREADS].w;
        const int ke    = exact_bin(b->cA, b->cB, b->n0, b->n1, iorig, jorig, b->L0, b->L1, b->L2, b->rm2, b->Rbin_d, b->Rm, b->Rbin);
        atomicAdd(&b->stats[0], 1u);
        if (ke == -2) atomicAdd(&b->stats[2], 1u);
        const int target = ke >= 0 ? ke : -1;   // -1 = trash word
        if (target != k1idx)
        {
            atomicSub(&b->hist[k1idx], 1u);
            atomicAdd(&b->hist[target], 1u);
            if (k1idx >= 0 && (k1idx < b->Rbin || target >= 0)) atomicAdd(&b->stats[1], 1u);
        }
    }
}

// ---- cell mode prep (r_max << L): counting sort of both groups by cell ---------------------------------
// Cells are the top `bits` bits of the fixed-point box fractions, so the cell of a molecule costs three shifts and the
// periodic wrap of cell indices is a mask.  key = slab * nc^3 + (cx * nc + cy) * nc + cz  (z fastest: the neighbour cells
// of one (x, y) column are one or two contiguous runs of the sorted array).
__global__ void __launch_bounds__(256) k3_cell_keys(const int4* __restrict__ fix, const int* __restrict__ slab_id, int nframes,
                                                    int n, int bits, unsigned nkeys, int* __restrict__ keys,
                                                    unsigned* __restrict__ counts)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || f >= nframes) return;
    const int4     v  = fix[(size_t)f * n + i];
    const unsigned sh = 32u - (unsigned)bits;
    const unsigned cx = (unsigned)v.x >> sh, cy = (unsigned)v.y >> sh, cz = (unsigned)v.z >> sh;
    const unsigned cell = (((cx << bits) | cy) << bits) | cz;
    int            key  = (int)cell;
    if (slab_id)
    {
        const int s = slab_id[(size_t)f * n + i];
        key         = s < 0 ? -1 : (int)((unsigned)s << (3 * bits) | cell);
    }
    keys[(size_t)f * n + i] = key;
    if (key >= 0) atomicAdd(&counts[(size_t)f * (nkeys + 1) + key], 1u);
}

// exclusive scan of counts[f][0..nkeys) -> starts[f][0..nkeys]; one block per frame
__global__ void __launch_bounds__(1024) k3_cell_scan(const unsigned* __restrict__ counts, unsigned nkeys, unsigned* __restrict__ starts)
{
    __shared__ unsigned part[1024];
    const unsigned* c = counts + (size_t)blockIdx.x * (nkeys + 1);
    unsigned*       o = starts + (size_t)blockIdx.x * (nkeys + 1);
    const unsigned  chunk = (nkeys + 1023u) / 1024u;
    const unsigned  b0 = threadIdx.x * chunk, b1 = min(nkeys, b0 + chunk);
    unsigned        sum = 0;
    for (unsigned k = b0; k < b1; ++k) sum += c[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1)
    {
        const unsigned v = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - sum;   // exclusive prefix of this thread's chunk
    for (unsigned k = b0; k < b1; ++k) { o[k] = run; run += c[k]; }
    if (threadIdx.x == 1023) o[nkeys] = part[1023];
}

__global__ void __launch_bounds__(256) k3_cell_scatter(const int4* __restrict__ fix, const int* __restrict__ keys, int nframes,
                                                       int n, unsigned nkeys, const unsigned* __restrict__ starts,
                                                       unsigned* __restrict__ cursor, int4* __restrict__ sorted)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || f >= nframes) return;
    const int key = keys[(size_t)f * n + i];
    if (key < 0) return;
    const unsigned pos = starts[(size_t)f * (nkeys + 1) + key] + atomicAdd(&cursor[(size_t)f * (nkeys + 1) + key], 1u);
    int4           v   = fix[(size_t)f * n + i];
    v.w                = i;
    sorted[(size_t)f * n + pos] = v;
}

struct CellDev   // cell mode parameters (by value)
{
    int             bits;          // cells per dimension = 1 << bits
    int             rx, ry, rz;    // neighbour reach in cells per dimension (covers Rm plus the fixed-point slop)
    unsigned        nkeysA, nkeysB;
    const unsigned* startA;        // [frame][nkeysA + 1]
    const unsigned* startB;        // [frame][nkeysB + 1]
    const int4*     sortedB;       // [frame][n1]
    // z reach per neighbour column (dx, dy) = (ix - rx, iy - ry): the cut-off sphere, not its bounding cube.  Cells whose closest
    // points are further apart than the cut-off (plus the fixed-point slop) hold no countable pair: corner columns need fewer z
    // layers or none at all (-1).  Symmetric in the signs of dx, dy, dz, as the symmetric evaluation requires.  Host-computed;
    // all entries = rz when a reach covers its whole axis.
    signed char     rzc[15 * 15];
    int             use_rzc;       // 0: reaches above 7 cells (the table does not cover them): every column uses rz
    int             split;         // k3_rdf_cellw: the neighbour columns of a tile are dealt to this many blocks
};

// ---- main kernel ------------------------------------------------------------------------------------------
// MODE 0 (full) : grid x = i-tile (over slabs), y = j-split, z = frame
// MODE 1 (sym)  : grid x = work item (i-tile, j-chunk beyond the tile | diagonal), z = frame; frames that are not
//                 eligible (not all molecules in one slab) are skipped here and done by the full kernel, and vice versa
// (the cell-binned evaluation for r_max << L is a kernel of its own: k3_rdf_cellw in b2a_kernels_rdf_cell.cuh)
// (launch bound: 6 resident blocks of the 8 x 128 tile = 80 registers instead of the 96 the compiler takes unconstrained, 36 bytes
//  of spills: 634.7 -> 652.4 frames/s on config 2; 7 blocks = 72 registers spill into the hot loop: 603.7)
// DBUF: the j-tile is double-buffered -- tile t+1 is copied global -> shared asynchronously (cp.async, LDGSTS) while tile t is
// evaluated, one barrier per tile instead of two.  +1.4 % on config 2 (long j ranges); it costs 8 KB of shared memory, so the
// host selects it only when the resident-block count is unchanged (profiles/r02_k3_tuning.md: config 3 / 5 lose 2.7 % with it).
template <int IPT, int THREADS, bool CLAMP, bool CUBIC, bool PARANOID, int MODE, bool DBUF = false>
__global__ void __launch_bounds__(THREADS, (IPT == 8 && THREADS == 128) ? 6 : 1) k3_rdf(const int4* __restrict__ sortedA, const int4* __restrict__ fixB,
                                                  const double* __restrict__ cenA, const double* __restrict__ cenB,
                                                  const unsigned* __restrict__ slab_cnt,
                                                  const RdfFrame* __restrict__ frames, RdfDev p,
                                                  unsigned long long* __restrict__ ghist,
                                                  unsigned long long* __restrict__ gstats)
{
    constexpr int  TI  = IPT * THREADS;
    constexpr bool SYM = MODE == 1;
    constexpr bool kDbuf = DBUF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4*     sj    = reinterpret_cast<int4*>(smem_raw);                                   // kTileJ (x 2 when double-buffered)
    int4*     sjbuf = sj;
    unsigned* queue = reinterpret_cast<unsigned*>(smem_raw + (kDbuf ? 2 : 1) * kTileJ * sizeof(int4));       // kQueueCap
    unsigned* hist  = queue + kQueueCap + 4;                                               // hist[-1] = trash
    __shared__ unsigned q_count[2];
    __shared__ RdfBlock blk;

    const int f = blockIdx.z;
    int       slab = -1, tstart = 0, tcount = 0, j0 = 0, j1 = 0, weight = 1;
    {
        // is this frame eligible for symmetric evaluation?  (every molecule passed the filter into ONE slab)
        int sym_slab = -1;
        if (p.sym_enabled)
            for (int s = 0; s < p.nslab; ++s)
                if (slab_cnt[(size_t)f * p.nslab + s] == (unsigned)p.n0) sym_slab = s;
        if (SYM != (sym_slab >= 0)) return;   // block-uniform
        if (SYM)
        {
            // work items in order: for tile t = 0..T-1 : [diagonal], then chunks of [(t+1)*TI, n0)
            const int T = (p.n0 + TI - 1) / TI;
            int       w = blockIdx.x;
            for (int t = 0; t < T; ++t)
            {
                const int beyond = max(0, p.n0 - (t + 1) * TI);
                const int nch    = (beyond + p.jchunk - 1) / p.jchunk;
                if (w < 1 + nch)
                {
                    slab = sym_slab; tstart = t * TI; tcount = min(TI, p.n0 - tstart);
                    if (w == 0) { j0 = tstart; j1 = tstart + tcount; weight = 1; }
                    else { j0 = (t + 1) * TI + (w - 1) * p.jchunk; j1 = min(p.n0, j0 + p.jchunk); weight = 2; }
                    break;
                }
                w -= 1 + nch;
            }
        }
        else
        {
            int      t   = blockIdx.x;
            unsigned off = 0;
            for (int s = 0; s < p.nslab; ++s)
            {
                const unsigned c  = slab_cnt[(size_t)f * p.nslab + s];
                const int      nt = (int)((c + TI - 1) / TI);
                if (t < nt) { slab = s; tstart = (int)off + t * TI; tcount = min(TI, (int)c - t * TI); break; }
                t -= nt;
                off += c;
            }
            j0 = blockIdx.y * p.jchunk;
            j1 = min(p.n1, j0 + p.jchunk);
        }
        if (slab < 0 || j1 <= j0) return;   // block-uniform
    }

    const RdfFrame fr = frames[f];
    const int4*    B  = MODE == 1 ? sortedA + (size_t)f * p.n0 : fixB + (size_t)f * p.n1;
    for (unsigned k = threadIdx.x; k < p.kstride; k += THREADS) hist[(int)k - 1] = 0;
    if (threadIdx.x == 0)
    {
        q_count[0] = q_count[1] = 0;
        blk.A = sortedA + (size_t)f * p.n0 + tstart; blk.Bsorted = MODE != 0 ? B : nullptr;
        blk.cA = cenA + (size_t)f * 3 * p.n0;
        blk.cB = cenB + (size_t)f * 3 * p.n1;
        blk.hist = hist; blk.tcount = tcount; blk.n0 = p.n0; blk.n1 = p.n1; blk.Rbin = p.Rbin;
        blk.rbin_bits = p.rbin_bits; blk.lowk_bits = fr.lowk_bits;
        blk.s_lo = fr.s_lo; blk.s_hi = fr.s_hi; blk.ay = fr.ay; blk.az = fr.az;
        blk.r2max = fr.r2max;
        blk.L0 = fr.L[0]; blk.L1 = fr.L[1]; blk.L2 = fr.L[2];
        blk.rm2 = p.rm2; blk.Rbin_d = p.Rbin_d; blk.Rm = p.Rm;
        blk.stats[0] = blk.stats[1] = blk.stats[2] = blk.stats[3] = 0;
    }

    const unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist);
    const unsigned cb_ok   = hist_sa - kMagicBits * 4u;
    const float    r2max   = fr.r2max;
    const float    s_lo = fr.s_lo, s_hi = fr.s_hi, ay = fr.ay, az = fr.az;
    const unsigned lowk = fr.lowk_bits;
    const unsigned long long s_lo2 = f2pack(s_lo, s_lo), s_hi2 = f2pack(s_hi, s_hi), magic2 = f2pack(kMagic, kMagic),
                             ay2 = f2pack(ay, ay), az2 = f2pack(az, az);

    int                xi[IPT], yi[IPT], zi[IPT];
    int                sub = 0;   // sub-tile counter: selects the ping-pong queue counter
    unsigned long long npairs = 0, nsel = 0;

    // Candidate queue: entries are (global j << 7 | thread).  A drain re-examines only FULL rounds of THREADS entries -- a partly
    // filled round costs the same latency (scattered FP64 loads, a dependent FP64 chain) as a full one, and with ~1 candidate per j
    // the old "drain everything every 256 j" always paid for one (config 2: 257 entries = three rounds for two rounds' worth) -- and
    // carries the remainder (`carry` entries at the front of the queue, block-uniform) into the next phase.  Entries name j
    // globally, so they survive the tile they were found in; everything is drained before the i-tile changes and at the end.
    unsigned carry = 0;
    auto drain = [&](bool all) {
        __syncthreads();                                            // every push of the phase that just ended is in the queue
        const unsigned n   = min(carry + q_count[sub & 1], (unsigned)kQueueCap);
        const unsigned rem = all ? 0u : (n % (unsigned)THREADS);
        if (threadIdx.x == 0) q_count[(sub + 1) & 1] = 0;           // the next phase's counter (its last readers passed a barrier long ago)
#ifdef B2A_EXP_NODRAIN   // experiment only (WRONG counts): what the in-kernel re-evaluation of the candidates costs
        if (false)
#endif
        for (unsigned e = rem + threadIdx.x; e < n; e += THREADS)
        {
            const unsigned ent = queue[e];
            const int      jg  = (int)(ent >> 7);
            recheck<IPT, THREADS, CLAMP, CUBIC, PARANOID>(&blk, (int)(ent & 127u), B[jg], jg);
        }
        carry = rem;
        ++sub;
        __syncthreads();
    };

    // stream the j-molecules [ja, jb) of B against the i-molecules currently held in registers
    auto process_range = [&](int ja, int jb) {
        const bool partial = tcount < TI;
        auto prefetch = [&](int4* dst, int jt) {   // kDbuf: asynchronous global -> shared copy of one j-tile (LDGSTS, 16 bytes per thread and request)
            const int cnt = min(kTileJ, jb - jt);
            for (int t = threadIdx.x; t < cnt; t += THREADS)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + t)), "l"(B + jt + t) : "memory");
        };
        if (kDbuf && ja < jb)
        {
            __syncthreads();            // the previous range's last drain may still read the buffer about to be refilled
            sj = sjbuf;
            prefetch(sj, ja);
        }
        for (int jt = ja; jt < jb; jt += kTileJ)
        {
            const int cnt = min(kTileJ, jb - jt);
            if (kDbuf)
            {
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncthreads();        // tile jt is visible to everybody; everybody has left tile jt - kTileJ
                if (jt + kTileJ < jb) prefetch(sj == sjbuf ? sjbuf + kTileJ : sjbuf, jt + kTileJ);   // lands while this tile is evaluated
            }
            else
            {
                __syncthreads();
                for (int t = threadIdx.x; t < cnt; t += THREADS) sj[t] = B[jt + t];
                __syncthreads();
            }
            for (int js = 0; js < cnt; js += kSubJ)
            {
                const int je = min(cnt, js + kSubJ);
                unsigned* qc = &q_count[sub & 1];
                if (!partial)
                {
#pragma unroll 2
                    for (int j = js; j < je; ++j)
                    {
                        const int4 pj   = sj[j];
                        unsigned   dacc = 0, kmin = 0xffffffffu;
                        if (IPT % 2 == 0 && kPacked)
                        {
#pragma unroll
                            for (int q = 0; q < IPT; q += 2)
                            {
                                unsigned k1a, k2a, k1b, k2b;
                                pair_bins2<CLAMP, CUBIC>(xi[q], yi[q], zi[q], xi[q + 1], yi[q + 1], zi[q + 1], pj, s_lo2, s_hi2, magic2, ay2, az2,
                                                         r2max, k1a, k2a, k1b, k2b);
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(k1a * 4u + cb_ok) : "memory");
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(k1b * 4u + cb_ok) : "memory");
                                dacc += (k2a - k1a) + (k2b - k1b);
                                if (!B2A_EXP_NOKMIN) kmin = min(kmin, min(k1a, k1b));
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int q = 0; q < IPT; ++q)
                            {
                                unsigned k1, k2;
                                pair_bins<CLAMP, CUBIC>(xi[q], yi[q], zi[q], pj, s_lo, s_hi, ay, az, r2max, k1, k2);
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(k1 * 4u + cb_ok) : "memory");
                                dacc += k2 - k1;
                                kmin = min(kmin, k1);
                            }
                        }
                        if (PARANOID || dacc != 0 || kmin <= lowk)
                        {
                            const unsigned slot = carry + atomicAdd(qc, 1u);
                            if (slot < (unsigned)kQueueCap) queue[slot] = ((unsigned)(jt + j) << 7) | threadIdx.x;
                            else
                            {
                                recheck<IPT, THREADS, CLAMP, CUBIC, PARANOID>(&blk, threadIdx.x, pj, jt + j);
                                atomicAdd(&blk.stats[3], 1u);
                            }
                        }
                    }
                }
                else
                {
                    // partial tile (the last tile of a slab): same packed arithmetic; i-slots beyond tcount add into the trash word and
                    // never raise a candidate.  Full mode evaluates only the register slots that hold molecules: a tile of at most a
                    // quarter / a half of TI molecules runs on 2 / 4 of the 8 slots (config 3: each of the five slabs ends in a tile of
                    // ~100 molecules; +1.2 %).  The symmetric mode keeps the single loop: there the extra copies cost config 2 0.5 %.
                    auto part_loop = [&](auto qn_c) {
                        constexpr int QN = decltype(qn_c)::value;
                        bool          okq[QN];
#pragma unroll
                        for (int q = 0; q < QN; ++q) okq[q] = (int)(threadIdx.x + q * THREADS) < tcount;
#pragma unroll 1
                        for (int j = js; j < je; ++j)
                        {
                            const int4 pj   = sj[j];
                            unsigned   dacc = 0, kmin = 0xffffffffu;
                            if (QN % 2 == 0 && kPacked)
                            {
#pragma unroll
                                for (int q = 0; q < QN; q += 2)
                                {
                                    unsigned k1a, k2a, k1b, k2b;
                                    pair_bins2<CLAMP, CUBIC>(xi[q], yi[q], zi[q], xi[q + 1], yi[q + 1], zi[q + 1], pj, s_lo2, s_hi2, magic2, ay2, az2,
                                                             r2max, k1a, k2a, k1b, k2b);
                                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(okq[q] ? (k1a * 4u + cb_ok) : (hist_sa - 4u)) : "memory");
                                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(okq[q + 1] ? (k1b * 4u + cb_ok) : (hist_sa - 4u)) : "memory");
                                    dacc += (okq[q] ? (k2a - k1a) : 0u) + (okq[q + 1] ? (k2b - k1b) : 0u);
                                    kmin = min(kmin, min(okq[q] ? k1a : 0xffffffffu, okq[q + 1] ? k1b : 0xffffffffu));
                                }
                            }
                            else
                            {
#pragma unroll
                                for (int q = 0; q < QN; ++q)
                                {
                                    unsigned k1, k2;
                                    pair_bins<CLAMP, CUBIC>(xi[q], yi[q], zi[q], pj, s_lo, s_hi, ay, az, r2max, k1, k2);
                                    const unsigned addr = okq[q] ? (k1 * 4u + cb_ok) : (hist_sa - 4u);
                                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                                    dacc += okq[q] ? (k2 - k1) : 0u;
                                    kmin = okq[q] ? min(kmin, k1) : kmin;
                                }
                            }
                            if (PARANOID || dacc != 0 || kmin <= lowk)
                            {
                                const unsigned slot = carry + atomicAdd(qc, 1u);
                                if (slot < (unsigned)kQueueCap) queue[slot] = ((unsigned)(jt + j) << 7) | threadIdx.x;
                                else
                                {
                                    recheck<IPT, THREADS, CLAMP, CUBIC, PARANOID>(&blk, threadIdx.x, pj, jt + j);
                                    atomicAdd(&blk.stats[3], 1u);
                                }
                            }
                        }
                    };
                    if (MODE == 0 && IPT >= 8 && tcount <= (IPT / 4) * THREADS) part_loop(std::integral_constant<int, (MODE == 0 && IPT >= 8 ? IPT / 4 : IPT)>{});
                    else if (MODE == 0 && IPT >= 4 && tcount <= (IPT / 2) * THREADS) part_loop(std::integral_constant<int, (MODE == 0 && IPT >= 4 ? IPT / 2 : IPT)>{});
                    else part_loop(std::integral_constant<int, IPT>{});
                }
                drain(false);
            }
            if (kDbuf) sj = sj == sjbuf ? sjbuf + kTileJ : sjbuf;
        }
        npairs += (unsigned long long)tcount * (unsigned long long)(jb > ja ? jb - ja : 0);
    };

    auto load_tile = [&]() {
        const int4* A = sortedA + (size_t)f * p.n0 + tstart;
#pragma unroll
        for (int q = 0; q < IPT; ++q)
        {
            const int  il = threadIdx.x + q * THREADS;
            const int4 a  = il < tcount ? A[il] : make_int4(0, 0, 0, 0);
            xi[q] = a.x; yi[q] = a.y; zi[q] = a.z;
        }
    };

    load_tile();
    process_range(j0, j1);
    drain(true);
    nsel = SYM ? (weight == 1 ? tcount : 0) : (blockIdx.y == 0 ? tcount : 0);
    __syncthreads();
    unsigned long long* gh = ghist + (size_t)slab * p.Rbin;
    for (int k = threadIdx.x; k < p.Rbin; k += THREADS)
    {
        const unsigned v = hist[k];
        if (v) atomicAdd(&gh[k], (unsigned long long)v * (unsigned long long)weight);
    }
    if (threadIdx.x == 0)
    {
        if (blk.stats[0]) atomicAdd(&gstats[2], (unsigned long long)blk.stats[0]);
        if (blk.stats[1]) atomicAdd(&gstats[3], (unsigned long long)blk.stats[1] * (unsigned long long)weight);
        if (blk.stats[2]) atomicAdd(&gstats[1], (unsigned long long)blk.stats[2] * (unsigned long long)weight);
        if (blk.stats[3]) atomicAdd(&gstats[6], (unsigned long long)blk.stats[3]);
        atomicAdd(&gstats[5], npairs);
        if (nsel) atomicAdd(&gstats[4], nsel);
    }
}

} // namespace b2a
