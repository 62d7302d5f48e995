// K3: all-pairs minimum-image RDF histogram (reference src/general/calc_rdf_region.cpp:73-95) for sm_100a.
//
// Design (measured on B200, see profiles/r01_ubench.md):
//  * Issue-bound, not HBM- or atomic-bound: every pair costs ~16 issue slots, so the kernel is built to
//    minimise instructions per pair.
//  * Coordinates are 32-bit box fractions (x_fix = rint(x * 2^32/L) mod 2^32, written by K1).  The integer
//    difference wraps mod 2^32, which IS the minimum image: no floor/round per dimension.
//  * i-molecules live in registers (IPT per thread), j-molecules are streamed through shared memory and
//    read as one broadcast LDS.128 per j per warp.
//  * bin = mantissa of fma(r, scale, 2^23 - 0.5): float->int conversion and floor in the FMA itself.
//    Two FMAs with scale*(1-eps) and scale*(1+eps) bracket the FP32 error; if they agree the bin is proven
//    (DESIGN.md "K3 error bound"), otherwise the pair is queued and re-evaluated in FP64 exactly as the
//    reference does (double periodicity loops, un-fused multiply-adds, IEEE sqrt and divide).  Counts are
//    therefore bit-exact with no exceptions.
//  * The shared histogram covers every reachable bin (r up to the half box diagonal), so the atomic is
//    unconditional: no range test, no predicate, no divergence.  Bins >= Rbin are simply never flushed.
//    Queued pairs are counted first (in the low-estimate bin) and CORRECTED by the exact path.
//  * Slabs (`meZ`) and the region filter depend on i only (calc_rdf_region.cpp:75-77).  A prep pass drops
//    filtered-out i and groups the rest by slab, so a block sees one slab and needs one histogram.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace b2a
{

constexpr unsigned kMagicBits = 0x4B000000u;   // bits of 2^23 (float)
constexpr float    kMagic     = 8388607.5f;    // 2^23 - 0.5: fma(r, s, kMagic) rounds to floor-ish(r*s) in the mantissa
constexpr int      kQueueCap  = 1024;          // exact-path queue entries per block
constexpr int      kTileJ     = 1024;          // j-molecules staged per shared-memory tile
constexpr int      kSubJ      = 256;           // j-iterations between queue-drain checks

struct RdfFrame   // per frame (the box may change frame to frame); host-computed
{
    float    s_lo, s_hi;   // bins per x-fixed-unit, times (1-eps) rounded down / (1+eps) rounded up
    float    ay, az;       // Ly/Lx, Lz/Lx (exactly 1 when cubic)
    unsigned lowk_bits;    // kMagicBits + LOWK: bins below LOWK are always re-evaluated (absolute error term)
    unsigned pad;
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
    int      jchunk, jsplit;
    int      clamp;
};

// ---- prep: slab id per i, counts per slab -----------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_slab_count(const double* __restrict__ cenA, int nframes, RdfDev p,
                                                     int* __restrict__ slab_id, unsigned* __restrict__ slab_cnt)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n0 || f >= nframes) return;
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
    if (s >= 0) atomicAdd(&slab_cnt[(size_t)f * p.nslab + s], 1u);
}

// scatter: sortedA[f][off[slab] + k] = {fix.xyz, i}; order inside a slab is arbitrary (integer sums commute)
__global__ void __launch_bounds__(256) k3_slab_scatter(const int4* __restrict__ fixA, const int* __restrict__ slab_id,
                                                       int nframes, int n0, int nslab,
                                                       const unsigned* __restrict__ slab_cnt,
                                                       unsigned* __restrict__ cursor, int4* __restrict__ sortedA)
{
    extern __shared__ unsigned soff[];
    const int f = blockIdx.y;
    if (f >= nframes) return;
    if (threadIdx.x == 0)
    {
        unsigned acc = 0;
        for (int s = 0; s < nslab; ++s) { soff[s] = acc; acc += slab_cnt[(size_t)f * nslab + s]; }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    const int s = slab_id[(size_t)f * n0 + i];
    if (s < 0) return;
    const unsigned pos = soff[s] + atomicAdd(&cursor[(size_t)f * nslab + s], 1u);
    int4           v   = fixA[(size_t)f * n0 + i];
    v.w                = i;
    sortedA[(size_t)f * n0 + pos] = v;
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
__device__ __forceinline__ int exact_bin(const double* __restrict__ cA, const double* __restrict__ cB, int n0, int n1,
                                         int i, int j, const double L[3], const RdfDev& p)
{
    double R = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const double r = periodicity_d(__dsub_rn(cA[(size_t)k * n0 + i], cB[(size_t)k * n1 + j]), L[k]);
        R              = __dadd_rn(R, __dmul_rn(r, r));
    }
    if (!(1e-6 < R && R < p.rm2)) return -1;
    const int meR = __double2int_rz(__ddiv_rn(__dmul_rn(p.Rbin_d, __dsqrt_rn(R)), p.Rm));
    return (meR >= p.Rbin) ? -2 : meR;
}

// ---- main kernel ------------------------------------------------------------------------------------------
// grid: x = i-tile (over slabs), y = j-split, z = frame.  SYM: same group on both sides, i<j tiles only.
template <int IPT, int THREADS, bool CLAMP, bool CUBIC, bool PARANOID>
__global__ void __launch_bounds__(THREADS) k3_rdf(const int4* __restrict__ sortedA, const int4* __restrict__ fixB,
                                                  const double* __restrict__ cenA, const double* __restrict__ cenB,
                                                  const unsigned* __restrict__ slab_cnt,
                                                  const RdfFrame* __restrict__ frames, RdfDev p,
                                                  unsigned long long* __restrict__ ghist,
                                                  unsigned long long* __restrict__ gstats)
{
    constexpr int TI = IPT * THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4*               sj    = reinterpret_cast<int4*>(smem_raw);                                    // kTileJ
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(smem_raw + kTileJ * sizeof(int4)); // kQueueCap
    unsigned*           hist  = reinterpret_cast<unsigned*>(queue + kQueueCap) + 4;                   // hist[-1] = trash
    __shared__ unsigned q_count;
    __shared__ unsigned s_stats[4];   // exact, moved, dropped, overflow

    const int f = blockIdx.z;
    // locate this block's tile: slabs in order, ceil(cnt/TI) tiles each
    int      slab = -1, tstart = 0, tcount = 0;
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
    }
    if (slab < 0) return;   // block-uniform

    for (unsigned k = threadIdx.x; k < p.kstride; k += THREADS) hist[(int)k - 1] = 0;
    if (threadIdx.x == 0) q_count = 0;
    if (threadIdx.x < 4) s_stats[threadIdx.x] = 0;

    const RdfFrame fr   = frames[f];
    const int4*    A    = sortedA + (size_t)f * p.n0 + tstart;
    const int4*    B    = fixB + (size_t)f * p.n1;
    const double*  cA   = cenA + (size_t)f * 3 * p.n0;
    const double*  cB   = cenB + (size_t)f * 3 * p.n1;
    const unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist);

    int      xi[IPT], yi[IPT], zi[IPT];
    unsigned mulq[IPT], cbq[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q)
    {
        const int  il = threadIdx.x + q * THREADS;
        const bool ok = il < tcount;
        const int4 a  = ok ? A[il] : make_int4(0, 0, 0, 0);
        xi[q] = a.x; yi[q] = a.y; zi[q] = a.z;
        mulq[q] = ok ? 4u : 0u;                                         // padding lanes hit the trash word only
        cbq[q]  = ok ? (hist_sa - kMagicBits * 4u) : (hist_sa - 4u);
    }

    const float clampf = __uint_as_float(p.rbin_bits);   // 2^23 + Rbin

    // exact re-evaluation + correction of one queued pair (counted earlier in bin k1idx)
    auto fix_pair = [&](int il, int j, int k1idx) {
        const int i  = A[il].w;
        const int ke = exact_bin(cA, cB, p.n0, p.n1, i, j, fr.L, p);
        atomicAdd(&s_stats[0], 1u);
        if (ke == -2) atomicAdd(&s_stats[2], 1u);
        const int target = ke >= 0 ? ke : -1;          // -1 = trash
        if (target != k1idx)
        {
            atomicSub(&hist[k1idx], 1u);
            atomicAdd(&hist[target], 1u);
            if (k1idx >= 0 && (k1idx < p.Rbin || target >= 0)) atomicAdd(&s_stats[1], 1u);
        }
    };

    const int j0 = blockIdx.y * p.jchunk;
    const int j1 = min(p.n1, j0 + p.jchunk);
    for (int jt = j0; jt < j1; jt += kTileJ)
    {
        __syncthreads();
        const int cnt = min(kTileJ, j1 - jt);
        for (int t = threadIdx.x; t < cnt; t += THREADS) sj[t] = B[jt + t];
        __syncthreads();
        for (int js = 0; js < cnt; js += kSubJ)
        {
            const int je = min(cnt, js + kSubJ);
#pragma unroll 2
            for (int j = js; j < je; ++j)
            {
                const int4 pj = sj[j];
                unsigned   k1[IPT], k2[IPT];
                unsigned   dacc = 0, kmin = 0xffffffffu;
#pragma unroll
                for (int q = 0; q < IPT; ++q)
                {
                    float fx = (float)(xi[q] - pj.x);   // wraps mod 2^32 = minimum image
                    float fy = (float)(yi[q] - pj.y);
                    float fz = (float)(zi[q] - pj.z);
                    if (!CUBIC) { fy *= fr.ay; fz *= fr.az; }
                    const float r2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
                    float       r;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
                    float v1 = fmaf(r, fr.s_lo, kMagic);
                    float v2 = fmaf(r, fr.s_hi, kMagic);
                    if (CLAMP) { v1 = fminf(v1, clampf); v2 = fminf(v2, clampf); }
                    k1[q] = __float_as_uint(v1);
                    k2[q] = __float_as_uint(v2);
                    unsigned addr;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(k1[q]), "r"(mulq[q]), "r"(cbq[q]));
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                    dacc += k2[q] - k1[q];
                    kmin = min(kmin, k1[q]);
                }
                if (PARANOID || dacc != 0 || kmin <= fr.lowk_bits)
                {
#pragma unroll
                    for (int q = 0; q < IPT; ++q)
                    {
                        if (mulq[q] == 0) continue;
                        const bool amb = PARANOID || (k1[q] != k2[q] && k1[q] < p.rbin_bits) || k1[q] <= fr.lowk_bits;
                        if (!amb) continue;
                        const int      k1idx = (int)(k1[q] - kMagicBits);
                        const unsigned il    = threadIdx.x + q * THREADS;
                        const unsigned slot  = atomicAdd(&q_count, 1u);
                        if (slot < (unsigned)kQueueCap)
                            queue[slot] = ((unsigned long long)(unsigned)(jt + j) << 32) | ((unsigned long long)(unsigned)(k1idx + 1) << 12) | il;
                        else
                        {
                            fix_pair((int)il, jt + j, k1idx);
                            atomicAdd(&s_stats[3], 1u);
                        }
                    }
                }
            }
            // drain the exact queue when it is at least half full, and after the last j.  Threads sample q_count
            // at different times, so the decision is made block-uniform by the barrier's OR-reduction.
            const bool last = (je == cnt) && (jt + kTileJ >= j1);
            if (__syncthreads_or(last || q_count >= (unsigned)(kQueueCap / 2)))
            {
                const unsigned n = min(q_count, (unsigned)kQueueCap);   // stable: every thread is past its pushes
                for (unsigned e = threadIdx.x; e < n; e += THREADS)
                {
                    const unsigned long long ent = queue[e];
                    fix_pair((int)(ent & 0xfffu), (int)(ent >> 32), (int)((ent >> 12) & 0xfffffu) - 1);
                }
                __syncthreads();
                if (threadIdx.x == 0) q_count = 0;
                __syncthreads();
            }
        }
    }
    __syncthreads();
    unsigned long long* gh = ghist + (size_t)slab * p.Rbin;
    for (int k = threadIdx.x; k < p.Rbin; k += THREADS)
    {
        const unsigned v = hist[k];
        if (v) atomicAdd(&gh[k], (unsigned long long)v);
    }
    if (threadIdx.x == 0)
    {
        if (s_stats[0]) atomicAdd(&gstats[2], (unsigned long long)s_stats[0]);
        if (s_stats[1]) atomicAdd(&gstats[3], (unsigned long long)s_stats[1]);
        if (s_stats[2]) atomicAdd(&gstats[1], (unsigned long long)s_stats[2]);
        if (s_stats[3]) atomicAdd(&gstats[6], (unsigned long long)s_stats[3]);
        atomicAdd(&gstats[5], (unsigned long long)tcount * (unsigned long long)(j1 > j0 ? j1 - j0 : 0));
        if (blockIdx.y == 0) atomicAdd(&gstats[4], (unsigned long long)tcount);
    }
}

} // namespace b2a
