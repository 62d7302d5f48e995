// K5: coordination numbers -- for every selected molecule i of group 0, the number of molecules j of group 1 with
// 1e-3 < d(i,j) <= Rmin (reference src/mofs/calc_pair_dist_bulk.cpp:14-23, 66-92; src/mofs/calc_pair_dist.cpp:80-106).
//
// Same machinery as K3 (b2a_kernels_rdf.cuh): 32-bit box-fraction coordinates whose integer difference IS the minimum
// image, i-molecules in registers, j-molecules streamed through a shared tile, the whole-box or cell-list traversal.
// What differs is the per-pair work: no histogram, no square root -- one compare-and-count against a threshold on r^2.
// The FP32 value decides a pair only when it is outside a guard band around the threshold (relative FP32 error of r^2
// plus the absolute error of the 2^32 grid, both bounded on the host); pairs inside the band, and pairs closer than the
// 1e-3 nm self-pair cut, are queued per (j, thread) and re-evaluated in FP64 exactly as the reference computes them
// (periodicity loops, un-fused squares, IEEE sqrt), and the difference is applied to the count.  Counts are bit-exact.
#pragma once

#include "b2a_kernels_rdf.cuh"

namespace b2a
{

struct CoordFrame   // per frame, host-computed (x fixed units^2)
{
    float  t_in;      // r2 <= t_in  : proven  d <= Rmin
    float  t_mid, w;  // |r2 - t_mid| <= w : guard band around Rmin^2 -> undecided
    float  t_small;   // r2 <= t_small : may violate d > 1e-3 -> undecided
    float  ay, az;    // Ly/Lx, Lz/Lx
    double L[3];
};

struct CoordDev
{
    int    n0, n1;
    double rmin;            // double(float Rmin): `distance <= Rmin` compares in double
    int    jchunk, jsplit;  // brute-force mode: j range per block
};

struct CoordSelect   // region filter on i (calc_pair_dist_bulk.cpp:71; calc_pair_dist.cpp:82-86)
{
    int    dim, use_cyl;
    double low, up, cx, cy, rlow, rup;   // the program's floats widened
};

// slab_id[f][i] = 0 (selected) / -1, slab_cnt[f][0] = numA
__global__ void __launch_bounds__(256) k5_select(const double* __restrict__ cenA, int nframes, int n0, CoordSelect s,
                                                 int* __restrict__ slab_id, unsigned* __restrict__ slab_cnt)
{
    __shared__ unsigned scnt;
    const int f = blockIdx.y;
    if (threadIdx.x == 0) scnt = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n0 && f < nframes)
    {
        const double* c  = cenA + (size_t)f * 3 * n0;
        const double  pd = c[(size_t)s.dim * n0 + i];
        bool          ok = s.low <= pd && pd <= s.up;
        if (ok && s.use_cyl)
        {
            const double ax = __dsub_rn(c[i], s.cx), ay = __dsub_rn(c[(size_t)n0 + i], s.cy);
            const double L1 = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
            ok              = s.rlow <= L1 && L1 <= s.rup;
        }
        slab_id[(size_t)f * n0 + i] = ok ? 0 : -1;
        if (ok) atomicAdd(&scnt, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0 && f < nframes && scnt) atomicAdd(&slab_cnt[f], scnt);
}

template <bool CUBIC>
__device__ __forceinline__ float pair_r2(int xi, int yi, int zi, const int4& pj, float ay, float az)
{
    float fx = __int2float_rn(xi - pj.x);   // wraps mod 2^32 = minimum image
    float fy = __int2float_rn(yi - pj.y);
    float fz = __int2float_rn(zi - pj.z);
    if (!CUBIC) { fy = __fmul_rn(fy, ay); fz = __fmul_rn(fz, az); }
    return __fmaf_rn(fz, fz, __fmaf_rn(fy, fy, __fmul_rn(fx, fx)));
}

// two i-molecules against one j with packed f32x2 arithmetic (see pair_bins2): bit-identical lanes, half the issue slots
template <bool CUBIC>
__device__ __forceinline__ void pair_r2x2(int xa, int ya, int za, int xb, int yb, int zb, const int4& pj, unsigned long long ay2,
                                          unsigned long long az2, float& r2a, float& r2b)
{
    unsigned long long fx = f2pack(__int2float_rn(xa - pj.x), __int2float_rn(xb - pj.x));
    unsigned long long fy = f2pack(__int2float_rn(ya - pj.y), __int2float_rn(yb - pj.y));
    unsigned long long fz = f2pack(__int2float_rn(za - pj.z), __int2float_rn(zb - pj.z));
    if (!CUBIC) { fy = f2mul(fy, ay2); fz = f2mul(fz, az2); }
    const unsigned long long r2 = f2fma(fz, fz, f2fma(fy, fy, f2mul(fx, fx)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r2a), "=f"(r2b) : "l"(r2));
}

// calcDistance (calc_pair_dist_bulk.cpp:14-23) + the test at :81, operation for operation
__device__ __noinline__ int coord_exact(const double* __restrict__ cA, const double* __restrict__ cB, int n0, int n1, int i, int j,
                                        double L0, double L1, double L2, double rmin)
{
    const double L[3] = { L0, L1, L2 };
    double       ret  = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const double r = periodicity_d(__dsub_rn(cA[(size_t)k * n0 + i], cB[(size_t)k * n1 + j]), L[k]);
        ret            = __dadd_rn(ret, __dmul_rn(r, r));
    }
    const double distance = __dsqrt_rn(ret);
    return (distance > 1e-3 && distance <= rmin) ? 1 : 0;
}

struct CoordBlock   // block-uniform state for the cold path (shared memory)
{
    const int4*   A;
    const int4*   Bsorted;   // cell mode: j indexes the sorted array
    const double* cA;
    const double* cB;
    int*          cnt;       // global per-i counts of this frame
    int           tcount, n0, n1;
    float         t_in, t_mid, w, t_small, ay, az;
    double        L0, L1, L2, rmin;
    unsigned      nexact, noverflow;
};

template <int IPT, int THREADS, bool CUBIC>
__device__ __noinline__ void coord_recheck(CoordBlock* b, int t, int4 pj, int jglobal)
{
    const int jorig = b->Bsorted ? b->Bsorted[jglobal].w : jglobal;
#pragma unroll 1
    for (int q = 0; q < IPT; ++q)
    {
        const int il = t + q * THREADS;
        if (il >= b->tcount) break;
        const int4  a  = b->A[il];
        const float r2 = pair_r2<CUBIC>(a.x, a.y, a.z, pj, b->ay, b->az);
        if (!(fabsf(r2 - b->t_mid) <= b->w || r2 <= b->t_small)) continue;
        const int fast  = r2 <= b->t_in ? 1 : 0;
        const int exact = coord_exact(b->cA, b->cB, b->n0, b->n1, a.w, jorig, b->L0, b->L1, b->L2, b->rmin);
        atomicAdd(&b->nexact, 1u);
        if (exact != fast) atomicAdd(&b->cnt[a.w], exact - fast);
    }
}

// MODE 0: grid x = i-tile, y = j-split, z = frame.  MODE 2: grid x = cell of the i side, y = sub-tile, z = frame.
template <int IPT, int THREADS, bool CUBIC, int MODE>
__global__ void __launch_bounds__(THREADS) k5_coord(const int4* __restrict__ sortedA, const int4* __restrict__ fixB,
                                                    const double* __restrict__ cenA, const double* __restrict__ cenB,
                                                    const unsigned* __restrict__ slab_cnt, const CoordFrame* __restrict__ frames,
                                                    CoordDev p, CellDev cd, int* __restrict__ cnt_all,
                                                    unsigned long long* __restrict__ gstats)
{
    constexpr int TI = IPT * THREADS;
    __shared__ __align__(16) int4 sj[kTileJ];
    __shared__ unsigned           queue[kQueueCap];
    __shared__ unsigned           q_count[2];
    __shared__ CoordBlock         blk;

    const int f = blockIdx.z;
    int       tstart = 0, tcount = 0, j0 = 0, j1 = 0, cell_a0 = 0, cell_a1 = 0, ccx = 0, ccy = 0, ccz = 0;
    if (MODE == 0)
    {
        const int c = (int)slab_cnt[f];
        tstart      = blockIdx.x * TI;
        if (tstart >= c) return;
        tcount = min(TI, c - tstart);
        j0     = blockIdx.y * p.jchunk;
        j1     = min(p.n1, j0 + p.jchunk);
        if (j1 <= j0) return;
    }
    else
    {
        const unsigned  key = blockIdx.x;
        const unsigned* sa  = cd.startA + (size_t)f * (cd.nkeysA + 1);
        cell_a0 = (int)sa[key];
        cell_a1 = (int)sa[key + 1];
        if (cell_a1 <= cell_a0) return;
        const unsigned mask = (1u << cd.bits) - 1u;
        ccz    = (int)(key & mask);
        ccy    = (int)((key >> cd.bits) & mask);
        ccx    = (int)((key >> (2 * cd.bits)) & mask);
        tstart = cell_a0 + (int)blockIdx.y * TI;
        if (tstart >= cell_a1) return;
        tcount = min(TI, cell_a1 - tstart);
    }

    const CoordFrame fr = frames[f];
    const int4*      B  = MODE == 2 ? cd.sortedB + (size_t)f * p.n1 : fixB + (size_t)f * p.n1;
    if (threadIdx.x == 0)
    {
        q_count[0] = q_count[1] = 0;
        blk.A = sortedA + (size_t)f * p.n0 + tstart; blk.Bsorted = MODE == 2 ? B : nullptr;
        blk.cA = cenA + (size_t)f * 3 * p.n0; blk.cB = cenB + (size_t)f * 3 * p.n1;
        blk.cnt = cnt_all + (size_t)f * p.n0;
        blk.tcount = tcount; blk.n0 = p.n0; blk.n1 = p.n1;
        blk.t_in = fr.t_in; blk.t_mid = fr.t_mid; blk.w = fr.w; blk.t_small = fr.t_small; blk.ay = fr.ay; blk.az = fr.az;
        blk.L0 = fr.L[0]; blk.L1 = fr.L[1]; blk.L2 = fr.L[2]; blk.rmin = p.rmin;
        blk.nexact = blk.noverflow = 0;
    }
    const float t_in = fr.t_in, t_mid = fr.t_mid, wband = fr.w, t_small = fr.t_small, ay = fr.ay, az = fr.az;
    const unsigned long long ay2 = f2pack(ay, ay), az2 = f2pack(az, az);

    int                xi[IPT], yi[IPT], zi[IPT], cnt[IPT];
    int                sub = 0;
    unsigned long long npairs = 0;

    auto process_range = [&](int ja, int jb) {
        for (int jt = ja; jt < jb; jt += kTileJ)
        {
            __syncthreads();
            const int n = min(kTileJ, jb - jt);
            for (int t = threadIdx.x; t < n; t += THREADS) sj[t] = B[jt + t];
            __syncthreads();
            for (int js = 0; js < n; js += kSubJ, ++sub)
            {
                const int je = min(n, js + kSubJ);
                unsigned* qc = &q_count[sub & 1];
#pragma unroll 2
                for (int j = js; j < je; ++j)
                {
                    const int4 pj = sj[j];
                    float      m1 = 3.0e38f, m2 = 3.0e38f;
                    if (IPT % 2 == 0 && kPacked)
                    {
#pragma unroll
                        for (int q = 0; q < IPT; q += 2)
                        {
                            float r2a, r2b;
                            pair_r2x2<CUBIC>(xi[q], yi[q], zi[q], xi[q + 1], yi[q + 1], zi[q + 1], pj, ay2, az2, r2a, r2b);
                            cnt[q] += r2a <= t_in ? 1 : 0;
                            cnt[q + 1] += r2b <= t_in ? 1 : 0;
                            m1 = fminf(m1, fminf(fabsf(r2a - t_mid), fabsf(r2b - t_mid)));
                            m2 = fminf(m2, fminf(r2a, r2b));
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int q = 0; q < IPT; ++q)
                        {
                            const float r2 = pair_r2<CUBIC>(xi[q], yi[q], zi[q], pj, ay, az);
                            cnt[q] += r2 <= t_in ? 1 : 0;
                            m1 = fminf(m1, fabsf(r2 - t_mid));
                            m2 = fminf(m2, r2);
                        }
                    }
                    if (m1 <= wband || m2 <= t_small)   // some pair of this thread with j is undecided (lanes beyond tcount may trip it too)
                    {
                        const unsigned slot = atomicAdd(qc, 1u);
                        if (slot < (unsigned)kQueueCap) queue[slot] = ((unsigned)j << 16) | threadIdx.x;
                        else
                        {
                            coord_recheck<IPT, THREADS, CUBIC>(&blk, threadIdx.x, pj, jt + j);
                            atomicAdd(&blk.noverflow, 1u);
                        }
                    }
                }
                __syncthreads();
                const unsigned nq = min(*qc, (unsigned)kQueueCap);
                for (unsigned e = threadIdx.x; e < nq; e += THREADS)
                {
                    const unsigned ent = queue[e];
                    const int      jl  = (int)(ent >> 16);
                    coord_recheck<IPT, THREADS, CUBIC>(&blk, (int)(ent & 0xffffu), sj[jl], jt + jl);
                }
                if (threadIdx.x == 0) q_count[(sub + 1) & 1] = 0;
                __syncthreads();
            }
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
            xi[q] = a.x; yi[q] = a.y; zi[q] = a.z; cnt[q] = 0;
        }
    };
    auto store_tile = [&]() {
        const int4* A = sortedA + (size_t)f * p.n0 + tstart;
        int*        c = cnt_all + (size_t)f * p.n0;
#pragma unroll
        for (int q = 0; q < IPT; ++q)
        {
            const int il = threadIdx.x + q * THREADS;
            if (il < tcount && cnt[q]) atomicAdd(&c[A[il].w], cnt[q]);
        }
    };

    if (MODE == 0)
    {
        load_tile();
        process_range(j0, j1);
        store_tile();
    }
    else
    {
        const int       nc = 1 << cd.bits;
        const unsigned* sb = cd.startB + (size_t)f * (cd.nkeysB + 1);
        const int       nx = min(2 * cd.rx + 1, nc), ny = min(2 * cd.ry + 1, nc);
        for (; tstart < cell_a1; tstart += TI * (int)gridDim.y)
        {
            tcount = min(TI, cell_a1 - tstart);
            __syncthreads();
            if (threadIdx.x == 0) { blk.A = sortedA + (size_t)f * p.n0 + tstart; blk.tcount = tcount; }
            load_tile();
            for (int ix = 0; ix < nx; ++ix)
                for (int iy = 0; iy < ny; ++iy)
                {
                    const int      cx  = (ccx - (nx >> 1) + ix + nc) & (nc - 1);
                    const int      cy  = (ccy - (ny >> 1) + iy + nc) & (nc - 1);
                    const unsigned col = ((unsigned)cx << cd.bits | (unsigned)cy) << cd.bits;
                    if (2 * cd.rz + 1 >= nc) { process_range((int)sb[col], (int)sb[col + nc]); continue; }
                    const int z0 = ccz - cd.rz, z1 = ccz + cd.rz;
                    if (z0 < 0)
                    {
                        process_range((int)sb[col + z0 + nc], (int)sb[col + nc]);
                        process_range((int)sb[col], (int)sb[col + z1 + 1]);
                    }
                    else if (z1 >= nc)
                    {
                        process_range((int)sb[col + z0], (int)sb[col + nc]);
                        process_range((int)sb[col], (int)sb[col + z1 - nc + 1]);
                    }
                    else process_range((int)sb[col + z0], (int)sb[col + z1 + 1]);
                }
            store_tile();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (blk.nexact) atomicAdd(&gstats[2], (unsigned long long)blk.nexact);
        if (blk.noverflow) atomicAdd(&gstats[6], (unsigned long long)blk.noverflow);
        atomicAdd(&gstats[5], npairs);
    }
}

// counts -> per-frame histogram of counts (tmpPair of calc_pair_dist_bulk.cpp:86) and the running total over frames.
// The values cluster on a handful of bins, so they are privatised in shared memory and flushed once per block.
__global__ void __launch_bounds__(256) k5_hist(const int* __restrict__ cnt_all, const int* __restrict__ slab_id, int nframes, int n0,
                                               int max_count, unsigned* __restrict__ table, unsigned long long* __restrict__ hist,
                                               unsigned long long* __restrict__ gstats)
{
    extern __shared__ unsigned sh5[];   // max_count bins + selected + dropped
    const int f = blockIdx.y;
    for (int b = threadIdx.x; b < max_count + 2; b += blockDim.x) sh5[b] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n0 && f < nframes && slab_id[(size_t)f * n0 + i] >= 0)
    {
        const int c = cnt_all[(size_t)f * n0 + i];
        atomicAdd(&sh5[max_count], 1u);
        if (c < 0 || c >= max_count) atomicAdd(&sh5[max_count + 1], 1u);
        else atomicAdd(&sh5[c], 1u);
    }
    __syncthreads();
    if (f >= nframes) return;
    for (int b = threadIdx.x; b < max_count; b += blockDim.x)
        if (sh5[b])
        {
            atomicAdd(&table[(size_t)f * max_count + b], sh5[b]);
            atomicAdd(&hist[b], (unsigned long long)sh5[b]);
        }
    if (threadIdx.x == 0)
    {
        if (sh5[max_count]) atomicAdd(&gstats[4], (unsigned long long)sh5[max_count]);
        if (sh5[max_count + 1]) atomicAdd(&gstats[1], (unsigned long long)sh5[max_count + 1]);
    }
}

} // namespace b2a
