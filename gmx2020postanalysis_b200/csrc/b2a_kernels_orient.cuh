// K4: per-molecule orientation-angle histogram (reference src/mofs/calc_orient_mof.cpp:14-63, 150-200),
// fused with the centre reduction: one pass over the float rvec stream (12 B/atom), centres and the three
// reference atoms never leave registers.  HBM-bound.
//
// The angle bin is decided WITHOUT evaluating acos on the device: the host tabulates, with its own libm
// and long double arithmetic (exactly what the reference executes), the largest cos value belonging to
// each bin edge; the device compares.  All other arithmetic is FP64 with un-fused multiply-adds in the
// reference's operation order, so the histogram is bit-exact.
#pragma once

#include "b2a_kernels_center.cuh"

namespace b2a
{

struct OrientDev
{
    int    DIMN, nAngle, nbin1, tp1, tp2, tp3, DIMNOrient, method;   // tp* are 0-based here
    double low, up, c1, c2, dr, CR, Cr;
};

__device__ __forceinline__ double periodicity_o(double dx, double box)
{
    const double h = box / 2.0;   // gmx_handle.ipp:410-417
    while (dx > h) dx = __dsub_rn(dx, box);
    while (dx < -h) dx = __dadd_rn(dx, box);
    return dx;
}

__global__ void __launch_bounds__(256) k4_orient(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                 const int* __restrict__ index, int nmol, int napm,
                                                 const double* __restrict__ w, double wsum,
                                                 const FrameGeom* __restrict__ geom, OrientDev d,
                                                 const double* __restrict__ cosedge,
                                                 unsigned long long* __restrict__ hist,
                                                 unsigned long long* __restrict__ stats, int smem_bins)
{
    extern __shared__ double sm4[];
    double*   sw    = sm4;                    // napm
    double*   sedge = sm4 + napm;             // nAngle + 1
    unsigned* shist = reinterpret_cast<unsigned*>(sedge + d.nAngle + 1);
    for (int j = threadIdx.x; j < napm; j += blockDim.x) sw[j] = w[j];
    for (int j = threadIdx.x; j <= d.nAngle; j += blockDim.x) sedge[j] = cosedge[j];
    for (int b = threadIdx.x; b < smem_bins; b += blockDim.x) shist[b] = 0;
    __syncthreads();
    const int  ncells   = d.nbin1 * d.nAngle;              // orient cells, then nbin1 NCMDens cells
    const bool use_smem = smem_bins >= ncells + d.nbin1;
    unsigned   ndrop = 0, nsel = 0;
    const long long total = (long long)nframes * nmol;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x)
    {
        const int       f = (int)(n / nmol), i = (int)(n % nmol);
        const FrameGeom g = geom[f];
        const float*    p = x + (size_t)f * frame_stride + (size_t)__ldg(index + (size_t)i * napm) * 3;
        // slab / cylinder filter on the centre (:158-177)
        const double cz = center_component(p, d.DIMN, napm, sw, wsum, g.L[d.DIMN], g.half[d.DIMN]);
        if (!(d.low <= cz && cz <= d.up)) continue;
        const int    ka = d.DIMN == 0 ? 1 : 0, kb = d.DIMN == 2 ? 1 : 2;
        const double t1 = __dsub_rn(center_component(p, ka, napm, sw, wsum, g.L[ka], g.half[ka]), d.c1);
        const double t2 = __dsub_rn(center_component(p, kb, napm, sw, wsum, g.L[kb], g.half[kb]), d.c2);
        const double R  = __dsqrt_rn(__dadd_rn(__dmul_rn(t1, t1), __dmul_rn(t2, t2)));
        if (!(d.Cr <= R && R <= d.CR)) continue;
        nsel++;
        const int meZ = __double2int_rz(__ddiv_rn(__dsub_rn(R, d.Cr), d.dr));
        if (meZ < 0 || meZ >= d.nbin1) { ndrop++; continue; }
        if (use_smem) atomicAdd(&shist[ncells + meZ], 1u);
        else atomicAdd(&hist[ncells + meZ], 1ull);
        // getPoint + calcVector (:14-38): unwrapped atoms tp1..3, then min-imaged differences
        double v1[3], v2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const double ref = (double)__ldg(p + k);
            const double p1  = unwrap((double)__ldg(p + 3 * d.tp1 + k), ref, g.L[k], g.half[k]);
            const double p2  = unwrap((double)__ldg(p + 3 * d.tp2 + k), ref, g.L[k], g.half[k]);
            const double p3  = unwrap((double)__ldg(p + 3 * d.tp3 + k), ref, g.L[k], g.half[k]);
            v1[k]            = periodicity_o(__dsub_rn(p1, p2), g.L[k]);
            v2[k]            = periodicity_o(__dsub_rn(p1, p3), g.L[k]);
        }
        double vo[3];
        if (d.method == 2)
        {
            vo[0] = __dsub_rn(__dmul_rn(v1[1], v2[2]), __dmul_rn(v1[2], v2[1]));
            vo[1] = __dsub_rn(__dmul_rn(v1[2], v2[0]), __dmul_rn(v1[0], v2[2]));
            vo[2] = __dsub_rn(__dmul_rn(v1[0], v2[1]), __dmul_rn(v1[1], v2[0]));
        }
        else
        {
            vo[0] = __dadd_rn(v1[0], v2[0]); vo[1] = __dadd_rn(v1[1], v2[1]); vo[2] = __dadd_rn(v1[2], v2[2]);
        }
        // calcAngle (:41-47) up to the cosine; wall = unit vector along DIMNOrient so r2 == 1 and the dot
        // product is vo[DIMNOrient] exactly (the other terms are signed zeros)
        const double r1 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vo[0], vo[0]), __dmul_rn(vo[1], vo[1])), __dmul_rn(vo[2], vo[2])));
        const double wl[3] = { d.DIMNOrient == 0 ? 1.0 : 0.0, d.DIMNOrient == 1 ? 1.0 : 0.0, d.DIMNOrient == 2 ? 1.0 : 0.0 };
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn(vo[0], wl[0]), __dmul_rn(vo[1], wl[1])), __dmul_rn(vo[2], wl[2]));
        const double cs  = __ddiv_rn(dot, __dmul_rn(r1, 1.0));
        if (!(cs >= -1.0 && cs <= 1.0)) { ndrop++; continue; }   // acos -> NaN in the reference (undefined bin)
        // bin = #{k in 1..nAngle : cs <= edge[k]} by binary search over the decreasing edge table
        int lo = 0, hi = d.nAngle + 1;   // invariant: cs <= edge[lo] (edge[0] = 2), cs > edge[hi] (virtual)
        while (hi - lo > 1)
        {
            const int mid = (lo + hi) >> 1;
            if (cs <= sedge[mid]) lo = mid; else hi = mid;
        }
        const int meA = lo;
        if (meA >= d.nAngle) { ndrop++; continue; }              // e.g. angle == 180 (:194-196) -> out of bounds
        const int cell = meZ + meA * d.nbin1;
        if (use_smem) atomicAdd(&shist[cell], 1u);
        else atomicAdd(&hist[cell], 1ull);
    }
    __syncthreads();
    if (use_smem)
        for (int b = threadIdx.x; b < ncells + d.nbin1; b += blockDim.x)
            if (shist[b]) atomicAdd(&hist[b], (unsigned long long)shist[b]);
    for (int o = 16; o > 0; o >>= 1)
    {
        ndrop += __shfl_xor_sync(0xffffffffu, ndrop, o);
        nsel += __shfl_xor_sync(0xffffffffu, nsel, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (ndrop) atomicAdd(&stats[1], (unsigned long long)ndrop);
        if (nsel) atomicAdd(&stats[4], (unsigned long long)nsel);
    }
}

} // namespace b2a
