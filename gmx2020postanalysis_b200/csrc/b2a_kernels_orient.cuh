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
    double   low, up, c1, c2, CR, Cr;
    DivConst dr;
};

__device__ __forceinline__ double periodicity_o(double dx, double box, float hsafe)
{
    if (fabsf((float)dx) < hsafe) return dx;   // |dx| < L/2 proven in FP32 (hsafe is a few ulp below L/2): loops would not run
    const double h = box / 2.0;                // gmx_handle.ipp:410-417
    while (dx > h) dx = __dsub_rn(dx, box);
    while (dx < -h) dx = __dadd_rn(dx, box);
    return dx;
}

// One molecule: NCMDens cell (cell_z = ncells + meZ) and orient cell (cell_a = meZ + meA*nbin1), -1 where nothing is
// counted.  Mirrors calc_orient_mof.cpp:156-199 statement by statement.
template <bool SMEM>
__device__ __forceinline__ void orient_molecule(const float* p, int napm, const double* sw, const DivConst& wsum,
                                                const FrameGeom& g, const OrientDev& d, const double* sedge, int& cell_z,
                                                int& cell_a, unsigned& nsel, unsigned& ndrop)
{
    cell_z = cell_a = -1;
    const float r0 = ldf<SMEM>(p), r1 = ldf<SMEM>(p + 1), r2 = ldf<SMEM>(p + 2);
    double      c0, c1, c2;
    mol_centers3<SMEM>(p, napm, sw, wsum, g, c0, c1, c2);
    // slab / cylinder filter on the centre (:158-177)
    const double cz = d.DIMN == 0 ? c0 : (d.DIMN == 1 ? c1 : c2);
    if (!(d.low <= cz && cz <= d.up)) return;
    const double t1 = __dsub_rn(d.DIMN == 0 ? c1 : c0, d.c1);
    const double t2 = __dsub_rn(d.DIMN == 2 ? c1 : c2, d.c2);
    const double R  = __dsqrt_rn(__dadd_rn(__dmul_rn(t1, t1), __dmul_rn(t2, t2)));
    if (!(d.Cr <= R && R <= d.CR)) return;
    nsel++;
    const int meZ = __double2int_rz(div_const(__dsub_rn(R, d.Cr), d.dr));
    if (meZ < 0 || meZ >= d.nbin1) { ndrop++; return; }
    cell_z = d.nbin1 * d.nAngle + meZ;
    // getPoint + calcVector (:14-38): unwrapped atoms tp1..3, min-imaged differences
    const float rr[3] = { r0, r1, r2 };
    double      v1[3], v2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const double p1 = unwrap_f(ldf<SMEM>(p + 3 * d.tp1 + k), rr[k], g.L[k], g.half[k], g.hsafe[k]);
        const double p2 = unwrap_f(ldf<SMEM>(p + 3 * d.tp2 + k), rr[k], g.L[k], g.half[k], g.hsafe[k]);
        const double p3 = unwrap_f(ldf<SMEM>(p + 3 * d.tp3 + k), rr[k], g.L[k], g.half[k], g.hsafe[k]);
        v1[k]           = periodicity_o(__dsub_rn(p1, p2), g.L[k], g.hsafe[k]);
        v2[k]           = periodicity_o(__dsub_rn(p1, p3), g.L[k], g.hsafe[k]);
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
    // calcAngle (:41-47) up to the cosine; wall = unit vector along DIMNOrient so r2 == 1 and the dot product is
    // vo[DIMNOrient] exactly (the other terms are signed zeros)
    const double rv = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vo[0], vo[0]), __dmul_rn(vo[1], vo[1])), __dmul_rn(vo[2], vo[2])));
    const double wl[3] = { d.DIMNOrient == 0 ? 1.0 : 0.0, d.DIMNOrient == 1 ? 1.0 : 0.0, d.DIMNOrient == 2 ? 1.0 : 0.0 };
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(vo[0], wl[0]), __dmul_rn(vo[1], wl[1])), __dmul_rn(vo[2], wl[2]));
    const double cs  = __ddiv_rn(dot, __dmul_rn(rv, 1.0));
    if (!(cs >= -1.0 && cs <= 1.0)) { ndrop++; return; }   // acos -> NaN in the reference (undefined bin)
    // bin = #{k in 1..nAngle : cs <= edge[k]} by binary search over the decreasing edge table
    int lo = 0, hi = d.nAngle + 1;   // invariant: cs <= edge[lo] (edge[0] = 2), cs > edge[hi] (virtual)
    while (hi - lo > 1)
    {
        const int mid = (lo + hi) >> 1;
        if (cs <= sedge[mid]) lo = mid; else hi = mid;
    }
    if (lo >= d.nAngle) { ndrop++; return; }                 // e.g. angle == 180 (:194-196) -> out of bounds
    cell_a = meZ + lo * d.nbin1;
}

// direct-load kernel (index group not contiguous); the TMA-staged variant is k_stream<2>
__global__ void __launch_bounds__(256) k4_orient(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                 const int* __restrict__ index, int nmol, int napm,
                                                 const double* __restrict__ w, DivConst wsum,
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
    const int  i  = blockIdx.x * blockDim.x + threadIdx.x;
    const int  f0 = blockIdx.y * kFramesPerBlock, f1 = min(nframes, f0 + kFramesPerBlock);
    const size_t moff = i < nmol ? (size_t)__ldg(index + (size_t)i * napm) * 3 : 0;
    for (int f = f0; f < f1 && i < nmol; ++f)
    {
        int cz, ca;
        const FrameGeom g = geom[f];
        orient_molecule<false>(x + (size_t)f * frame_stride + moff, napm, sw, wsum, g, d, sedge, cz, ca, nsel, ndrop);
        if (cz >= 0) { if (use_smem) atomicAdd(&shist[cz], 1u); else atomicAdd(&hist[cz], 1ull); }
        if (ca >= 0) { if (use_smem) atomicAdd(&shist[ca], 1u); else atomicAdd(&hist[ca], 1ull); }
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
