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
    float    Crf, inv_drf;   // FP32 copies for the starting guess of the radial bin
};

__device__ __forceinline__ double periodicity_o(double dx, double box, float hsafe)
{
    if (fabsf((float)dx) < hsafe) return dx;   // |dx| < L/2 proven in FP32 (hsafe is a few ulp below L/2): loops would not run
    const double h = box / 2.0;                // gmx_handle.ipp:410-417
    while (dx > h) dx = __dsub_rn(dx, box);
    while (dx < -h) dx = __dadd_rn(dx, box);
    return dx;
}

// Tables the host prepares with its own libm so that the device never evaluates sqrt/acos/divide on the common path
// and still reproduces the reference's bins exactly:
//   sedge[k], k = 0..nAngle  (double): largest cos value whose reference angle bin is >= k (sedge[0] = 2)
//   fedge[k]                 (float) : the same edges rounded to float, for the FP32 bracket
//   stab[k],  k = 0..nbin1   (double): smallest s = t1^2 + t2^2 whose R = RN(sqrt(s)) passes Cr <= R and has
//                                      int((R - Cr) / dr) >= k;  stab[nbin1 + 1] = largest s with RN(sqrt(s)) <= CR
struct OrientTabs
{
    const double* sedge;
    const float*  fedge;
    const double* stab;
};

// exact angle bin of calc_orient_mof.cpp:41-47,193-194 from dot = vecOut . wall and n2 = |vecOut|^2: FP64 sqrt and divide
__device__ __noinline__ int orient_exact_bin(double dot, double n2, const double* sedge, int nAngle)
{
    const double rv = __dsqrt_rn(n2);
    const double cs = __ddiv_rn(dot, __dmul_rn(rv, 1.0));
    if (!(cs >= -1.0 && cs <= 1.0)) return -1;   // acos -> NaN in the reference (undefined bin): dropped
    int lo = 0, hi = nAngle + 1;                 // invariant: cs <= sedge[lo] (sedge[0] = 2), cs > sedge[hi] (virtual)
    while (hi - lo > 1)
    {
        const int mid = (lo + hi) >> 1;
        if (cs <= sedge[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// One molecule: NCMDens cell (cell_z = ncells + meZ) and orient cell (cell_a = meZ + meA*nbin1), -1 where nothing is
// counted.  Mirrors calc_orient_mof.cpp:156-199 statement by statement.
template <bool SMEM>
__device__ __forceinline__ void orient_molecule(const float* p, int napm, const double* sw, const DivConst& wsum,
                                                const FrameGeom& g, const OrientDev& d, const OrientTabs& tb, int& cell_z,
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
    const double s  = __dadd_rn(__dmul_rn(t1, t1), __dmul_rn(t2, t2));   // the argument of std::sqrt at :175
    // Cr <= R && R <= CR, meZ = int((R - Cr) / dr) with R = RN(sqrt(s)): all monotone in s, so compare s against the
    // host-tabulated thresholds.  An FP32 estimate gives the starting bin, FP64 compares make it exact.
    if (!(s >= tb.stab[0] && s <= tb.stab[d.nbin1 + 1])) return;
    nsel++;
    int meZ = (int)((sqrtf((float)s) - d.Crf) * d.inv_drf);
    meZ     = max(0, min(meZ, d.nbin1));
    while (meZ > 0 && s < tb.stab[meZ]) --meZ;
    while (meZ < d.nbin1 && s >= tb.stab[meZ + 1]) ++meZ;
    if (meZ >= d.nbin1) { ndrop++; return; }
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
    // calcAngle (:41-47): wall = unit vector along DIMNOrient, so |wall| == 1 and the dot product is vo[DIMNOrient]
    // exactly (the other two products are signed zeros); cos = dot / |vo|
    const double n2  = __dadd_rn(__dadd_rn(__dmul_rn(vo[0], vo[0]), __dmul_rn(vo[1], vo[1])), __dmul_rn(vo[2], vo[2]));
    const double dot = d.DIMNOrient == 0 ? vo[0] : (d.DIMNOrient == 1 ? vo[1] : vo[2]);
    // FP32 bracket: |cs_f - cs_ref| < 1e-6 (two conversions, rsqrt.approx, one multiply: < 8 * 2^-24), so a value more
    // than kMargin away from both neighbouring edges (and from +-1, where the reference's acos turns NaN) is decided.
    constexpr float kMargin = 2e-6f;
    const float     csf     = (float)dot * rsqrtf((float)n2);
    int             lo = 0, hi = d.nAngle + 1;
    while (hi - lo > 1)
    {
        const int mid = (lo + hi) >> 1;
        if (csf <= tb.fedge[mid]) lo = mid; else hi = mid;
    }
    const bool proven = (fabsf(csf) < 1.0f - kMargin) && (lo == 0 || csf <= tb.fedge[lo] - kMargin) &&
                        (lo >= d.nAngle || csf >= tb.fedge[lo + 1] + kMargin);
    if (!proven) lo = orient_exact_bin(dot, n2, tb.sedge, d.nAngle);
    if (lo < 0 || lo >= d.nAngle) { ndrop++; return; }       // NaN angle, or angle == 180 (:194-196) -> out of bounds
    cell_a = meZ + lo * d.nbin1;
}

// Heterogeneous molecule sizes (GmxHandleFull groups, reference include/itp/gmx_bits/gmx_handle_full.ipp:115-199): the same
// literal FP64 evaluation per molecule, with the molecule's own atom count, weights (wfull + off[i]) and divisor (div[i]).
// The reference has no orientation program on GmxHandleFull; this completes the accumulator set for such groups.
__global__ void __launch_bounds__(256) k4_orient_csr(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                     const int* __restrict__ index, const int* __restrict__ off, int nmol,
                                                     const double* __restrict__ wfull, const DivConst* __restrict__ div,
                                                     const FrameGeom* __restrict__ geom, OrientDev d,
                                                     const double* __restrict__ cosedge, const float* __restrict__ fedge_g,
                                                     const double* __restrict__ stab_g,
                                                     unsigned long long* __restrict__ hist,
                                                     unsigned long long* __restrict__ stats, int smem_bins)
{
    extern __shared__ double sm4c[];
    double*   sedge = sm4c;                   // nAngle + 1
    double*   stab  = sedge + d.nAngle + 1;   // nbin1 + 2
    float*    fedge = reinterpret_cast<float*>(stab + d.nbin1 + 2);   // nAngle + 1 (+1 pad)
    unsigned* shist = reinterpret_cast<unsigned*>(fedge + ((d.nAngle + 2) & ~1));
    for (int j = threadIdx.x; j <= d.nAngle; j += blockDim.x) { sedge[j] = cosedge[j]; fedge[j] = fedge_g[j]; }
    for (int j = threadIdx.x; j < d.nbin1 + 2; j += blockDim.x) stab[j] = stab_g[j];
    const OrientTabs tb{ sedge, fedge, stab };
    for (int b = threadIdx.x; b < smem_bins; b += blockDim.x) shist[b] = 0;
    __syncthreads();
    const int  ncells   = d.nbin1 * d.nAngle;
    const bool use_smem = smem_bins >= ncells + d.nbin1;
    unsigned   ndrop = 0, nsel = 0;
    const int  i  = blockIdx.x * blockDim.x + threadIdx.x;
    const int  f0 = blockIdx.y * kFramesPerBlock, f1 = min(nframes, f0 + kFramesPerBlock);
    if (i < nmol)
    {
        const int      o0 = __ldg(off + i), na = __ldg(off + i + 1) - o0;
        const size_t   moff = (size_t)__ldg(index + o0) * 3;   // atoms are molN + j (full.ipp:187)
        const DivConst dv = div[i];
        for (int f = f0; f < f1; ++f)
        {
            int cz, ca;
            const FrameGeom g = geom[f];
            orient_molecule<false>(x + (size_t)f * frame_stride + moff, na, wfull + o0, dv, g, d, tb, cz, ca, nsel, ndrop);
            if (cz >= 0) { if (use_smem) atomicAdd(&shist[cz], 1u); else atomicAdd(&hist[cz], 1ull); }
            if (ca >= 0) { if (use_smem) atomicAdd(&shist[ca], 1u); else atomicAdd(&hist[ca], 1ull); }
        }
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

// direct-load kernel (index group not contiguous); the TMA-staged variant is k_stream<2>
__global__ void __launch_bounds__(256) k4_orient(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                 const int* __restrict__ index, int nmol, int napm,
                                                 const double* __restrict__ w, DivConst wsum,
                                                 const FrameGeom* __restrict__ geom, OrientDev d,
                                                 const double* __restrict__ cosedge, const float* __restrict__ fedge_g,
                                                 const double* __restrict__ stab_g,
                                                 unsigned long long* __restrict__ hist,
                                                 unsigned long long* __restrict__ stats, int smem_bins)
{
    extern __shared__ double sm4[];
    double*   sw    = sm4;                    // napm
    double*   sedge = sm4 + napm;             // nAngle + 1
    double*   stab  = sedge + d.nAngle + 1;   // nbin1 + 2
    float*    fedge = reinterpret_cast<float*>(stab + d.nbin1 + 2);   // nAngle + 1 (+1 pad)
    unsigned* shist = reinterpret_cast<unsigned*>(fedge + ((d.nAngle + 2) & ~1));
    for (int j = threadIdx.x; j < napm; j += blockDim.x) sw[j] = w[j];
    for (int j = threadIdx.x; j <= d.nAngle; j += blockDim.x) { sedge[j] = cosedge[j]; fedge[j] = fedge_g[j]; }
    for (int j = threadIdx.x; j < d.nbin1 + 2; j += blockDim.x) stab[j] = stab_g[j];
    const OrientTabs tb{ sedge, fedge, stab };
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
        orient_molecule<false>(x + (size_t)f * frame_stride + moff, napm, sw, wsum, g, d, tb, cz, ca, nsel, ndrop);
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
