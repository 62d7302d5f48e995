// K1 (centres), K1b (unwrapped positions) and K2 (fused centre -> density histogram) for sm_100a.
//
// All three are HBM-bound streams over the float rvec array (12 B/atom).  The arithmetic
// is the reference's, operation for operation, in FP64 with FMA contraction disabled
// (__dmul_rn/__dadd_rn/__ddiv_rn): the reference binary is baseline x86-64 (no FMA), so
// `tmpCenter += tmpPos * w` rounds twice (reference include/itp/gmx_bits/gmx_handle.ipp:193) and the
// result here is bit-identical, not merely within tolerance.  The j-summation order per (molecule, k) is
// the reference's sequential order, which is why this is a thread-per-molecule loop and not a shuffle tree.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace b2a
{

struct FastFrame   // per frame, FP32 view of the box for the bracketed fast paths (b2a_kernels_fast.cuh)
{
    float L[3];      // the float box diagonal itself (Lbox is its widening, so this is exact)
    float invL[3];   // fl(1 / L)
    float hs2[3];    // (L/2)(1 - 1e-5): |offset| below this proves the whole-box shift count
    float pad[3];
};

// The same box, laid out for packed FP32 arithmetic (fma.rn.f32x2 works on register PAIRS): every constant appears twice
// in a row, so one LDS.64 / LDS.128 delivers a ready-made pair.  Used by the three-site orientation fast path.
struct __align__(16) FastPack
{
    float negL2[6];   // {-L0, -L0, -L1, -L1, -L2, -L2}
    float invL2[6];   // {1/L0, 1/L0, ...}
    float hsmin;      // min_k hs2[k]; 0 when the box is too elongated for the one-threshold shift proof (everything undecided)
    float dmax;       // 2 max_k L[k]: raw atom-atom differences beyond this are not handled by the packed path
    float pad[2];
};

struct FrameGeom   // per frame; filled on the host at submit time from the float box diagonal (ipp:72-74)
{
    double L[3];         // Lbox[k] = double(fr->box[k][k])
    double half[3];      // Lbox[k] / 2  (ipp:128, 176)
    double inv_unit[3];  // 2^32 / Lbox[k]: box-fraction fixed point used by the RDF fast path
    float  hsafe[3];     // float slightly BELOW Lbox[k]/2: |x - ref| < hsafe in FP32 proves that no unwrap shift is needed
    float  pad;
    FastFrame ff;
    double pad16;        // keeps `fk` on a 16-byte boundary
    FastPack fk;         // sizeof == 208: the streaming kernels fetch one FrameGeom per stage with a 16-byte-granular bulk copy
};
static_assert(offsetof(FrameGeom, fk) % 16 == 0, "FastPack is read with 128-bit shared loads");
static_assert(sizeof(FrameGeom) % 16 == 0, "FrameGeom is moved with cp.async.bulk");

// Division by a block-invariant divisor, correctly rounded, in 3 FP64 instructions instead of the ~35 of __ddiv_rn:
// with y = RN(1/d) from the host (IEEE division), q0 = RN(x*y), r = x - d*q0 (exact in an FMA), q = RN(q0 + r*y) is the
// correctly rounded quotient (Markstein, "Computation of elementary functions on the IBM RISC System/6000", 1990).
// `fast` is cleared by the host when d or 1/d is not a normal number (e.g. the zero charge sum of a neutral molecule,
// where the reference's x/0 = +-inf/NaN must be reproduced), and huge/tiny/non-finite quotients take the IEEE path.
struct DivConst
{
    double d, inv;
    int    fast, pad;
};

__device__ __forceinline__ double div_const(double x, const DivConst& c)
{
    const double q0 = __dmul_rn(x, c.inv);
    if (c.fast && fabs(q0) > 1e-280 && fabs(q0) < 1e280)
    {
        const double r = __fma_rn(-c.d, q0, x);
        return __fma_rn(r, c.inv, q0);
    }
    return __ddiv_rn(x, c.d);
}

__device__ __forceinline__ double unwrap(double t, double ref, double L, double half)
{
    // ipp:139-142 / 188-191: repeated +-L until within half a box of the molecule's first atom
    while (__dsub_rn(t, ref) > half) t = __dsub_rn(t, L);
    while (__dsub_rn(t, ref) < -half) t = __dadd_rn(t, L);
    return t;
}

// FP32 pre-test: (double)t - (double)ref is exact and differs from the rounded float difference by <= 2^-24 relative,
// so |tf - reff| < hsafe (a float a few ulp below L/2) proves -L/2 < t - ref < L/2: the reference's loops would not run.
__device__ __forceinline__ double unwrap_f(float tf, float reff, double L, double half, float hsafe)
{
    const double t = (double)tf;
    if (fabsf(tf - reff) < hsafe) return t;
    return unwrap(t, (double)reff, L, half);
}

__device__ __forceinline__ int to_fixed(double c, double inv_unit)
{
    // nearest multiple of L/2^32, wrapped mod 2^32: integer differences are then minimum-image for free
    return (int)(unsigned)(unsigned long long)__double2ll_rn(c * inv_unit);
}

template <bool SMEM>
__device__ __forceinline__ float ldf(const float* p) { return SMEM ? *p : __ldg(p); }

// ---- per-molecule centre, GmxHandle::loadPositionCenter (ipp:150-199): sequential j order per component ----------
template <bool SMEM>
__device__ __forceinline__ void mol_centers3(const float* p, int napm, const double* sw, const DivConst& wsum,
                                             const FrameGeom& g, double& c0, double& c1, double& c2)
{
    const float r0 = ldf<SMEM>(p), r1 = ldf<SMEM>(p + 1), r2 = ldf<SMEM>(p + 2);
    c0 = c1 = c2 = 0.0;
    constexpr int U = 4;
    int           j = 0;
    for (; j + U <= napm; j += U)
    {
        float v[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            v[u][0] = ldf<SMEM>(p + 3 * (j + u));
            v[u][1] = ldf<SMEM>(p + 3 * (j + u) + 1);
            v[u][2] = ldf<SMEM>(p + 3 * (j + u) + 2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const double wj = sw[j + u];
            c0 = __dadd_rn(c0, __dmul_rn(unwrap_f(v[u][0], r0, g.L[0], g.half[0], g.hsafe[0]), wj));
            c1 = __dadd_rn(c1, __dmul_rn(unwrap_f(v[u][1], r1, g.L[1], g.half[1], g.hsafe[1]), wj));
            c2 = __dadd_rn(c2, __dmul_rn(unwrap_f(v[u][2], r2, g.L[2], g.half[2], g.hsafe[2]), wj));
        }
    }
    for (; j < napm; ++j)
    {
        const double wj = sw[j];
        c0 = __dadd_rn(c0, __dmul_rn(unwrap_f(ldf<SMEM>(p + 3 * j), r0, g.L[0], g.half[0], g.hsafe[0]), wj));
        c1 = __dadd_rn(c1, __dmul_rn(unwrap_f(ldf<SMEM>(p + 3 * j + 1), r1, g.L[1], g.half[1], g.hsafe[1]), wj));
        c2 = __dadd_rn(c2, __dmul_rn(unwrap_f(ldf<SMEM>(p + 3 * j + 2), r2, g.L[2], g.half[2], g.hsafe[2]), wj));
    }
    c0 = div_const(c0, wsum);
    c1 = div_const(c1, wsum);
    c2 = div_const(c2, wsum);
}

template <bool SMEM>
__device__ __forceinline__ double center_component(const float* p, int k, int napm, const double* sw, const DivConst& wsum,
                                                   double L, double half, float hsafe)
{
    const float ref = ldf<SMEM>(p + k);
    double      c   = 0.0;
    for (int j = 0; j < napm; ++j) c = __dadd_rn(c, __dmul_rn(unwrap_f(ldf<SMEM>(p + 3 * j + k), ref, L, half, hsafe), sw[j]));
    return div_const(c, wsum);
}

// the same value (same operations in the same order), written so that only the additions sit on the dependent chain: the loads,
// conversions, unwrap pre-tests and multiplications of four atoms at a time are independent of the running sum
template <bool SMEM>
__device__ __forceinline__ double center_component_pipelined(const float* p, int k, int napm, const double* sw, const DivConst& wsum,
                                                             double L, double half, float hsafe)
{
    const float ref = ldf<SMEM>(p + k);
    double      c   = 0.0;
    int         j   = 0;
    for (; j + 4 <= napm; j += 4)
    {
        double t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = __dmul_rn(unwrap_f(ldf<SMEM>(p + 3 * (j + u) + k), ref, L, half, hsafe), sw[j + u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) c = __dadd_rn(c, t[u]);
    }
    for (; j < napm; ++j) c = __dadd_rn(c, __dmul_rn(unwrap_f(ldf<SMEM>(p + 3 * j + k), ref, L, half, hsafe), sw[j]));
    return div_const(c, wsum);
}

// ---- K1 (direct loads; used when the index group is not one contiguous run of atoms) -----------------------------
// cen: [frame][k][i] doubles (= itp::matd column-major per frame); fix: [frame][i] int4 {x,y,z fixed, i}
__global__ void __launch_bounds__(256) k1_centers(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                  const int* __restrict__ index, int nmol, int napm,
                                                  const double* __restrict__ w, DivConst wsum,
                                                  const FrameGeom* __restrict__ geom, double* __restrict__ cen,
                                                  int4* __restrict__ fix)
{
    extern __shared__ double sw[];
    for (int j = threadIdx.x; j < napm; j += blockDim.x) sw[j] = w[j];
    __syncthreads();
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmol || f >= nframes) return;
    const FrameGeom g = geom[f];
    const float*    p = x + (size_t)f * frame_stride + (size_t)__ldg(index + (size_t)i * napm) * 3;
    double          c0, c1, c2;
    mol_centers3<false>(p, napm, sw, wsum, g, c0, c1, c2);
    double* o = cen + (size_t)f * 3 * nmol;
    o[i]                    = c0;
    o[(size_t)nmol + i]     = c1;
    o[2 * (size_t)nmol + i] = c2;
    if (fix) fix[(size_t)f * nmol + i] = make_int4(to_fixed(c0, g.inv_unit[0]), to_fixed(c1, g.inv_unit[1]), to_fixed(c2, g.inv_unit[2]), i);
}

// ---- K1 for heterogeneous molecule sizes: GmxHandleFull::loadPositionCenter (reference
// include/itp/gmx_bits/gmx_handle_full.ipp:150-199).  off[i] = first index position of molecule i, wfull[n] = weight of
// the atom at index position n, div[i] = the molecule's own weight total (totMass / totCharge / napm, full.ipp:323-344)
__global__ void __launch_bounds__(256) k1_centers_csr(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                      const int* __restrict__ index, const int* __restrict__ off, int nmol,
                                                      const double* __restrict__ wfull, const DivConst* __restrict__ div,
                                                      const FrameGeom* __restrict__ geom, double* __restrict__ cen,
                                                      int4* __restrict__ fix)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmol || f >= nframes) return;
    const FrameGeom g  = geom[f];
    const int       o0 = __ldg(off + i), na = __ldg(off + i + 1) - o0;
    const float*    p  = x + (size_t)f * frame_stride + (size_t)__ldg(index + o0) * 3;   // atoms are molN + j (full.ipp:187)
    const double*   w  = wfull + o0;
    const float     r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2);
    double          c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int j = 0; j < na; ++j)
    {
        const double wj = __ldg(w + j);
        c0 = __dadd_rn(c0, __dmul_rn(unwrap_f(__ldg(p + 3 * j), r0, g.L[0], g.half[0], g.hsafe[0]), wj));
        c1 = __dadd_rn(c1, __dmul_rn(unwrap_f(__ldg(p + 3 * j + 1), r1, g.L[1], g.half[1], g.hsafe[1]), wj));
        c2 = __dadd_rn(c2, __dmul_rn(unwrap_f(__ldg(p + 3 * j + 2), r2, g.L[2], g.half[2], g.hsafe[2]), wj));
    }
    const DivConst d = div[i];
    c0 = div_const(c0, d); c1 = div_const(c1, d); c2 = div_const(c2, d);
    double* o = cen + (size_t)f * 3 * nmol;
    o[i]                    = c0;
    o[(size_t)nmol + i]     = c1;
    o[2 * (size_t)nmol + i] = c2;
    if (fix) fix[(size_t)f * nmol + i] = make_int4(to_fixed(c0, g.inv_unit[0]), to_fixed(c1, g.inv_unit[1]), to_fixed(c2, g.inv_unit[2]), i);
}

// GmxHandleFull::loadPosition (full.ipp:115-148): pos is boxd(nmol, maxNapm); slots j >= napm[i] are not touched
__global__ void __launch_bounds__(256) k1b_positions_csr(const float* __restrict__ xf, const int* __restrict__ index,
                                                         const int* __restrict__ off, const int* __restrict__ mol_of, int ngx,
                                                         int nmol, FrameGeom g, double* __restrict__ pos)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ngx) return;
    const int    i  = __ldg(mol_of + n), o0 = __ldg(off + i), j = n - o0;
    const float* pr = xf + (size_t)__ldg(index + o0) * 3;
    const float* pa = xf + (size_t)__ldg(index + n) * 3;
    double*      o  = pos + ((size_t)i + (size_t)j * nmol) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = unwrap_f(__ldg(pa + k), __ldg(pr + k), g.L[k], g.half[k], g.hsafe[k]);
}

// ---- K1b: unwrapped per-atom positions, GmxHandle::loadPosition (ipp:115-148) --------------------------------
// one frame; pos in itp::boxd memory order: element (i,j)[k] at (i + j*nmol)*3 + k
__global__ void __launch_bounds__(256) k1b_positions(const float* __restrict__ xf, const int* __restrict__ index,
                                                     int nmol, int napm, FrameGeom g, double* __restrict__ pos)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)nmol * napm) return;
    // thread n handles output slot n = i + j*nmol so the 24-byte stores of a warp are contiguous
    const int    i   = (int)(n % nmol), j = (int)(n / nmol);
    const int    mol = __ldg(index + (size_t)i * napm);
    const int    at  = __ldg(index + (size_t)i * napm + j);   // ipp:137 reads index[grp][n], not molN + j
    const float* pr  = xf + (size_t)mol * 3;
    const float* pa  = xf + (size_t)at * 3;
    double*      o   = pos + (size_t)n * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = unwrap_f(__ldg(pa + k), __ldg(pr + k), g.L[k], g.half[k], g.hsafe[k]);
}

// ---- velocity / force loaders: the same weighted mean and gather WITHOUT unwrapping -------------------------------
// GmxHandle::loadVelocityCenter / loadForceCenter (ipp:228-269, 298-339): atoms molN + j, sequential j order, one
// division by atomCenter.sum() per component.  cen: [frame][k][i] like K1.
__global__ void __launch_bounds__(256) k1_field_centers(const float* __restrict__ v, size_t frame_stride, int nframes,
                                                        const int* __restrict__ index, int nmol, int napm,
                                                        const double* __restrict__ w, DivConst wsum, double* __restrict__ cen)
{
    extern __shared__ double sw[];
    for (int j = threadIdx.x; j < napm; j += blockDim.x) sw[j] = w[j];
    __syncthreads();
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmol || f >= nframes) return;
    const float* p  = v + (size_t)f * frame_stride + (size_t)__ldg(index + (size_t)i * napm) * 3;
    double       c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int j = 0; j < napm; ++j)
    {
        const double wj = sw[j];
        c0 = __dadd_rn(c0, __dmul_rn((double)__ldg(p + 3 * j), wj));
        c1 = __dadd_rn(c1, __dmul_rn((double)__ldg(p + 3 * j + 1), wj));
        c2 = __dadd_rn(c2, __dmul_rn((double)__ldg(p + 3 * j + 2), wj));
    }
    double* o = cen + (size_t)f * 3 * nmol;
    o[i]                    = div_const(c0, wsum);
    o[(size_t)nmol + i]     = div_const(c1, wsum);
    o[2 * (size_t)nmol + i] = div_const(c2, wsum);
}

// GmxHandleFull::loadVelocityCenter (gmx_handle_full.ipp:201-250): per-molecule atom counts and weight totals
__global__ void __launch_bounds__(256) k1_field_centers_csr(const float* __restrict__ v, size_t frame_stride, int nframes,
                                                            const int* __restrict__ index, const int* __restrict__ off, int nmol,
                                                            const double* __restrict__ wfull, const DivConst* __restrict__ div,
                                                            double* __restrict__ cen)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmol || f >= nframes) return;
    const int     o0 = __ldg(off + i), na = __ldg(off + i + 1) - o0;
    const float*  p  = v + (size_t)f * frame_stride + (size_t)__ldg(index + o0) * 3;
    const double* w  = wfull + o0;
    double        c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int j = 0; j < na; ++j)
    {
        const double wj = __ldg(w + j);
        c0 = __dadd_rn(c0, __dmul_rn((double)__ldg(p + 3 * j), wj));
        c1 = __dadd_rn(c1, __dmul_rn((double)__ldg(p + 3 * j + 1), wj));
        c2 = __dadd_rn(c2, __dmul_rn((double)__ldg(p + 3 * j + 2), wj));
    }
    const DivConst d = div[i];
    double*        o = cen + (size_t)f * 3 * nmol;
    o[i]                    = div_const(c0, d);
    o[(size_t)nmol + i]     = div_const(c1, d);
    o[2 * (size_t)nmol + i] = div_const(c2, d);
}

// GmxHandle::loadVelocity / loadForce (ipp:201-226, 271-296): vel(i,j)[k] = fr->v[index[grp][n]][k], boxd memory order
__global__ void __launch_bounds__(256) k1b_field_values(const float* __restrict__ vf, const int* __restrict__ index, int nmol,
                                                        int napm, double* __restrict__ out)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)nmol * napm) return;
    const int    i  = (int)(n % nmol), j = (int)(n / nmol);
    const float* pa = vf + (size_t)__ldg(index + (size_t)i * napm + j) * 3;
    double*      o  = out + (size_t)n * 3;
    o[0] = (double)__ldg(pa); o[1] = (double)__ldg(pa + 1); o[2] = (double)__ldg(pa + 2);
}

// ---- K2: centre of the needed dimension(s) -> 1-D histogram, never materialising the centres ----------------
// README.md:143-150, general/calc_density_region.cpp:59-71, general/density_slit_jhl.cpp:59-77
struct DensityDev
{
    int      dim, dim2, nbin, upper_inclusive;
    double   low, up, low2, up2;   // the program's float `real` bounds widened (comparison is done in double)
    DivConst dbin;
};

// returns the bin of this molecule, or -1 (filtered out, or outside the array: counted in ndrop)
template <bool SMEM>
__device__ __forceinline__ int density_molecule(const float* p, int napm, const double* sw, const DivConst& wsum,
                                                const FrameGeom& g, const DensityDev& d, unsigned& nsel, unsigned& ndrop)
{
    const double c  = center_component<SMEM>(p, d.dim, napm, sw, wsum, g.L[d.dim], g.half[d.dim], g.hsafe[d.dim]);
    bool         ok = (d.low <= c) && (d.upper_inclusive ? (c <= d.up) : (c < d.up));
    if (ok && d.dim2 >= 0)
    {
        const double c2 = center_component<SMEM>(p, d.dim2, napm, sw, wsum, g.L[d.dim2], g.half[d.dim2], g.hsafe[d.dim2]);
        ok              = (d.low2 <= c2) && (d.upper_inclusive ? (c2 <= d.up2) : (c2 < d.up2));
    }
    if (!ok) return -1;
    nsel++;
    const int me = __double2int_rz(div_const(__dsub_rn(c, d.low), d.dbin));
    if (me < 0 || me >= d.nbin) { ndrop++; return -1; }
    return me;
}

// density from already materialised centres (heterogeneous groups: K1 csr writes the centres, this bins them)
__global__ void __launch_bounds__(256) k2_density_from_centers(const double* __restrict__ cen, int nframes, int nmol, DensityDev d,
                                                               unsigned long long* __restrict__ hist,
                                                               unsigned long long* __restrict__ stats)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmol || f >= nframes) return;
    const double* c0 = cen + (size_t)f * 3 * nmol;
    const double  c  = c0[(size_t)d.dim * nmol + i];
    bool          ok = (d.low <= c) && (d.upper_inclusive ? (c <= d.up) : (c < d.up));
    if (ok && d.dim2 >= 0)
    {
        const double c2 = c0[(size_t)d.dim2 * nmol + i];
        ok              = (d.low2 <= c2) && (d.upper_inclusive ? (c2 <= d.up2) : (c2 < d.up2));
    }
    if (!ok) return;
    atomicAdd(&stats[4], 1ull);
    const int me = __double2int_rz(div_const(__dsub_rn(c, d.low), d.dbin));
    if (me < 0 || me >= d.nbin) { atomicAdd(&stats[1], 1ull); return; }
    atomicAdd(&hist[me], 1ull);
}

// ---- general region counter / histogram over materialised centres (b2a_region_spec) -----------------------------------------
struct RegionDev
{
    int    nfilter, fdim[3], fincl[3], wrap, cyl, nbdim, bdim[2], nb[2];
    float  flow[3], fup[3];
    double cx, cy, R2, blow[2], dbin[2];
};

__global__ void __launch_bounds__(256) k2g_region(const double* __restrict__ cen, int nframes, int nmol, RegionDev d,
                                                  const FrameGeom* __restrict__ geom, unsigned long long* __restrict__ hist,
                                                  unsigned long long* __restrict__ stats, unsigned long long* __restrict__ perframe,
                                                  size_t ncounters)
{
    const int  f  = blockIdx.y;
    const int  i  = blockIdx.x * blockDim.x + threadIdx.x;
    bool       ok = i < nmol && f < nframes;
    long long  cell = 0;
    bool       oob  = false;
    if (ok)
    {
        const double* c0 = cen + (size_t)f * 3 * nmol;
        for (int q = 0; q < d.nfilter && ok; ++q)
        {
            const double p = c0[(size_t)d.fdim[q] * nmol + i];
            if (d.wrap)
            {
                // calc_numdens_pore.cpp:15-22: `real tmp` wrapped with the double box length, bounds compared in float
                const double L = geom[f].L[d.fdim[q]];
                float        t = (float)p;
                while ((double)t > L) t = (float)__dsub_rn((double)t, L);
                while (t < 0.f) t = (float)__dadd_rn((double)t, L);
                ok = (d.flow[q] <= t) && (d.fincl[q] ? (t <= d.fup[q]) : (t < d.fup[q]));
            }
            else
                ok = ((double)d.flow[q] <= p) && (d.fincl[q] ? (p <= (double)d.fup[q]) : (p < (double)d.fup[q]));
        }
        if (ok && d.cyl)
        {
            const double dx = __dsub_rn(c0[i], d.cx), dy = __dsub_rn(c0[(size_t)nmol + i], d.cy);
            ok              = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= d.R2;     // std::pow(v, 2) is v * v, correctly rounded
        }
        if (ok)
        {
            long long stride = 1;
            for (int b = 0; b < d.nbdim; ++b)
            {
                double p;
                if (d.bdim[b] == 3)
                {   // radial bin: std::sqrt(pow(x - cx, 2) + pow(y - cy, 2)) (calc_radius_distribution_in_pore.cpp:47-49)
                    const double dx = __dsub_rn(c0[i], d.cx), dy = __dsub_rn(c0[(size_t)nmol + i], d.cy);
                    p               = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                }
                else p = c0[(size_t)d.bdim[b] * nmol + i];
                const double q  = d.bdim[b] == 3 && d.blow[b] == 0.0 ? __ddiv_rn(p, d.dbin[b]) : __ddiv_rn(__dsub_rn(p, d.blow[b]), d.dbin[b]);
                const int    me = (q >= -2.0e9 && q <= 2.0e9) ? __double2int_rz(q) : -1;
                if (me < 0 || me >= d.nb[b]) oob = true;     // the reference indexes outside its array here
                cell += (long long)me * stride;
                stride *= d.nb[b];
            }
        }
    }
    // one counter (nbdim == 0): count per warp, one atomic per warp
    if (d.nbdim == 0)
    {
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && m)
        {
            atomicAdd(&hist[0], (unsigned long long)__popc(m));
            atomicAdd(&stats[4], (unsigned long long)__popc(m));
            if (perframe) atomicAdd(&perframe[(size_t)f], (unsigned long long)__popc(m));
        }
        return;
    }
    if (!ok) return;
    atomicAdd(&stats[4], 1ull);
    if (oob) { atomicAdd(&stats[1], 1ull); return; }
    atomicAdd(&hist[cell], 1ull);
    if (perframe) atomicAdd(&perframe[(size_t)f * ncounters + cell], 1ull);
}

// grid: x = molecule block, y = frame group.  A thread owns ONE molecule and walks the frames of its group.
constexpr int kFramesPerBlock = 8;

__global__ void __launch_bounds__(256) k2_density(const float* __restrict__ x, size_t frame_stride, int nframes,
                                                  const int* __restrict__ index, int nmol, int napm,
                                                  const double* __restrict__ w, DivConst wsum,
                                                  const FrameGeom* __restrict__ geom, DensityDev d,
                                                  unsigned long long* __restrict__ hist,
                                                  unsigned long long* __restrict__ stats, int smem_bins)
{
    extern __shared__ double sm2[];
    double*   sw    = sm2;
    unsigned* shist = reinterpret_cast<unsigned*>(sm2 + napm);
    for (int j = threadIdx.x; j < napm; j += blockDim.x) sw[j] = w[j];
    for (int b = threadIdx.x; b < smem_bins; b += blockDim.x) shist[b] = 0;
    __syncthreads();
    const bool use_smem = smem_bins >= d.nbin;
    unsigned   ndrop = 0, nsel = 0;
    const int  i  = blockIdx.x * blockDim.x + threadIdx.x;
    const int  f0 = blockIdx.y * kFramesPerBlock, f1 = min(nframes, f0 + kFramesPerBlock);
    if (i < nmol)
    {
        const size_t moff = (size_t)__ldg(index + (size_t)i * napm) * 3;
        for (int f = f0; f < f1; ++f)
        {
            const int me = density_molecule<false>(x + (size_t)f * frame_stride + moff, napm, sw, wsum, geom[f], d, nsel, ndrop);
            if (me < 0) continue;
            if (use_smem) atomicAdd(&shist[me], 1u);
            else atomicAdd(&hist[me], 1ull);
        }
    }
    __syncthreads();
    if (use_smem)
        for (int b = threadIdx.x; b < d.nbin; b += blockDim.x)
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
