// FP32-bracketed fast paths for the fused streaming accumulators (K2 density, K4 orientation).
//
// The reference evaluates centres, bins and angles in FP64 (README.md:143-150, src/mofs/calc_orient_mof.cpp:150-200).
// Done literally that costs 200 (K2) to 700 (K4) warp-instructions per molecule, while HBM delivers a molecule's
// 36 bytes every ~200 issue slots: the literal kernels are instruction-bound at 0.2-0.6 of the HBM roofline
// (profiles/r01_hbm.md).  Here every molecule is first evaluated in FP32 together with a rigorous bound on the
// distance between the FP32 value and the reference's FP64 value.  A decision (region filter, bin index, angle bin) is
// PROVEN when the FP32 value is further than that bound from every threshold involved; otherwise the molecule is
// queued and re-evaluated by the literal FP64 code (density_molecule / orient_molecule).  Counts stay bit-exact.
//
// Error model (eps = 2^-24; all quantities per coordinate; D = x_j - x_0 exact, u = D - n L the reference's offset of
// atom j from the molecule's first atom after its unwrap loops, gmx_handle.ipp:139-142,188-191):
//   d  = fl(x_j - x_0)                       |d - D|  <= eps |D|
//   n  = rint(d / L) by the 1.5*2^23 trick, u^ = fl(d - n L) (one FMA rounding)
//   if |u^| < hs2 = (L/2)(1 - 1e-5) then n is the reference's shift count (the only integer with |D - nL| < L/2) and
//                                            |u^ - u| <= eps (|d| + |u^|)(1 + 2 eps)  =: eu
//   centre  c^ = fl(x_0 + fl(invW * sum_j fl(w_j u^_j)))   vs   c = x_0 + sum w_j u_j / W  (the reference's value up to
//   FP64 rounding, < 1e-13 relative, absorbed by the 5 % slack):
//       |c^ - c| <= rho [ eu + (napm + 3) eps m_u ] + eps |c^|,   rho = sum|w_j| / |W|,  m_u = max |u^|, m_d = max |d|
// Everything downstream (bin coordinate, radius, bisector / normal vector, cosine) propagates these bounds with
// first-order terms inflated by >= 5 % and second-order terms bounded explicitly; see each function.
#pragma once

#include "b2a_kernels_orient.cuh"
#include "b2a_packed.cuh"

namespace b2a
{

constexpr float kEps      = 5.9604645e-8f;   // 2^-24
constexpr float kRndMagic = 12582912.0f;     // 1.5 * 2^23: fl(v + magic) - magic == rint(v) for |v| < 2^22
constexpr int   kCosLut   = 4096;            // K4: uniform-in-cos cells -> first candidate angle bin

struct FastParams   // block-uniform, host-computed per launch
{
    int   enabled;    // 0: literal FP64 only; 1: bracket + queue; 2: validation (both, compare)
    float invW;       // fl(1 / W)
    float e1;         // 1.05 eps                       (x |c^|)
    float e2;         // 1.05 eps rho                   (x m_d)
    float e3;         // 1.05 eps rho (napm + 4)        (x m_u)
};

// offset of one coordinate from the reference atom, shifted by whole boxes into (-L/2, L/2); updates the maxima that
// prove the shift (m_u) and bound the error (m_d)
__device__ __forceinline__ float fast_offset(float t, float ref, float L, float invL, float& m_u, float& m_d)
{
    const float d = t - ref;
    const float n = __fadd_rn(__fmaf_rn(d, invL, kRndMagic), -kRndMagic);
    const float u = __fmaf_rn(-n, L, d);
    m_u = fmaxf(m_u, fabsf(u));
    m_d = fmaxf(m_d, fabsf(d));
    return u;
}

// ---- one centre component -------------------------------------------------------------------------------------------
// returns c^_k; m_u / m_d accumulate over the atoms of the molecule, hs_ok is cleared when a shift is not proven
template <bool SMEM>
__device__ __forceinline__ float fast_center_component(const float* p, int k, int napm, const float* swf, float invW, float L,
                                                       float invL, float hs2, float& m_u, float& m_d, bool& hs_ok)
{
    const float ref = ldf<SMEM>(p + k);
    float       acc = 0.f, mu = 0.f;
    for (int j = 1; j < napm; ++j)
    {
        const float u = fast_offset(ldf<SMEM>(p + 3 * j + k), ref, L, invL, mu, m_d);
        acc = __fmaf_rn(swf[j], u, acc);
    }
    hs_ok = hs_ok && (mu < hs2);
    m_u   = fmaxf(m_u, mu);
    return __fmaf_rn(acc, invW, ref);
}

__device__ __forceinline__ float fast_center_err(float c, float m_u, float m_d, const FastParams& fp)
{
    return __fmaf_rn(fabsf(c), fp.e1, __fmaf_rn(m_d, fp.e2, m_u * fp.e3));
}

// three-way test of lo <= c <= hi (or c < hi: identical away from the edges): 1 proven inside, 0 proven outside, -1 undecided
__device__ __forceinline__ int fast_in_range(float c, float E, float lo, float hi)
{
    if (c - E > lo && c + E < hi) return 1;
    if (c + E < lo || c - E > hi) return 0;
    return -1;
}

// ---- K2 fast path ---------------------------------------------------------------------------------------------------
struct DensityFast
{
    float low, up, low2, up2;   // the program's own float bounds (exact)
    float inv_dbin;             // fl(1 / dbin)
};

// returns: >= 0 proven bin; -1 proven "not counted" (filtered out, or outside the array: ndrop/nsel updated like the
// literal code); -2 undecided -> caller queues the molecule for density_molecule()
template <bool SMEM>
__device__ __forceinline__ int density_fast(const float* p, int napm, const float* swf, const FastParams& fp, const FastFrame& ff,
                                            const DensityDev& d, const DensityFast& df, unsigned& nsel, unsigned& ndrop)
{
    float m_u = 0.f, m_d = 0.f;
    bool  ok  = true;
    const float c = fast_center_component<SMEM>(p, d.dim, napm, swf, fp.invW, ff.L[d.dim], ff.invL[d.dim], ff.hs2[d.dim], m_u, m_d, ok);
    float       c2 = 0.f;
    if (d.dim2 >= 0)
        c2 = fast_center_component<SMEM>(p, d.dim2, napm, swf, fp.invW, ff.L[d.dim2], ff.invL[d.dim2], ff.hs2[d.dim2], m_u, m_d, ok);
    if (!ok) return -2;
    const float E  = fast_center_err(c, m_u, m_d, fp);
    const int   in = fast_in_range(c, E, df.low, df.up);
    if (in == 0) return -1;
    if (in < 0) return -2;
    if (d.dim2 >= 0)
    {
        const float E2  = fast_center_err(c2, m_u, m_d, fp);
        const int   in2 = fast_in_range(c2, E2, df.low2, df.up2);
        if (in2 == 0) return -1;
        if (in2 < 0) return -2;
    }
    // me = int((c - low) / dbin): q^ = fl(fl(c^ - low) * inv_dbin); |q^ - q| <= E inv_dbin (1 + 3 eps) + 3 eps q
    const float q  = (c - df.low) * df.inv_dbin;
    const float Eq = __fmaf_rn(E, df.inv_dbin * 1.01f, q * (3.2f * kEps));
    const float k  = floorf(q);
    const float fr = q - k;
    if (!(fr > Eq && fr < 1.0f - Eq) || !(q < 2.0e9f)) return -2;
    nsel++;
    const int me = (int)k;
    if (me >= d.nbin) { ndrop++; return -1; }
    return me;
}

// ---- K4 fast path ---------------------------------------------------------------------------------------------------
struct OrientFast
{
    float low, up, c1, c2, Cr, CR, inv_dr;   // the program's float parameters (exact) and fl(1 / dr)
    float cr_lo;                             // Cr when Cr > 0, else -huge: `Cr <= R` is then always true (R >= 0)
};

// returns 1: proven, cell_z / cell_a / nsel / ndrop set exactly like orient_molecule(); 0: undecided (nothing touched).
// Written branch-light: every "cannot decide" condition is accumulated into one predicate and tested once; only the two
// proven-outside cases (slab filter, radial selection) leave early.
template <bool SMEM>
__device__ __forceinline__ int orient_fast(const float* p, int napm, const float* swf, const FastParams& fp, const FastFrame& ff,
                                           const OrientDev& d, const OrientFast& of, const float* fedge, const unsigned char* lut,
                                           int& cell_z, int& cell_a, unsigned& nsel, unsigned& ndrop)
{
    const float r0 = ldf<SMEM>(p), r1 = ldf<SMEM>(p + 1), r2 = ldf<SMEM>(p + 2);
    const float L0 = ff.L[0], L1 = ff.L[1], L2 = ff.L[2], iL0 = ff.invL[0], iL1 = ff.invL[1], iL2 = ff.invL[2];
    float       a0 = 0.f, a1 = 0.f, a2 = 0.f, mu0 = 0.f, mu1 = 0.f, mu2 = 0.f, m_d = 0.f;
#pragma unroll 2
    for (int j = 1; j < napm; ++j)
    {
        const float u0 = fast_offset(ldf<SMEM>(p + 3 * j), r0, L0, iL0, mu0, m_d);
        const float u1 = fast_offset(ldf<SMEM>(p + 3 * j + 1), r1, L1, iL1, mu1, m_d);
        const float u2 = fast_offset(ldf<SMEM>(p + 3 * j + 2), r2, L2, iL2, mu2, m_d);
        const float w  = swf[j];
        a0 = __fmaf_rn(w, u0, a0); a1 = __fmaf_rn(w, u1, a1); a2 = __fmaf_rn(w, u2, a2);
    }
    // vectors p1 - p2, p1 - p3 (:31-38).  The reference subtracts the unwrapped atoms and applies periodicity(): the
    // result is the representative of x_tp1 - x_tp2 (mod L) in [-L/2, L/2], which fast_offset() computes directly; it is
    // proven (and unique) when |v^| < hs2, and then |v^ - v| <= eps (|d| + |v^|)
    float v1[3], v2[3], mv0 = 0.f, mv1 = 0.f, mv2 = 0.f, m_dv = 0.f;
    {
        const float* q1 = p + 3 * d.tp1; const float* q2 = p + 3 * d.tp2; const float* q3 = p + 3 * d.tp3;
        const float  x0 = ldf<SMEM>(q1), x1 = ldf<SMEM>(q1 + 1), x2 = ldf<SMEM>(q1 + 2);
        v1[0] = fast_offset(x0, ldf<SMEM>(q2), L0, iL0, mv0, m_dv);
        v1[1] = fast_offset(x1, ldf<SMEM>(q2 + 1), L1, iL1, mv1, m_dv);
        v1[2] = fast_offset(x2, ldf<SMEM>(q2 + 2), L2, iL2, mv2, m_dv);
        v2[0] = fast_offset(x0, ldf<SMEM>(q3), L0, iL0, mv0, m_dv);
        v2[1] = fast_offset(x1, ldf<SMEM>(q3 + 1), L1, iL1, mv1, m_dv);
        v2[2] = fast_offset(x2, ldf<SMEM>(q3 + 2), L2, iL2, mv2, m_dv);
    }
    const float h0 = ff.hs2[0], h1 = ff.hs2[1], h2 = ff.hs2[2];
    bool        und = !(fmaxf(mu0, mv0) < h0 && fmaxf(mu1, mv1) < h1 && fmaxf(mu2, mv2) < h2);   // a shift count is not proven
    const float m_u = fmaxf(mu0, fmaxf(mu1, mu2));
    const float c0 = __fmaf_rn(a0, fp.invW, r0), c1 = __fmaf_rn(a1, fp.invW, r1), c2 = __fmaf_rn(a2, fp.invW, r2);
    // one error bound for all three centre components (|c| replaced by the largest of them)
    const float Ec = fast_center_err(fmaxf(fabsf(c0), fmaxf(fabsf(c1), fabsf(c2))), m_u, m_d, fp);
    // slab filter lowPos <= posc(i, DIMN) <= upPos (:158)
    const float cz = d.DIMN == 0 ? c0 : (d.DIMN == 1 ? c1 : c2);
    if (!und && (cz + Ec < of.low || cz - Ec > of.up)) { cell_z = cell_a = -1; return 1; }   // proven outside
    und = und || !(cz - Ec > of.low && cz + Ec < of.up);
    // R = sqrt(t1^2 + t2^2) (:160-175); |R^ - R| <= sqrt2 (Ec + eps |t|) + 4 eps R (sqrt.approx: 2^-23 relative)  <  1.5 Ec + 8 eps R
    const float ca = d.DIMN == 0 ? c1 : c0, cb = d.DIMN == 2 ? c1 : c2;
    const float ta = ca - of.c1, tb = cb - of.c2;
    float       R;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(R) : "f"(__fmaf_rn(ta, ta, tb * tb)));
    const float ER = __fmaf_rn(R, 8.0f * kEps, 1.5f * Ec);
    if (!und && (R + ER < of.cr_lo || R - ER > of.CR)) { cell_z = cell_a = -1; return 1; }   // proven not selected
    und = und || !(R - ER > of.cr_lo && R + ER < of.CR);
    const float q  = (R - of.Cr) * of.inv_dr;
    const float Eq = __fmaf_rn(ER, of.inv_dr * 1.01f, fabsf(q) * (3.2f * kEps));
    const float k  = floorf(q);
    const float fr = q - k;
    und = und || !(fr > Eq && fr < 1.0f - Eq && q < 2.0e9f && q >= 0.f);
    const float m_v = fmaxf(mv0, fmaxf(mv1, mv2));
    const float Ev  = (1.05f * kEps) * (m_dv + m_v);
    float vo0, vo1, vo2, Evo;
    if (d.method == 2)
    {
        vo0 = __fmaf_rn(v1[1], v2[2], -(v1[2] * v2[1]));
        vo1 = __fmaf_rn(v1[2], v2[0], -(v1[0] * v2[2]));
        vo2 = __fmaf_rn(v1[0], v2[1], -(v1[1] * v2[0]));
        // |d(ab - cd)| <= Ev (|a|+|b|+|c|+|d|) + 2 Ev^2 + 2 eps (|ab| + |cd|)
        Evo = 1.05f * (4.0f * m_v * Ev + 2.0f * Ev * Ev + 4.0f * kEps * m_v * m_v);
    }
    else
    {
        vo0 = v1[0] + v2[0]; vo1 = v1[1] + v2[1]; vo2 = v1[2] + v2[2];
        Evo = __fmaf_rn(2.1f * kEps, m_v, 2.1f * Ev);
    }
    // cos = vo[DIMNOrient] / |vo| (:41-47 with the unit wall vector): |dcos| <= (1 + sqrt3) Evo / |vo| + rounding
    const float n2   = __fmaf_rn(vo2, vo2, __fmaf_rn(vo1, vo1, vo0 * vo0));
    float       rinv;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(n2));
    const float dot = d.DIMNOrient == 0 ? vo0 : (d.DIMNOrient == 1 ? vo1 : vo2);
    const float cs  = dot * rinv;
    const float rel = Evo * rinv;
    const float Ecs = __fmaf_rn(rel, 3.0f, 10.0f * kEps + 1.3e-7f);   // + the float rounding of the tabulated edges
    // |vo| must be well above its error and far from denormal (ftz), cos away from +-1 where the reference's acos may be NaN
    und = und || !(rel < 0.05f && n2 > 1.0e-30f && fabsf(cs) < 1.0f - Ecs);
    // angle bin: start from the first candidate of the cos cell, walk down the (decreasing) edge table
    int ci = (int)((cs + 1.0f) * (0.5f * kCosLut));
    ci     = min(max(ci, 0), kCosLut - 1);
    int lo = lut[ci];
    while (lo < d.nAngle && cs <= fedge[lo + 1]) ++lo;
    const float e_up = fedge[lo], e_dn = fedge[min(lo + 1, d.nAngle)];
    und = und || !((lo == 0 || cs <= e_up - Ecs) && (lo >= d.nAngle || cs >= e_dn + Ecs));
    if (und) return 0;
    // proven: account exactly like orient_molecule()
    const int meZ = (int)k;
    cell_z = cell_a = -1;
    nsel++;
    if (meZ >= d.nbin1) { ndrop++; return 1; }
    cell_z = d.nbin1 * d.nAngle + meZ;
    if (lo >= d.nAngle) { ndrop++; return 1; }
    cell_a = meZ + lo * d.nbin1;
    return 1;
}

// ---- K4 fast path for three-site molecules whose first vector atom is the molecule's reference atom -----------------------
// (water: tp1 = O, {tp2, tp3} = the two H).  Same bracket as orient_fast(), restructured around what the instruction
// stream of that function showed (profiles/r01_hbm.md: ~300 instructions per molecule, twice the HBM-rate budget):
//  * With tp1 the reference atom, the vectors are the unwrapped offsets themselves: x~_0 - x~_j = -u_j, and periodicity()
//    leaves them alone because the reference's unwrap loops already made |u_j| <= L/2 (gmx_handle.ipp:139-142,
//    calc_orient_mof.cpp:31-38).  The six extra minimum-image evaluations disappear; m_v = m_u and m_dv = m_d.
//  * The offsets of atoms 1 and 2 are evaluated together, per dimension, in packed FP32 (FADD2 / FFMA2): 12 issue slots
//    instead of 24.  Each lane is bit-identical to fast_offset().
//  * One threshold proves all six shift counts (min_k hs2[k], FastPack::hsmin) and one bounds all raw differences
//    (FastPack::dmax); the maxima are FMNMX3 trees over |.| operands.
//  * The angle bin is found without a loop: the look-up cell names the first candidate bin, one compare against the next
//    edge moves up at most one bin, and the margin test against the two enclosing edges doubles as the proof that no
//    further step was needed (fedge[0] = 2 and fedge[nAngle + 1 ..] = -3 make the end bins pass without special cases).
// `wsgn` is the sign that turns the computed vector into the reference's vecOut: method 1: -(u1 + u2) -> -1;
// method 2: (-u_tp2) x (-u_tp3) = +-(u1 x u2) -> +1 for (tp2, tp3) = (2nd, 3rd atom), -1 when swapped.
template <bool SMEM>
__device__ __forceinline__ int orient_fast3(const float* p, float w1, float w2, float wsgn, const FastParams& fp, const FastPack& fk,
                                            const OrientDev& d, const OrientFast& of, const float* fedge, const unsigned char* lut,
                                            int& cell_z, int& cell_a, unsigned& nsel, unsigned& ndrop)
{
    const float x0[3] = { ldf<SMEM>(p), ldf<SMEM>(p + 1), ldf<SMEM>(p + 2) };
    const unsigned long long magic2 = f2pack(kRndMagic, kRndMagic);
    float u1[3], u2[3], m_u = 0.f, m_d = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const unsigned long long nL2 = *reinterpret_cast<const unsigned long long*>(&fk.negL2[2 * k]);
        const unsigned long long iL2 = *reinterpret_cast<const unsigned long long*>(&fk.invL2[2 * k]);
        const unsigned long long X   = f2pack(ldf<SMEM>(p + 3 + k), ldf<SMEM>(p + 6 + k));
        const unsigned long long D   = f2sub(X, f2pack(x0[k], x0[k]));            // d = fl(x_j - x_0)
        const unsigned long long N   = f2sub(f2fma(D, iL2, magic2), magic2);     // rint(d / L)
        const unsigned long long U   = f2fma(N, nL2, D);                         // u^ = fl(d - n L)
        float da, db;
        f2unpack(D, da, db);
        f2unpack(U, u1[k], u2[k]);
        m_d = fmaxf(fmaxf(m_d, fabsf(da)), fabsf(db));
        m_u = fmaxf(fmaxf(m_u, fabsf(u1[k])), fabsf(u2[k]));
    }
    bool und = !(m_u < fk.hsmin && m_d <= fk.dmax);   // a shift count is not proven / outside the error model
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = __fmaf_rn(__fmaf_rn(w2, u2[k], w1 * u1[k]), fp.invW, x0[k]);
    const float Ec = fast_center_err(fmaxf(fabsf(c[0]), fmaxf(fabsf(c[1]), fabsf(c[2]))), m_u, m_d, fp);
    // slab filter lowPos <= posc(i, DIMN) <= upPos (:158)
    const float cz = d.DIMN == 0 ? c[0] : (d.DIMN == 1 ? c[1] : c[2]);
    if (!(cz - Ec > of.low && cz + Ec < of.up))       // not proven inside: proven outside, or undecided
    {
        if (!und && (cz + Ec < of.low || cz - Ec > of.up)) { cell_z = cell_a = -1; return 1; }
        und = true;
    }
    // R = sqrt(t1^2 + t2^2) (:160-175), bounds as in orient_fast()
    const float ca = d.DIMN == 0 ? c[1] : c[0], cb = d.DIMN == 2 ? c[1] : c[2];
    const float ta = ca - of.c1, tb = cb - of.c2;
    float       R;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(R) : "f"(__fmaf_rn(ta, ta, tb * tb)));
    const float ER = __fmaf_rn(R, 8.0f * kEps, 1.5f * Ec);
    if (!(R - ER > of.cr_lo && R + ER < of.CR))       // not proven selected: proven not selected, or undecided
    {
        if (!und && (R + ER < of.cr_lo || R - ER > of.CR)) { cell_z = cell_a = -1; return 1; }
        und = true;
    }
    const float q  = (R - of.Cr) * of.inv_dr;
    const float Eq = __fmaf_rn(ER, of.inv_dr * 1.01f, fabsf(q) * (3.2f * kEps));
    const float kq = floorf(q);
    const float fr = q - kq;
    und = und || !(fr > Eq && fr < 1.0f - Eq && q < 2.0e9f && q >= 0.f);
    // vecOut up to the sign wsgn; |v^ - v| <= Ev per component with m_v = m_u, m_dv = m_d
    const float Ev = (1.05f * kEps) * (m_d + m_u);
    float vo0, vo1, vo2, Evo;
    if (d.method == 2)
    {
        vo0 = __fmaf_rn(u1[1], u2[2], -(u1[2] * u2[1]));
        vo1 = __fmaf_rn(u1[2], u2[0], -(u1[0] * u2[2]));
        vo2 = __fmaf_rn(u1[0], u2[1], -(u1[1] * u2[0]));
        Evo = 1.05f * (4.0f * m_u * Ev + 2.0f * Ev * Ev + 4.0f * kEps * m_u * m_u);
    }
    else
    {
        vo0 = u1[0] + u2[0]; vo1 = u1[1] + u2[1]; vo2 = u1[2] + u2[2];
        Evo = __fmaf_rn(2.1f * kEps, m_u, 2.1f * Ev);
    }
    const float n2 = __fmaf_rn(vo2, vo2, __fmaf_rn(vo1, vo1, vo0 * vo0));
    float       rinv;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(n2));
    const float dot = d.DIMNOrient == 0 ? vo0 : (d.DIMNOrient == 1 ? vo1 : vo2);
    const float cs  = dot * (rinv * wsgn);
    const float rel = Evo * rinv;
    const float Ecs = __fmaf_rn(rel, 3.0f, 10.0f * kEps + 1.3e-7f);
    und = und || !(rel < 0.05f && n2 > 1.0e-30f && fabsf(cs) < 1.0f - Ecs);
    // angle bin, loop-free (see the header comment); a NaN or out-of-range cs has set `und` above and only needs a valid cell
    const unsigned ci = min((unsigned)__float2int_rz(__fmaf_rn(cs, 0.5f * kCosLut, 0.5f * kCosLut)), (unsigned)(kCosLut - 1));
    int            lo = lut[ci];
    const float    ea = fedge[lo], eb = fedge[lo + 1], ec = fedge[lo + 2];
    const bool     step = cs <= eb;
    lo += step ? 1 : 0;
    const float e_up = step ? eb : ea, e_dn = step ? ec : eb;
    und = und || !(cs <= e_up - Ecs && cs >= e_dn + Ecs);
    if (und) return 0;
    const int meZ = (int)kq;
    cell_z = cell_a = -1;
    nsel++;
    if (meZ >= d.nbin1) { ndrop++; return 1; }
    cell_z = d.nbin1 * d.nAngle + meZ;
    if (lo >= d.nAngle) { ndrop++; return 1; }
    cell_a = meZ + lo * d.nbin1;
    return 1;
}

} // namespace b2a
