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
// stream of that function showed (profiles/r01_hbm.md: ~300 instructions per molecule, against the ~200 issue slots per
// molecule that HBM leaves):
//  * With tp1 the reference atom, the vectors are the unwrapped offsets themselves: x~_0 - x~_j = -u_j, and periodicity()
//    leaves them alone because the reference's unwrap loops already made |u_j| <= L/2 (gmx_handle.ipp:139-142,
//    calc_orient_mof.cpp:31-38).  The six extra minimum-image evaluations disappear; m_v = m_u and m_dv = m_d.
//  * TWO molecules per thread, one in each half of a 64-bit register pair: every add / multiply / fma of the chain (offsets,
//    centre, error bounds, radius, bin coordinate, vector, cosine) is one packed FP32 instruction (FADD2 / FMUL2 / FFMA2)
//    for both molecules.  Each half is the IEEE operation the scalar code would execute.  Only what has no packed form
//    stays per molecule: |.|-maxima (FMNMX3), compares, MUFU.SQRT / RSQ, float -> int, table look-ups.
//  * One threshold proves all six shift counts (min_k hs2[k], FastPack::hsmin) and one bounds all raw differences
//    (FastPack::dmax).
//  * The fraction test of the radial bin and the angle-bin test are |value - midpoint| < halfwidth - error: one compare
//    instead of two.  The angle bin is found without a
//    loop: the look-up cell names the first candidate bin, one compare against the next edge moves up at most one bin, and
//    the interval test against that bin's {midpoint, halfwidth} (host table, halfwidth shrunk by the float roundings of
//    the table) doubles as the proof that no further step was needed.
// `wsgn` is the sign that turns the computed vector into the reference's vecOut: method 1: -(u1 + u2) -> -1;
// method 2: (-u_tp2) x (-u_tp3) = +-(u1 x u2) -> +1 for (tp2, tp3) = (2nd, 3rd atom), -1 when swapped.
__device__ __forceinline__ unsigned long long f2bc(float a) { return f2pack(a, a); }

// the inputs of orient_fast3x2(): coordinates of two molecules (A in the low, B in the high half of every pair) and the packed
// frame constants.  (Handing the stage back to the producer right after these loads, before the arithmetic, was tried and
// is slower: 0.097 vs 0.090 ms per 64 frames -- the kernel waits on its own dependency chains, not on the feed.)
struct Mol3x2
{
    unsigned long long X0[3], X1[3], X2[3];   // atom 0 (reference), 1, 2; component k
    unsigned long long nL2[3], iL2[3];        // (-L_k, -L_k), (1/L_k, 1/L_k)
    float              hsmin, dmax;
};

template <bool SMEM>
__device__ __forceinline__ void load3x2(Mol3x2& m, const float* pA, const float* pB, const FastPack& fk)
{
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        m.nL2[k] = *reinterpret_cast<const unsigned long long*>(&fk.negL2[2 * k]);
        m.iL2[k] = *reinterpret_cast<const unsigned long long*>(&fk.invL2[2 * k]);
        m.X0[k]  = f2pack(ldf<SMEM>(pA + k), ldf<SMEM>(pB + k));
        m.X1[k]  = f2pack(ldf<SMEM>(pA + 3 + k), ldf<SMEM>(pB + 3 + k));
        m.X2[k]  = f2pack(ldf<SMEM>(pA + 6 + k), ldf<SMEM>(pB + 6 + k));
    }
    m.hsmin = fk.hsmin; m.dmax = fk.dmax;
}

// ---- the part of the two-molecule bracket that follows the centres and the two vectors -----------------------------------------
// C: centre components; MU / MD: per-molecule maxima of |offset| and |raw difference| behind the centre (error of the centre);
// V1, V2: the two vectors up to the sign vsgn; MV / MDV: the maxima behind them (error of the vectors)
// SPEC = 1: the program's default axes and method (calc_orient_mof.cpp:86-92: DIMN = 2, DIMNOrient = 2, method = 1) as compile-time
// constants -- the component selects and the cross-product branch fold away; 0: whatever the spec says, at run time.
template <int SPEC = 0>
__device__ __forceinline__ void orient_tail_x2(const unsigned long long C[3], unsigned long long MU, unsigned long long MD,
                                               const unsigned long long V1[3], const unsigned long long V2[3], unsigned long long MV,
                                               unsigned long long MDV, bool undA, bool undB, bool validA, bool validB, float vsgn,
                                               const FastParams& fp, const OrientDev& d, const OrientFast& of, const float* fedge,
                                               const float2* fmh, const unsigned char* lut, int res[2], int meZ[2], int lo[2])
{
    typedef unsigned long long P;
    float c0a, c0b, c1a, c1b, c2a, c2b;
    f2unpack(C[0], c0a, c0b); f2unpack(C[1], c1a, c1b); f2unpack(C[2], c2a, c2b);
    const P CM = f2pack(fmaxf(fabsf(c0a), fmaxf(fabsf(c1a), fabsf(c2a))), fmaxf(fabsf(c0b), fmaxf(fabsf(c1b), fabsf(c2b))));
    const P Ec = f2fma(CM, f2bc(fp.e1), f2fma(MD, f2bc(fp.e2), f2mul(MU, f2bc(fp.e3))));      // fast_center_err()
    // slab filter (:158) and radial selection (:160-177); DIMN picks the components
    const int DIMN = SPEC ? 2 : d.DIMN, DIMNOrient = SPEC ? 2 : d.DIMNOrient, method = SPEC ? 1 : d.method;
    const P Cz = DIMN == 0 ? C[0] : (DIMN == 1 ? C[1] : C[2]);
    const P Ca = DIMN == 0 ? C[1] : C[0], Cb = DIMN == 2 ? C[1] : C[2];
    bool    outA = false, outB = false;   // proven "not counted"
    {
        const P LO = f2sub(Cz, Ec), HI = f2add(Cz, Ec);
        float   la, lb, ha, hb;
        f2unpack(LO, la, lb); f2unpack(HI, ha, hb);
        // inside: proven; definitely outside (and the shifts proven): "not counted"; anything else: undecided.  Written as predicate
        // algebra (no branches: the two molecules of a thread rarely agree, and a divergent region costs more than the logic)
        const bool inA = la > of.low && ha < of.up, doA = ha < of.low || la > of.up;
        const bool inB = lb > of.low && hb < of.up, doB = hb < of.low || lb > of.up;
        outA = !inA & !undA & doA; undA = undA | (!inA & !doA);
        outB = !inB & !undB & doB; undB = undB | (!inB & !doB);
    }
    const P Ta = f2sub(Ca, f2bc(of.c1)), Tb = f2sub(Cb, f2bc(of.c2));
    const P S  = f2fma(Ta, Ta, f2mul(Tb, Tb));
    float   sa, sb, ra, rb;
    f2unpack(S, sa, sb);
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(sa));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(sb));
    const P R  = f2pack(ra, rb);
    const P ER = f2fma(R, f2bc(8.0f * kEps), f2mul(Ec, f2bc(1.5f)));    // |R^ - R| < 1.5 Ec + 8 eps R
    {
        const P LO = f2sub(R, ER), HI = f2add(R, ER);
        float   la, lb, ha, hb;
        f2unpack(LO, la, lb); f2unpack(HI, ha, hb);
        const bool inA = la > of.cr_lo && ha < of.CR, doA = ha < of.cr_lo || la > of.CR;
        const bool inB = lb > of.cr_lo && hb < of.CR, doB = hb < of.cr_lo || lb > of.CR;
        outA = outA | (!inA & !undA & doA); undA = undA | (!inA & !doA);
        outB = outB | (!inB & !undB & doB); undB = undB | (!inB & !doB);
    }
    // radial bin coordinate q = (R - Cr) / dr: proven when its fraction is further than Eq from 0 and 1 and 0 <= q < 2e9
    const P Q  = f2mul(f2sub(R, f2bc(of.Cr)), f2bc(of.inv_dr));
    const P Eq = f2fma(ER, f2bc(of.inv_dr * 1.01f), f2mul(Q, f2bc(3.2f * kEps)));
    float   qa, qb;
    f2unpack(Q, qa, qb);
    const float kqa = floorf(qa), kqb = floorf(qb);
    {
        // |fr - 1/2| < 1/2 - Eq  <=>  Eq < fr < 1 - Eq; q - floor(q) is exact, the two roundings here are inside the 4 eps
        const P G = f2sub(f2sub(Q, f2pack(kqa, kqb)), f2bc(0.5f)), H = f2sub(f2bc(0.5f - 4.0f * kEps), Eq);
        float   ga, gb, ha, hb;
        f2unpack(G, ga, gb); f2unpack(H, ha, hb);
        // as unsigned integers the bit patterns of [+0, 2e9) are ordered and everything negative / NaN lies above
        undA = undA || !(fabsf(ga) < ha) || !(__float_as_uint(qa) < __float_as_uint(2.0e9f));
        undB = undB || !(fabsf(gb) < hb) || !(__float_as_uint(qb) < __float_as_uint(2.0e9f));
    }
    // vecOut up to the sign wsgn; |v^ - v| <= Ev per component
    const P Ev = f2mul(f2add(MDV, MV), f2bc(1.05f * kEps));
    P       VO[3], Evo;
    if (method == 2)
    {
        VO[0] = f2sub(f2mul(V1[1], V2[2]), f2mul(V1[2], V2[1]));
        VO[1] = f2sub(f2mul(V1[2], V2[0]), f2mul(V1[0], V2[2]));
        VO[2] = f2sub(f2mul(V1[0], V2[1]), f2mul(V1[1], V2[0]));
        // |d(ab - cd)| <= Ev (|a|+|b|+|c|+|d|) + 2 Ev^2 + 2 eps (|ab| + |cd|), inflated by 5 %
        Evo = f2mul(f2fma(f2mul(MV, f2bc(4.0f)), Ev, f2fma(f2add(Ev, Ev), Ev, f2mul(f2mul(MV, f2bc(4.0f * kEps)), MV))), f2bc(1.05f));
    }
    else
    {
        VO[0] = f2add(V1[0], V2[0]); VO[1] = f2add(V1[1], V2[1]); VO[2] = f2add(V1[2], V2[2]);
        Evo   = f2fma(MV, f2bc(2.1f * kEps), f2mul(Ev, f2bc(2.1f)));
    }
    const P N2 = f2fma(VO[2], VO[2], f2fma(VO[1], VO[1], f2mul(VO[0], VO[0])));
    float   n2a, n2b, ria, rib;
    f2unpack(N2, n2a, n2b);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ria) : "f"(n2a));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rib) : "f"(n2b));
    const P RI  = f2pack(ria, rib);
    const P DOT = DIMNOrient == 0 ? VO[0] : (DIMNOrient == 1 ? VO[1] : VO[2]);
    const P CS  = f2mul(DOT, f2mul(RI, f2bc(vsgn)));
    const P REL = f2mul(Evo, RI);
    const P Ecs = f2fma(REL, f2bc(3.0f), f2bc(10.0f * kEps));   // the float roundings of the edge table are inside fmh[].y
    const P OM  = f2sub(f2bc(1.0f), Ecs);
    const P CI  = f2fma(CS, f2bc(0.5f * kCosLut), f2bc(0.5f * kCosLut));
    float   csa, csb, rela, relb, ecsa, ecsb, oma, omb, cia, cib;
    f2unpack(CS, csa, csb); f2unpack(REL, rela, relb); f2unpack(Ecs, ecsa, ecsb); f2unpack(OM, oma, omb); f2unpack(CI, cia, cib);
    undA = undA || !(rela < 0.05f && n2a > 1.0e-30f && fabsf(csa) < oma);
    undB = undB || !(relb < 0.05f && n2b > 1.0e-30f && fabsf(csb) < omb);
    // angle bin; a NaN or out-of-range cosine has set `und` above and only needs a valid look-up cell
    int loA = lut[min((unsigned)__float2int_rz(cia), (unsigned)(kCosLut - 1))];
    int loB = lut[min((unsigned)__float2int_rz(cib), (unsigned)(kCosLut - 1))];
    loA += csa <= fedge[loA + 1] ? 1 : 0;
    loB += csb <= fedge[loB + 1] ? 1 : 0;
    const float2 mhA = fmh[loA], mhB = fmh[loB];
    undA = undA || !(fabsf(csa - mhA.x) < mhA.y - ecsa);
    undB = undB || !(fabsf(csb - mhB.x) < mhB.y - ecsb);
    // per molecule: 0 = undecided, 1 = proven "not counted", 2 = proven selected with radial bin meZ and angle bin lo (either may
    // still lie outside its array: the caller accounts for that exactly like orient_molecule())
    res[0] = (outA || !validA) ? 1 : (undA ? 0 : 2);   // a molecule past the end of the group counts as "not counted"
    res[1] = (outB || !validB) ? 1 : (undB ? 0 : 2);
    meZ[0] = (int)kqa; meZ[1] = (int)kqb;
    lo[0] = loA; lo[1] = loB;
}

// Three-site molecule, vectors from the reference atom to the other two: the vectors ARE the unwrapped offsets (see above).
template <int SPEC = 0>
__device__ __forceinline__ void orient_fast3x2(const Mol3x2& m, bool validA, bool validB, float w1, float w2, float wsgn, const FastParams& fp,
                                               const OrientDev& d, const OrientFast& of, const float* fedge, const float2* fmh,
                                               const unsigned char* lut, int res[2], int meZ[2], int lo[2])
{
    typedef unsigned long long P;   // (molecule A, molecule B)
    const P magic2 = f2bc(kRndMagic);
    P       U1[3], U2[3];
    float   mdA = 0.f, mdB = 0.f, muA = 0.f, muB = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const P D1 = f2sub(m.X1[k], m.X0[k]);   // d = fl(x_j - x_0)
        const P D2 = f2sub(m.X2[k], m.X0[k]);
        U1[k]      = f2fma(f2sub(f2fma(D1, m.iL2[k], magic2), magic2), m.nL2[k], D1);   // u^ = fl(d - rint(d / L) L)
        U2[k]      = f2fma(f2sub(f2fma(D2, m.iL2[k], magic2), magic2), m.nL2[k], D2);
        float a, b, e, f;
        f2unpack(D1, a, b); f2unpack(D2, e, f);
        mdA = fmaxf(fmaxf(mdA, fabsf(a)), fabsf(e)); mdB = fmaxf(fmaxf(mdB, fabsf(b)), fabsf(f));
        f2unpack(U1[k], a, b); f2unpack(U2[k], e, f);
        muA = fmaxf(fmaxf(muA, fabsf(a)), fabsf(e)); muB = fmaxf(fmaxf(muB, fabsf(b)), fabsf(f));
    }
    const bool undA = !(muA < m.hsmin && mdA <= m.dmax), undB = !(muB < m.hsmin && mdB <= m.dmax);   // shift count not proven
    const P MU = f2pack(muA, muB), MD = f2pack(mdA, mdB);
    const P W1 = f2bc(w1), W2 = f2bc(w2), IW = f2bc(fp.invW);
    P       C[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) C[k] = f2fma(f2fma(W2, U2[k], f2mul(W1, U1[k])), IW, m.X0[k]);
    orient_tail_x2<SPEC>(C, MU, MD, U1, U2, MU, MD, undA, undB, validA, validB, wsgn, fp, d, of, fedge, fmh, lut, res, meZ, lo);
}

// Any molecule, any three vector atoms, two molecules per thread: the offsets of every atom from the reference atom feed the
// weighted centre sum as they are produced (nothing per atom stays live), the two vectors are the minimum images of
// x_tp1 - x_tp2 and x_tp1 - x_tp3 (proven and unique below hsmin, exactly as in orient_fast()).
// `rot` (0 when the stride is odd): with one molecule per lane the lanes of a warp read their atom j at a stride of 3 * napm
// words, which for even napm puts 2 to 16 lanes on the same shared-memory bank.  The FP32 sums here do not depend on the
// order of the atoms, so lane block m = lane / (32 / g), g = gcd(3 napm, 32), starts its walk at atom m: the g lanes that
// would collide read g different atoms, 3 m words apart, and 3 m mod g is a permutation.  The walk then includes the
// reference atom itself (offset exactly 0, contributes nothing), so that the wrap-around is a whole molecule = a multiple of g.
template <bool SMEM, int SPEC = 0>
__device__ __forceinline__ void orient_fastNx2(const float* pA, const float* pB, bool validA, bool validB, int napm, int rot, const float* swf,
                                               const FastParams& fp, const FastPack& fk, const OrientDev& d, const OrientFast& of,
                                               const float* fedge, const float2* fmh, const unsigned char* lut, int res[2], int meZ[2],
                                               int lo[2])
{
    typedef unsigned long long P;
    const P magic2 = f2bc(kRndMagic);
    P       X0[3], nL2[3], iL2[3], ACC[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        nL2[k] = *reinterpret_cast<const P*>(&fk.negL2[2 * k]);
        iL2[k] = *reinterpret_cast<const P*>(&fk.invL2[2 * k]);
        X0[k]  = f2pack(ldf<SMEM>(pA + k), ldf<SMEM>(pB + k));
        ACC[k] = f2bc(0.f);
    }
    float mdA = 0.f, mdB = 0.f, muA = 0.f, muB = 0.f;
    auto offset = [&](P X, P REF, int k, float& ma_d, float& mb_d, float& ma_u, float& mb_u) -> P {
        const P D = f2sub(X, REF);
        const P U = f2fma(f2sub(f2fma(D, iL2[k], magic2), magic2), nL2[k], D);
        float   a, b;
        f2unpack(D, a, b); ma_d = fmaxf(ma_d, fabsf(a)); mb_d = fmaxf(mb_d, fabsf(b));
        f2unpack(U, a, b); ma_u = fmaxf(ma_u, fabsf(a)); mb_u = fmaxf(mb_u, fabsf(b));
        return U;
    };
    int jj = rot;   // rot < napm (host)
#pragma unroll 2
    for (int j = rot ? 0 : 1; j < napm; ++j)
    {
        const int ja = rot ? jj : j;
        const P   W  = f2bc(swf[ja]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            ACC[k] = f2fma(W, offset(f2pack(ldf<SMEM>(pA + 3 * ja + k), ldf<SMEM>(pB + 3 * ja + k)), X0[k], k, mdA, mdB, muA, muB), ACC[k]);
        jj = jj + 1 == napm ? 0 : jj + 1;
    }
    float mdvA = 0.f, mdvB = 0.f, mvA = 0.f, mvB = 0.f;
    P     V1[3], V2[3], C[3];
    const P IW = f2bc(fp.invW);
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const P T1 = f2pack(ldf<SMEM>(pA + 3 * d.tp1 + k), ldf<SMEM>(pB + 3 * d.tp1 + k));
        V1[k] = offset(T1, f2pack(ldf<SMEM>(pA + 3 * d.tp2 + k), ldf<SMEM>(pB + 3 * d.tp2 + k)), k, mdvA, mdvB, mvA, mvB);
        V2[k] = offset(T1, f2pack(ldf<SMEM>(pA + 3 * d.tp3 + k), ldf<SMEM>(pB + 3 * d.tp3 + k)), k, mdvA, mdvB, mvA, mvB);
        C[k]  = f2fma(ACC[k], IW, X0[k]);
    }
    const bool undA = !(fmaxf(muA, mvA) < fk.hsmin && fmaxf(mdA, mdvA) <= fk.dmax);
    const bool undB = !(fmaxf(muB, mvB) < fk.hsmin && fmaxf(mdB, mdvB) <= fk.dmax);
    orient_tail_x2<SPEC>(C, f2pack(muA, muB), f2pack(mdA, mdB), V1, V2, f2pack(mvA, mvB), f2pack(mdvA, mdvB), undA, undB, validA, validB, 1.0f,
                   fp, d, of, fedge, fmh, lut, res, meZ, lo);
}

} // namespace b2a
