/*
 * oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's per-frame analysis path.
 * See oracle.h for the pinning statement and the citation of every function.
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA, SSE2 doubles;
 * reference CMakeLists.txt:12-29 has no -march), so every multiply-add rounds twice.
 */
#include "oracle.h"

#include <math.h>
#include <stddef.h>

double orc_periodicity(double dx, double box)
{
    while (dx > box / 2.0) dx -= box;
    while (dx < -box / 2.0) dx += box;
    return dx;
}

double orc_eigen_sum(const double* w, int n)
{
    /* Eigen::internal::redux_impl<sum, Evaluator, LinearVectorizedTraversal, NoUnrolling>: two packet
     * accumulators of 2 doubles each, combined, horizontally added, then the scalar tail. */
    const int packet = 2;
    const int aligned2 = (n / (2 * packet)) * (2 * packet);
    const int aligned1 = (n / packet) * packet;
    double    res;
    if (aligned1)
    {
        double p0[2] = { w[0], w[1] };
        if (aligned1 > packet)
        {
            double p1[2] = { w[2], w[3] };
            for (int i = 2 * packet; i < aligned2; i += 2 * packet)
            {
                p0[0] += w[i];
                p0[1] += w[i + 1];
                p1[0] += w[i + 2];
                p1[1] += w[i + 3];
            }
            p0[0] += p1[0];
            p0[1] += p1[1];
            if (aligned1 > aligned2)
            {
                p0[0] += w[aligned2];
                p0[1] += w[aligned2 + 1];
            }
        }
        res = p0[0] + p0[1];
        for (int i = aligned1; i < n; ++i) res += w[i];
    }
    else
    {
        res = w[0];
        for (int i = 1; i < n; ++i) res += w[i];
    }
    return res;
}

void orc_load_position(const float* x, const int* index, int nmol, int napm, const double Lbox[3], double* pos)
{
    double half[3];
    for (int k = 0; k < 3; ++k) half[k] = Lbox[k] / 2;
    int n = 0;
    for (int i = 0; i < nmol; ++i)
    {
        const int molN = index[n];
        for (int j = 0; j < napm; ++j)
        {
            for (int k = 0; k < 3; ++k)
            {
                double       t   = x[(size_t)index[n] * 3 + k];
                const double ref = x[(size_t)molN * 3 + k];
                while (t - ref > half[k]) t -= Lbox[k];
                while (t - ref < -half[k]) t += Lbox[k];
                pos[((size_t)i + (size_t)j * nmol) * 3 + k] = t;
            }
            n++;
        }
    }
}

static void center_core(const float* x, const int* index, int nmol, int napm, const double* w, double wsum,
                        const double Lbox[3], double* posc)
{
    double half[3];
    for (int k = 0; k < 3; ++k) half[k] = Lbox[k] / 2;
    int n = 0;
    for (int i = 0; i < nmol; ++i)
    {
        const int molN = index[n];
        for (int k = 0; k < 3; ++k)
        {
            double       c   = 0.0;
            const double ref = x[(size_t)molN * 3 + k];
            for (int j = 0; j < napm; ++j)
            {
                double t = x[(size_t)(molN + j) * 3 + k]; /* contiguity assumed by the reference, ipp:187 */
                while (t - ref > half[k]) t -= Lbox[k];
                while (t - ref < -half[k]) t += Lbox[k];
                c += t * w[j];
            }
            posc[(size_t)i + (size_t)k * nmol] = c / wsum;
        }
        n += napm;
    }
}

void orc_load_position_center(const float* x, const int* index, int nmol, int napm, const double* w,
                              const double Lbox[3], double* posc)
{
    center_core(x, index, nmol, napm, w, orc_eigen_sum(w, napm), Lbox, posc);
}

void orc_dipole_center(const float* x, const int* index, int nmol, int napm, const double* q, const double Lbox[3],
                       double* posc)
{
    double s = 0.0; /* 9g_dist_bin_1803.cpp:217-219: plain sequential sum */
    for (int j = 0; j < napm; ++j) s += q[j];
    if (fabs(s) < 1e-6) s = 1.0;
    center_core(x, index, nmol, napm, q, s, Lbox, posc);
}

void orc_density_accum(const double* posc, int nmol, int dim, float low, float up, double dbin, int nbin,
                       int upper_inclusive, int dim2, float low2, float up2, uint64_t* hist, uint64_t* dropped)
{
    for (int i = 0; i < nmol; ++i)
    {
        const double p  = posc[(size_t)i + (size_t)dim * nmol];
        int          ok = (low <= p) && (upper_inclusive ? (p <= up) : (p < up));
        if (ok && dim2 >= 0)
        {
            const double p2 = posc[(size_t)i + (size_t)dim2 * nmol];
            ok              = (low2 <= p2) && (upper_inclusive ? (p2 <= up2) : (p2 < up2));
        }
        if (!ok) continue;
        const int me = (int)((p - low) / dbin);
        if (me < 0 || me >= nbin) { if (dropped) (*dropped)++; continue; }
        hist[me]++;
    }
}

float orc_rdf_default_rm(const double Lbox[3])
{
    float Rm = (float)(Lbox[0] < Lbox[1] ? Lbox[0] : Lbox[1]);         /* :54 */
    Rm       = (float)((double)Rm < Lbox[2] ? (double)Rm : Lbox[2]);   /* :55 */
    Rm /= 2;                                                           /* :56 */
    return Rm;
}

int orc_rdf_rbin(float Rm) { return (int)(Rm / 0.002); } /* :58 */

void orc_rdf_accum(const double* posc1, int n0, const double* posc2, int n1, const double Lbox[3], int nbin,
                   float lowPos, float upPos, int dim, int dim2, float lowPos2, float upPos2, float Rm, int Rbin,
                   uint64_t* rdf, uint64_t* stats)
{
    const double rm2   = (double)(Rm * Rm); /* float product, promoted for the compare (:88) */
    const float  width = upPos - lowPos;    /* float subtraction (:77) */
    uint64_t     counted = 0, oob = 0, npass = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : counted, oob, npass)
    for (int i = 0; i < n0; ++i)
    {
        const double pd = posc1[(size_t)i + (size_t)dim * n0], pd2 = posc1[(size_t)i + (size_t)dim2 * n0];
        if (!(lowPos < pd && pd < upPos && lowPos2 < pd2 && pd2 < upPos2)) continue;
        npass++;
        const int meZ = (int)(nbin * (pd - lowPos) / width);
        const double xi = posc1[i], yi = posc1[(size_t)i + n0], zi = posc1[(size_t)i + 2 * (size_t)n0];
        for (int j = 0; j < n1; ++j)
        {
            double R = 0, r;
            r = orc_periodicity(xi - posc2[j], Lbox[0]);
            R += r * r;
            r = orc_periodicity(yi - posc2[(size_t)j + n1], Lbox[1]);
            R += r * r;
            r = orc_periodicity(zi - posc2[(size_t)j + 2 * (size_t)n1], Lbox[2]);
            R += r * r;
            if (1e-6 < R && R < rm2)
            {
                const int meR = (int)(Rbin * sqrt(R) / Rm);
                if (meR >= Rbin || meZ < 0 || meZ >= nbin) { oob++; continue; }
                uint64_t* cell = &rdf[(size_t)meZ + (size_t)meR * nbin];
#pragma omp atomic
                (*cell)++;
                counted++;
            }
        }
    }
    if (stats) { stats[0] += counted; stats[1] += oob; stats[2] += npass; }
}

void orc_rdf_normalise(const uint64_t* rdf, int nbin, int Rbin, double* g)
{
    /* volumn[i] = pow(0.002*(i+1),3), then differences from the top down (:100-109) */
    for (int r = 0; r < Rbin; ++r)
    {
        double v = pow(0.002 * (r + 1), 3);
        if (r > 0) v = v - pow(0.002 * r, 3);
        for (int z = 0; z < nbin; ++z) g[(size_t)z + (size_t)r * nbin] = (double)rdf[(size_t)z + (size_t)r * nbin] / v;
    }
    /* mean over the last int(Rbin*0.1) columns, row by row (:113), then divide (:115) */
    const int tail = (int)(Rbin * 0.1);
    for (int z = 0; z < nbin; ++z)
    {
        double s = 0.0;
        for (int r = Rbin - tail; r < Rbin; ++r) s += g[(size_t)z + (size_t)r * nbin];
        const double mean = s / tail;
        for (int r = 0; r < Rbin; ++r) g[(size_t)z + (size_t)r * nbin] /= mean;
    }
}

static double angle_deg(const double v1[3], const double v2[3])
{
    /* calc_orient_mof.cpp:41-47; itp::pi is a long double constant, so the division is done in long double */
    const double      r1  = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    const double      r2  = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    const double      c   = (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / (r1 * r2);
    const long double pi  = 3.14159265358979323846264338327950288L;
    return (double)(acos(c) / pi * 180);
}

void orc_orient_accum(const double* pos, const double* posc, int nmol, int napm, const double Lbox[3], int DIMN,
                      float lowPos, float upPos, float c1, float c2, float dr, float CR, float Cr, int nAngle,
                      int tp1, int tp2, int tp3, int DIMNOrient, int method, uint64_t* orient, uint64_t* ncm,
                      uint64_t* dropped)
{
    (void)napm;
    const int    nbin1  = (int)((CR - Cr) / dr);   /* float arithmetic, :133 */
    const double dAngle = 180 / nAngle;            /* INTEGER division, :143 */
    double       wall[3] = { 0, 0, 0 };
    wall[DIMNOrient]     = 1.0;
    tp1--; tp2--; tp3--;
    for (int i = 0; i < nmol; ++i)
    {
        const double pz = posc[(size_t)i + (size_t)DIMN * nmol];
        if (!(lowPos <= pz && pz <= upPos)) continue;
        double t1 = 0, t2 = 0;
        if (DIMN == 0) { t1 = posc[(size_t)i + 1 * (size_t)nmol] - c1; t2 = posc[(size_t)i + 2 * (size_t)nmol] - c2; }
        if (DIMN == 1) { t1 = posc[(size_t)i + 0 * (size_t)nmol] - c1; t2 = posc[(size_t)i + 2 * (size_t)nmol] - c2; }
        if (DIMN == 2) { t1 = posc[(size_t)i + 0 * (size_t)nmol] - c1; t2 = posc[(size_t)i + 1 * (size_t)nmol] - c2; }
        const double R = sqrt(t1 * t1 + t2 * t2);
        if (!(Cr <= R && R <= CR)) continue;
        const int meZ = (int)((R - Cr) / dr);
        if (meZ < 0 || meZ >= nbin1) { if (dropped) (*dropped)++; continue; }
        ncm[meZ]++;
        double v1[3], v2[3], vo[3];
        for (int k = 0; k < 3; ++k)
        {
            const double p1 = pos[((size_t)i + (size_t)tp1 * nmol) * 3 + k];
            const double p2 = pos[((size_t)i + (size_t)tp2 * nmol) * 3 + k];
            const double p3 = pos[((size_t)i + (size_t)tp3 * nmol) * 3 + k];
            v1[k]           = orc_periodicity(p1 - p2, Lbox[k]);
            v2[k]           = orc_periodicity(p1 - p3, Lbox[k]);
        }
        if (method == 2)
        {
            vo[0] = v1[1] * v2[2] - v1[2] * v2[1];
            vo[1] = v1[2] * v2[0] - v1[0] * v2[2];
            vo[2] = v1[0] * v2[1] - v1[1] * v2[0];
        }
        else
        {
            vo[0] = v1[0] + v2[0]; vo[1] = v1[1] + v2[1]; vo[2] = v1[2] + v2[2];
        }
        const double ang = angle_deg(vo, wall);
        const int    meA = (int)(ang / dAngle);
        if (!(meA >= 0 && meA < nAngle)) { if (dropped) (*dropped)++; continue; }
        orient[(size_t)meZ + (size_t)meA * nbin1]++;
    }
}

/* ---- GmxHandleFull: heterogeneous molecule sizes (include/itp/gmx_bits/gmx_handle_full.ipp:115-199, 323-344) ---- */
void orc_full_load_position(const float* x, const int* index, int nmol, const int* napm, int max_napm, const double Lbox[3],
                            double* pos)
{
    (void)max_napm;
    double half[3];
    for (int k = 0; k < 3; ++k) half[k] = Lbox[k] / 2;
    int n = 0;
    for (int i = 0; i < nmol; ++i)
    {
        const int molN = index[n];
        for (int j = 0; j < napm[i]; ++j, ++n)
            for (int k = 0; k < 3; ++k)
            {
                double       t   = x[(size_t)index[n] * 3 + k];
                const double ref = x[(size_t)molN * 3 + k];
                while (t - ref > half[k]) t -= Lbox[k];
                while (t - ref < -half[k]) t += Lbox[k];
                pos[((size_t)i + (size_t)j * nmol) * 3 + k] = t;
            }
    }
}

/* w: weight per index position; com 1 (geometry) is given its evident intent: weights 1, divisor napm[i] */
void orc_full_load_position_center(const float* x, const int* index, int nmol, const int* napm, const double* w, int geometry,
                                   const double Lbox[3], double* posc)
{
    double half[3];
    for (int k = 0; k < 3; ++k) half[k] = Lbox[k] / 2;
    int n = 0;
    for (int i = 0; i < nmol; ++i)
    {
        const int molN = index[n];
        double    tot  = 0.0;                       /* totMass / totCharge: sequential sum, full.ipp:337-340 */
        for (int j = 0; j < napm[i]; ++j) tot += w[n + j];
        if (geometry) tot = (double)napm[i];
        for (int k = 0; k < 3; ++k)
        {
            double       c   = 0.0;
            const double ref = x[(size_t)molN * 3 + k];
            for (int j = 0; j < napm[i]; ++j)
            {
                double t = x[(size_t)(molN + j) * 3 + k];
                while (t - ref > half[k]) t -= Lbox[k];
                while (t - ref < -half[k]) t += Lbox[k];
                c += t * (geometry ? 1.0 : w[n + j]);
            }
            posc[(size_t)i + (size_t)k * nmol] = c / tot;
        }
        n += napm[i];
    }
}

void orc_field_center(const float* v, const int* index, int nmol, int napm, const double* w, double wsum, double* out)
{
    int n = 0;
    for (int i = 0; i < nmol; ++i)
    {
        const int molN = index[n];
        for (int k = 0; k < 3; ++k)
        {
            double c = 0.0;
            for (int j = 0; j < napm; ++j)
            {
                const double t = v[(size_t)(molN + j) * 3 + k];
                c += t * w[j];
            }
            out[(size_t)i + (size_t)k * nmol] = c / wsum;
        }
        n += napm;
    }
}

void orc_field_values(const float* v, const int* index, int nmol, int napm, double* out)
{
    int n = 0;
    for (int i = 0; i < nmol; ++i)
        for (int j = 0; j < napm; ++j, ++n)
            for (int k = 0; k < 3; ++k) out[((size_t)i + (size_t)j * nmol) * 3 + k] = v[(size_t)index[n] * 3 + k];
}

int orc_pair_count(const double* posc1, int n0, const double* posc2, int n1, const double Lbox[3], float Rmin, int dim,
                   float lowPos, float upPos, int use_cyl, float cx, float cy, float Rlow, float Rup, int* count)
{
    int numA = 0;
#pragma omp parallel for schedule(static) reduction(+ : numA)
    for (int i = 0; i < n0; ++i)
    {
        count[i]        = -1;
        const double pd = posc1[(size_t)i + (size_t)dim * n0];
        if (!(lowPos <= pd && pd <= upPos)) continue;
        if (use_cyl)
        {
            /* std::pow(x, 2) is x * x (gcc folds it; glibc's pow is exact for squares as well) */
            const double ax = posc1[i] - cx, ay = posc1[(size_t)i + n0] - cy;
            const double L1 = sqrt(ax * ax + ay * ay);
            if (!(Rlow <= L1 && L1 <= Rup)) continue;
        }
        numA++;
        int pairNumber = 0;
        for (int j = 0; j < n1; ++j)
        {
            double ret = 0.0;
            for (int k = 0; k < 3; ++k)
            {
                const double r = orc_periodicity(posc1[(size_t)i + (size_t)k * n0] - posc2[(size_t)j + (size_t)k * n1], Lbox[k]);
                ret += r * r;
            }
            const double distance = sqrt(ret);
            if (distance > 1e-3 && distance <= Rmin) pairNumber++;
        }
        count[i] = pairNumber;
    }
    return numA;
}
