/*
 * ref_handle_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" access to the reference's own
 * itp::GmxHandle loaders, compiled from the headers where they lie under /root/reference/include
 * (never copied).  Built into oracle/_ref/libref_handle.so by oracle/Makefile; used only by tests/
 * and tools/make_golden.py to pin oracle/port/oracle.c and to generate tests/golden fixtures.
 *
 * The handle's data members are public (reference include/itp/gmx_bits/gmx_handle.hpp:162-184), so
 * instead of going through init()/read_first_frame() the shim fills them directly from caller arrays
 * and then calls the unmodified loadPosition / loadPositionCenter / periodicity
 * (gmx_handle.ipp:115-148, 150-199, 410-417).
 */
#include <itp/gmx>

#include <cstring>

namespace
{
struct Shim
{
    itp::GmxHandle hd{ 0, nullptr };
    std::vector<std::vector<int>> idx;
};
} // namespace

extern "C" {

void* refshim_new(int ngrps)
{
    Shim* s      = new Shim;
    s->hd.ngrps  = ngrps;
    s->idx.resize(ngrps);
    s->hd.napm.resize(ngrps);
    s->hd.nmol.resize(ngrps);
    s->hd.mass.resize(ngrps);
    s->hd.charge.resize(ngrps);
    snew(s->hd.index, ngrps);
    snew(s->hd.ngx, ngrps);
    snew(s->hd.fr, 1);
    return s;
}

void refshim_free(void* h) { delete static_cast<Shim*>(h); }

void refshim_set_group(void* h, int grp, const int* index, int ngx, int napm, const double* mass, const double* charge)
{
    Shim* s = static_cast<Shim*>(h);
    s->idx[grp].assign(index, index + ngx);
    s->hd.index[grp] = s->idx[grp].data();
    s->hd.ngx[grp]   = ngx;
    s->hd.napm[grp]  = napm;
    s->hd.nmol[grp]  = ngx / napm;
    s->hd.mass[grp].resize(napm);
    s->hd.charge[grp].resize(napm);
    for (int j = 0; j < napm; ++j)
    {
        s->hd.mass[grp][j]   = mass[j];
        s->hd.charge[grp][j] = charge[j];
    }
}

/* x: [natoms][3] float, box: float[9] row-major; Lbox taken from the diagonal as readFirstFrame does (ipp:72-74) */
void refshim_set_frame(void* h, float* x, const float* box9)
{
    Shim* s    = static_cast<Shim*>(h);
    s->hd.fr->x = reinterpret_cast<rvec*>(x);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) s->hd.fr->box[a][b] = box9[3 * a + b];
    s->hd.Lbox[XX] = s->hd.fr->box[XX][XX];
    s->hd.Lbox[YY] = s->hd.fr->box[YY][YY];
    s->hd.Lbox[ZZ] = s->hd.fr->box[ZZ][ZZ];
}

/* out: nmol*napm*3 doubles in boxd memory order: element (i,j) at ((i + j*nmol)*3 + k) */
void refshim_load_position(void* h, int grp, double* out)
{
    Shim*     s   = static_cast<Shim*>(h);
    itp::boxd pos = s->hd.initPos(grp);
    s->hd.loadPosition(pos, grp);
    std::memcpy(out, pos.data(), sizeof(double) * 3 * pos.size());
}

/* out: nmol*3 doubles, column-major (i + k*nmol) */
void refshim_load_position_center(void* h, int grp, int com, double* out)
{
    Shim*     s    = static_cast<Shim*>(h);
    itp::matd posc = s->hd.initPosc(grp);
    s->hd.loadPositionCenter(posc, grp, com);
    std::memcpy(out, posc.data(), sizeof(double) * posc.size());
}

/* velocities / forces: point fr->v / fr->f at caller arrays, then the unmodified loaders (gmx_handle.ipp:201-339) */
void refshim_set_vf(void* h, float* v, float* f)
{
    Shim* s     = static_cast<Shim*>(h);
    s->hd.fr->v = reinterpret_cast<rvec*>(v);
    s->hd.fr->f = reinterpret_cast<rvec*>(f);
}
void refshim_load_field(void* h, int grp, int force, double* out)
{
    Shim*     s   = static_cast<Shim*>(h);
    itp::boxd val = s->hd.initPos(grp);
    if (force) s->hd.loadForce(val, grp); else s->hd.loadVelocity(val, grp);
    std::memcpy(out, val.data(), sizeof(double) * 3 * val.size());
}
void refshim_load_field_center(void* h, int grp, int force, int com, double* out)
{
    Shim*     s   = static_cast<Shim*>(h);
    itp::matd val = s->hd.initPosc(grp);
    if (force) s->hd.loadForceCenter(val, grp, com); else s->hd.loadVelocityCenter(val, grp, com);
    std::memcpy(out, val.data(), sizeof(double) * val.size());
}

/* ---- GmxHandleFull (heterogeneous molecule sizes), reference include/itp/gmx_bits/gmx_handle_full.ipp ---- */
struct ShimFull
{
    itp::GmxHandleFull hd{ 0, nullptr };
    std::vector<int>   idx;
};

void* refshim_full_new(const int* index, int ngx, int nmol, const int* napm, const double* mass, const double* charge)
{
    ShimFull* s = new ShimFull;
    s->hd.ngrps = 1;
    s->idx.assign(index, index + ngx);
    snew(s->hd.index, 1);
    snew(s->hd.ngx, 1);
    snew(s->hd.fr, 1);
    s->hd.index[0] = s->idx.data();
    s->hd.ngx[0]   = ngx;
    s->hd.napm.resize(1); s->hd.nmol.resize(1); s->hd.mass.resize(1); s->hd.charge.resize(1);
    s->hd.totMass.resize(1); s->hd.totCharge.resize(1);
    s->hd.nmol[0] = nmol;
    s->hd.napm[0].assign(napm, napm + nmol);
    s->hd.mass[0].resize(nmol); s->hd.charge[0].resize(nmol);
    s->hd.totMass[0].assign(nmol, 0.0); s->hd.totCharge[0].assign(nmol, 0.0);
    int n = 0;
    for (int i = 0; i < nmol; ++i)       /* same accumulation order as get_mass_charge, full.ipp:323-344 */
    {
        s->hd.mass[0][i].resize(napm[i]); s->hd.charge[0][i].resize(napm[i]);
        for (int j = 0; j < napm[i]; ++j, ++n)
        {
            s->hd.mass[0][i][j] = mass[n]; s->hd.charge[0][i][j] = charge[n];
            s->hd.totMass[0][i] += mass[n]; s->hd.totCharge[0][i] += charge[n];
        }
    }
    return s;
}
void refshim_full_free(void* h) { delete static_cast<ShimFull*>(h); }
void refshim_full_set_frame(void* h, float* x, const float* box9)
{
    ShimFull* s = static_cast<ShimFull*>(h);
    s->hd.fr->x = reinterpret_cast<rvec*>(x);
    s->hd.Lbox[XX] = box9[0]; s->hd.Lbox[YY] = box9[4]; s->hd.Lbox[ZZ] = box9[8];
}
/* out: nmol x maxNapm x 3 doubles in boxd order, untouched slots are returned as 0 */
void refshim_full_load_position(void* h, double* out, int maxNapm)
{
    ShimFull* s = static_cast<ShimFull*>(h);
    itp::boxd pos(s->hd.nmol[0], maxNapm);
    for (Eigen::Index i = 0; i < pos.size(); ++i) pos.data()[i].setZero();
    s->hd.loadPosition(pos, 0);
    std::memcpy(out, pos.data(), sizeof(double) * 3 * pos.size());
}
/* center 0 (mass) or 2 (charge); 1 indexes an empty vector in the reference (full.ipp:166-171) and is not callable */
void refshim_full_load_position_center(void* h, int center, double* out)
{
    ShimFull* s = static_cast<ShimFull*>(h);
    itp::matd posc;
    s->hd.initPosc(posc, 0);
    s->hd.loadPositionCenter(posc, 0, center);
    std::memcpy(out, posc.data(), sizeof(double) * posc.size());
}

double refshim_periodicity(double dx, double box) { return itp::GmxHandle::periodicity(dx, box); }

/* atomCenter.sum() exactly as ipp:195 evaluates it (Eigen 3.3.9 vectorised redux) */
double refshim_eigen_sum(const double* w, int n)
{
    itp::vecd v(n);
    for (int i = 0; i < n; ++i) v[i] = w[i];
    return v.sum();
}

/* rowwise mean of the last ncols columns of a (rows x cols) column-major matrix, as calc_rdf_region.cpp:113 */
void refshim_rightcols_rowmean(const double* m, int rows, int cols, int ncols, double* out)
{
    Eigen::Map<const itp::matd> M(m, rows, cols);
    itp::vecd                   mean = M.rightCols(ncols).rowwise().mean();
    for (int i = 0; i < rows; ++i) out[i] = mean[i];
}
}
