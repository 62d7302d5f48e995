/*
 * gmxstub_api.h -- stand-in for the slice of the GROMACS 2020 C API that the
 * itp::GmxHandle front end and its analysis programs touch.
 *
 * libgromacs 2020.1 is an external, un-vendored dependency of the reference
 * (reference src/general/CMakeLists.txt:2) and is absent from this image.  This header
 * plus gmxstub.cpp provide a link-compatible replacement for the handful of entry
 * points on the per-frame path (reference include/itp/gmx_bits/gmx_handle.ipp:17,26,34,
 * 70,83,397-407; gmx_config.hpp:7).  Type and function NAMES follow the public GROMACS
 * API so sources written against <gromacs/...> compile unchanged; the implementation
 * behind them is ours: trajectories, topologies and index groups are read from the
 * simple files documented in gmxstub/README.md instead of XTC/TRR/TPR.
 *
 * With a real GROMACS installation this directory is simply left off the include path.
 */
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

/* ---- scalar / vector basics ------------------------------------------------------- */
typedef float real; /* GROMACS mixed precision build */
typedef int   gmx_bool;
#ifndef FALSE
#define FALSE 0
#endif
#ifndef TRUE
#define TRUE 1
#endif
#define XX 0
#define YY 1
#define ZZ 2
#define DIM 3
typedef real   rvec[DIM];
typedef double dvec[DIM];
typedef int    ivec[DIM];
typedef real   matrix[DIM][DIM];

/* ---- command line ----------------------------------------------------------------- */
enum { etINT, etINT64, etREAL, etTIME, etSTR, etBOOL, etRVEC, etENUM, etNR };

struct t_pargs
{
    const char* option;
    gmx_bool    bSet;
    int         type;
    union {
        void*        v;
        int*         i;
        int64_t*     is;
        real*        r;
        const char** c;
        gmx_bool*    b;
        rvec*        rv;
    } u;
    const char* desc;
};

enum {
    efMDP, efTRX, efTRO, efTRN, efTRR, efCOMPRESSED, efXTC, efTNG, efEDR, efSTX, efSTO, efGRO,
    efG96, efPDB, efBRK, efENT, efESP, efPQR, efCPT, efLOG, efXVG, efOUT, efNDX, efTOP, efITP,
    efTPS, efTPR, efTEX, efRTP, efATP, efHDB, efDAT, efDLG, efMAP, efEPS, efMAT, efM2P, efMTX,
    efEDI, efCUB, efXPM, efRND, efCSV, efNR
};

#define ffSET (1 << 0)
#define ffREAD (1 << 1)
#define ffWRITE (1 << 2)
#define ffOPT (1 << 3)
#define ffLIB (1 << 4)
#define ffMULT (1 << 5)
#define ffALLOW_MISSING (1 << 6)
#define ffRW (ffREAD | ffWRITE)
#define ffOPTRD (ffREAD | ffOPT)
#define ffOPTWR (ffWRITE | ffOPT)
#define ffOPTRW (ffRW | ffOPT)
#define ffLIBRD (ffREAD | ffLIB)
#define ffLIBOPTRD (ffOPTRD | ffLIB)
#define ffRDMULT (ffREAD | ffMULT)
#define ffOPTRDMULT (ffRDMULT | ffOPT)
#define ffWRMULT (ffWRITE | ffMULT)
#define ffOPTWRMULT (ffWRMULT | ffOPT)

struct t_filenm
{
    int                      ftp;
    const char*              opt;
    const char*              fn;
    unsigned long            flag;
    std::vector<std::string> filenames;
};

#define PCA_CAN_VIEW (1 << 5)
#define PCA_CAN_BEGIN (1 << 6)
#define PCA_CAN_END (1 << 7)
#define PCA_CAN_DT (1 << 14)
#define PCA_CAN_TIME (PCA_CAN_BEGIN | PCA_CAN_END | PCA_CAN_DT)
#define PCA_TIME_UNIT (1 << 15)
#define PCA_KEEP_ARGS (1 << 8)
#define PCA_CAN_SET_DEFFNM (1 << 10)
#define PCA_NOEXIT_ON_ARGS (1 << 11)
#define PCA_QUIET (1 << 12)
#define PCA_NOT_READ_NODE (1 << 16)
#define PCA_DISABLE_INPUT_FILE_CHECKING (1 << 17)

struct gmx_output_env_t;

gmx_bool parse_common_args(int* argc, char* argv[], unsigned long Flags, int nfile, t_filenm fnm[],
                           int npargs, t_pargs* pa, int ndesc, const char** desc, int nbugs,
                           const char** bugs, gmx_output_env_t** oenv);

const char* opt2fn(const char* opt, int nfile, const t_filenm fnm[]);
const char* opt2fn_null(const char* opt, int nfile, const t_filenm fnm[]);
const char* ftp2fn(int ftp, int nfile, const t_filenm fnm[]);
const char* ftp2fn_null(int ftp, int nfile, const t_filenm fnm[]);
gmx_bool    opt2bSet(const char* opt, int nfile, const t_filenm fnm[]);
gmx_bool    opt2parg_bSet(const char* opt, int nargs, const t_pargs pa[]);

int gmx_run_cmain(int argc, char* argv[], int (*mainFunction)(int, char*[]));

/* ---- errors / memory -------------------------------------------------------------- */
#define FARGS 0, __FILE__, __LINE__
[[noreturn]] void gmx_fatal(int fatal_errno, const char* file, int line, const char* fmt, ...);

template <typename T>
static inline void gmxstub_snew(T*& ptr, size_t n)
{
    ptr = static_cast<T*>(std::calloc(n ? n : 1, sizeof(T)));
    if (!ptr) { std::fprintf(stderr, "gmxstub: out of memory\n"); std::abort(); }
}
template <typename T>
static inline void gmxstub_srenew(T*& ptr, size_t n)
{
    ptr = static_cast<T*>(std::realloc(ptr, (n ? n : 1) * sizeof(T)));
    if (!ptr) { std::fprintf(stderr, "gmxstub: out of memory\n"); std::abort(); }
}
#define snew(ptr, nelem) gmxstub_snew((ptr), (nelem))
#define srenew(ptr, nelem) gmxstub_srenew((ptr), (nelem))
#define sfree(ptr) std::free((void*)(ptr))

/* ---- topology --------------------------------------------------------------------- */
struct t_atom
{
    real           m, q;
    real           mB, qB;
    unsigned short type;
    unsigned short typeB;
    int            ptype;
    int            resind;
    int            atomnumber;
    char           elem[4];
};

struct t_resinfo
{
    char**        name;
    int           nr;
    unsigned char ic;
    int           chainnum;
    char          chainid;
    char**        rtp;
};

struct t_atoms
{
    int        nr;
    t_atom*    atom;
    char***    atomname;
    char***    atomtype;
    char***    atomtypeB;
    int        nres;
    t_resinfo* resinfo;
    void*      pdbinfo;
    gmx_bool   haveMass, haveCharge, haveType, haveBState, havePdbInfo;
};

struct t_block
{
    int  nr;
    int* index;
    int  nalloc_index;
};

struct t_atomtypes
{
    int  nr;
    int* atomnumber;
};

typedef union t_iparams {
    struct { real c6, c12; } lj;
    struct { real c6, c12, c6_14, c12_14; } lj14;
    struct { real a, b, c; } bham;
    struct { real buf[12]; } generic;
} t_iparams;

typedef int t_functype;

struct t_idef
{
    int         ntypes;
    int         atnr;
    t_functype* functype;
    t_iparams*  iparams;
    real        fudgeQQ;
};

struct t_symtab { int nr; void* symbuf; };

struct t_topology
{
    char**      name;
    t_idef      idef;
    t_atoms     atoms;
    t_atomtypes atomtypes;
    t_block     mols;
    gmx_bool    bIntermolecularInteractions;
    t_symtab    symtab;
};

struct t_inputrec
{
    int    eI;
    int64_t nsteps;
    double delta_t;
    int    ePBC;
    real   rcoulomb, rvdw;
};

int read_tpx_top(const char* fn, t_inputrec* ir, matrix box, int* natoms, rvec* x, rvec* v, t_topology* top);

void get_index(const t_atoms* atoms, const char* fnm, int ngrps, int isize[], int* index[], char* grpnames[]);
void rd_index(const char* statfile, int ngrps, int isize[], int* index[], char* grpnames[]);

/* ---- trajectory ------------------------------------------------------------------- */
#define TRX_READ_X (1 << 0)
#define TRX_NEED_X (1 << 1)
#define TRX_READ_V (1 << 2)
#define TRX_NEED_V (1 << 3)
#define TRX_READ_F (1 << 4)
#define TRX_NEED_F (1 << 5)
#define TRX_DONT_SKIP (1 << 6)

struct t_trxframe
{
    int      not_ok;
    gmx_bool bDouble;
    int      natoms;
    gmx_bool bStep;
    int64_t  step;
    gmx_bool bTime;
    real     time;
    gmx_bool bLambda;
    gmx_bool bFepState;
    real     lambda;
    int      fep_state;
    gmx_bool bAtoms;
    t_atoms* atoms;
    gmx_bool bPrec;
    real     prec;
    gmx_bool bX;
    rvec*    x;
    gmx_bool bV;
    rvec*    v;
    gmx_bool bF;
    rvec*    f;
    gmx_bool bBox;
    matrix   box;
    gmx_bool bPBC;
    int      ePBC;
    gmx_bool bIndex;
    int*     index;
};

struct t_trxstatus;

bool read_first_frame(const gmx_output_env_t* oenv, t_trxstatus** status, const char* fn, t_trxframe* fr, int flags);
bool read_next_frame(const gmx_output_env_t* oenv, t_trxstatus* status, t_trxframe* fr);
void close_trx(t_trxstatus* status);
int  nframes_read(t_trxstatus* status);

FILE* gmx_ffopen(const char* file, const char* mode);
int   gmx_ffclose(FILE* fp);

/* ---- stub-only extension: block access to the trajectory --------------------------- *
 * A batching front end wants K frames per call.  With real libgromacs this is a loop
 * over read_next_frame(); the stand-in also exposes it directly so a K-frame pinned
 * batch can be filled with one call.  Returns the number of frames delivered (0 at EOF).
 * x_out: [K][natoms][3] floats, box_out: [K][9] floats, time_out: [K] floats.            */
int gmxstub_read_frames(t_trxstatus* status, int max_frames, float* x_out, float* box_out, float* time_out);
int gmxstub_natoms(const t_trxstatus* status);
#define GMXSTUB_BATCH_READ 1
/* The stand-in reads its own raw format and real XTC / TRR files (gmxstub/xdrtraj.hpp; recognised by magic).
 * Writers and whole-file readers for tests and tools:                                                        */
#ifdef __cplusplus
extern "C" {
#endif
int gmxstub_write_xtc(const char* path, int natoms, int nframes, const float* x, const float* box9, int nbox, const float* times, float precision);
int gmxstub_write_trr(const char* path, int natoms, int nframes, const float* x, const float* v, const float* f, const float* box9, int nbox,
                      const float* times);
int gmxstub_read_all(const char* path, int max_frames, int want_flags, float* x_out, float* v_out, float* f_out, float* box_out, float* time_out,
                     int* natoms_out);
int gmxstub_read_batches(const char* path, int batch, int max_frames, float* x_out, float* box_out, float* time_out);
#ifdef __cplusplus
}
#endif
