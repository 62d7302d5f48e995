/*
 * gmxstub.cpp -- implementation of the libgromacs stand-in declared in
 * include/gromacs/gmxstub_api.h.  See gmxstub/README.md for the file formats.
 *
 * This is NOT GROMACS code: it implements, from scratch, just enough behaviour behind
 * the GROMACS entry-point names for itp::GmxHandle-style programs to run in an image
 * that has no libgromacs (command-line parsing of t_pargs / t_filenm tables, a binary
 * topology reader, an .ndx reader with stdin group selection, and a raw float trajectory
 * reader with -b/-e/-dt time selection).
 */
#include "gromacs/gmxstub_api.h"
#include "../xdrtraj.hpp"

#include <cctype>
#include <cmath>
#include <cstring>
#include <map>
#include <sstream>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>

struct gmx_output_env_t
{
    std::string program;
};

namespace
{
struct TimeSel
{
    bool   has_b = false, has_e = false, has_dt = false;
    double b = 0, e = 0, dt = 0;
} g_time;

std::string g_program = "gmxstub";

const char* default_ext(int ftp)
{
    switch (ftp)
    {
        case efTRX: case efTRO: case efXTC: case efCOMPRESSED: return ".xtc";
        case efTRR: case efTRN: return ".trr";
        case efTPR: case efTPS: return ".tpr";
        case efNDX: return ".ndx";
        case efXVG: return ".xvg";
        case efDAT: return ".dat";
        case efLOG: return ".log";
        case efOUT: return ".out";
        case efGRO: case efSTX: case efSTO: return ".gro";
        case efPDB: return ".pdb";
        case efTOP: return ".top";
        case efCSV: return ".csv";
        default: return ".dat";
    }
}

const char* default_base(int ftp)
{
    switch (ftp)
    {
        case efTRX: case efTRO: case efXTC: case efCOMPRESSED: case efTRR: case efTRN: return "traj";
        case efTPR: case efTPS: return "topol";
        case efNDX: return "index";
        default: return "out";
    }
}

bool has_extension(const std::string& s)
{
    size_t slash = s.find_last_of('/');
    size_t dot   = s.find_last_of('.');
    return dot != std::string::npos && (slash == std::string::npos || dot > slash) && dot + 1 < s.size();
}

std::string with_ext(const std::string& name, int ftp)
{
    return has_extension(name) ? name : name + default_ext(ftp);
}

void print_help(int nfile, t_filenm fnm[], int npargs, t_pargs* pa, int ndesc, const char** desc)
{
    std::printf("SYNOPSIS\n\n%s", g_program.c_str());
    for (int i = 0; i < nfile; ++i) std::printf(" [%s <file>]", fnm[i].opt);
    for (int i = 0; i < npargs; ++i) std::printf(" [%s ...]", pa[i].option);
    std::printf("\n\nDESCRIPTION\n\n");
    for (int i = 0; i < ndesc; ++i) std::printf("%s\n", desc[i]);
    std::printf("\nOPTIONS\n\n");
    for (int i = 0; i < nfile; ++i)
        std::printf(" %-8s %-20s %s\n", fnm[i].opt, fnm[i].filenames.empty() ? "" : fnm[i].filenames[0].c_str(),
                    (fnm[i].flag & ffOPT) ? "(Opt.)" : "");
    for (int i = 0; i < npargs; ++i) std::printf(" %-12s %s\n", pa[i].option, pa[i].desc ? pa[i].desc : "");
    std::printf(" %-12s %s\n %-12s %s\n %-12s %s\n", "-b", "first frame time (ps)", "-e", "last frame time (ps)", "-dt",
                "only use frames with t MOD dt == first time (ps)");
}

[[noreturn]] void usage_error(const char* what, const char* arg)
{
    std::fprintf(stderr, "\nError in user input:\n%s '%s'\n", what, arg);
    std::exit(1);
}

double to_double(const char* opt, const char* s)
{
    char*  end = nullptr;
    double v   = std::strtod(s, &end);
    if (end == s || *end != '\0') usage_error("Invalid value for option", opt);
    return v;
}
} // namespace

gmx_bool parse_common_args(int* argc, char* argv[], unsigned long Flags, int nfile, t_filenm fnm[], int npargs,
                           t_pargs* pa, int ndesc, const char** desc, int /*nbugs*/, const char** /*bugs*/,
                           gmx_output_env_t** oenv)
{
    if (*argc > 0 && argv[0]) g_program = argv[0];
    /* defaults for file options */
    for (int i = 0; i < nfile; ++i)
    {
        fnm[i].filenames.clear();
        std::string base = fnm[i].fn ? fnm[i].fn : default_base(fnm[i].ftp);
        fnm[i].filenames.push_back(with_ext(base, fnm[i].ftp));
        fnm[i].flag &= ~static_cast<unsigned long>(ffSET);
    }
    bool want_help = false;
    for (int a = 1; a < *argc; ++a)
    {
        const char* arg = argv[a];
        if (!std::strcmp(arg, "-h") || !std::strcmp(arg, "-help") || !std::strcmp(arg, "--help"))
        {
            want_help = true;
            continue;
        }
        bool matched = false;
        for (int i = 0; i < nfile && !matched; ++i)
        {
            if (!std::strcmp(arg, fnm[i].opt))
            {
                matched = true;
                fnm[i].flag |= ffSET;
                if (a + 1 < *argc && argv[a + 1][0] != '-')
                {
                    fnm[i].filenames[0] = with_ext(argv[++a], fnm[i].ftp);
                }
            }
        }
        for (int i = 0; i < npargs && !matched; ++i)
        {
            t_pargs& p = pa[i];
            if (p.type == etBOOL)
            {
                if (!std::strcmp(arg, p.option)) { *p.u.b = TRUE; p.bSet = TRUE; matched = true; }
                else if (!std::strncmp(arg, "-no", 3) && !std::strcmp(arg + 3, p.option + 1)) { *p.u.b = FALSE; p.bSet = TRUE; matched = true; }
                continue;
            }
            if (std::strcmp(arg, p.option)) continue;
            matched = true;
            p.bSet  = TRUE;
            int need = (p.type == etRVEC) ? 3 : 1;
            if (a + need >= *argc + 0 && a + need > *argc - 1 + 0)
            {
                if (a + need > *argc - 1) usage_error("Missing value for option", arg);
            }
            switch (p.type)
            {
                case etINT: *p.u.i = static_cast<int>(std::lround(to_double(arg, argv[++a]))); break;
                case etINT64: *p.u.is = static_cast<int64_t>(std::llround(to_double(arg, argv[++a]))); break;
                case etREAL:
                case etTIME: *p.u.r = static_cast<real>(to_double(arg, argv[++a])); break;
                case etSTR: *p.u.c = argv[++a]; break;
                case etENUM: p.u.c[0] = argv[++a]; break;
                case etRVEC:
                    for (int k = 0; k < 3; ++k) (*p.u.rv)[k] = static_cast<real>(to_double(arg, argv[++a]));
                    break;
                default: usage_error("Unsupported option type for", arg);
            }
        }
        if (!matched && (Flags & PCA_CAN_TIME))
        {
            if (!std::strcmp(arg, "-b") && a + 1 < *argc) { g_time.has_b = true; g_time.b = to_double(arg, argv[++a]); matched = true; }
            else if (!std::strcmp(arg, "-e") && a + 1 < *argc) { g_time.has_e = true; g_time.e = to_double(arg, argv[++a]); matched = true; }
            else if (!std::strcmp(arg, "-dt") && a + 1 < *argc) { g_time.has_dt = true; g_time.dt = to_double(arg, argv[++a]); matched = true; }
        }
        if (!matched) usage_error("Unknown command-line option", arg);
    }
    if (want_help)
    {
        print_help(nfile, fnm, npargs, pa, ndesc, desc);
        return FALSE;
    }
    /* required input files must exist */
    for (int i = 0; i < nfile; ++i)
    {
        if ((fnm[i].flag & ffREAD) && (!(fnm[i].flag & ffOPT) || (fnm[i].flag & ffSET)))
        {
            FILE* fp = std::fopen(fnm[i].filenames[0].c_str(), "rb");
            if (!fp) usage_error("Input file does not exist or is not accessible:", fnm[i].filenames[0].c_str());
            std::fclose(fp);
        }
    }
    if (oenv)
    {
        *oenv            = new gmx_output_env_t;
        (*oenv)->program = g_program;
    }
    return TRUE;
}

static const t_filenm* find_opt(const char* opt, int nfile, const t_filenm fnm[])
{
    for (int i = 0; i < nfile; ++i)
        if (!std::strcmp(opt, fnm[i].opt)) return &fnm[i];
    return nullptr;
}
static const t_filenm* find_ftp(int ftp, int nfile, const t_filenm fnm[])
{
    for (int i = 0; i < nfile; ++i)
        if (fnm[i].ftp == ftp) return &fnm[i];
    return nullptr;
}
static bool is_active(const t_filenm* f) { return f && (!(f->flag & ffOPT) || (f->flag & ffSET)); }

const char* opt2fn(const char* opt, int nfile, const t_filenm fnm[])
{
    const t_filenm* f = find_opt(opt, nfile, fnm);
    if (!f || f->filenames.empty()) gmx_fatal(FARGS, "opt2fn: no file for option %s", opt);
    return f->filenames[0].c_str();
}
const char* opt2fn_null(const char* opt, int nfile, const t_filenm fnm[])
{
    const t_filenm* f = find_opt(opt, nfile, fnm);
    return (is_active(f) && !f->filenames.empty()) ? f->filenames[0].c_str() : nullptr;
}
const char* ftp2fn(int ftp, int nfile, const t_filenm fnm[])
{
    const t_filenm* f = find_ftp(ftp, nfile, fnm);
    if (!f || f->filenames.empty()) gmx_fatal(FARGS, "ftp2fn: no file of type %d", ftp);
    return f->filenames[0].c_str();
}
const char* ftp2fn_null(int ftp, int nfile, const t_filenm fnm[])
{
    const t_filenm* f = find_ftp(ftp, nfile, fnm);
    return (is_active(f) && !f->filenames.empty()) ? f->filenames[0].c_str() : nullptr;
}
gmx_bool opt2bSet(const char* opt, int nfile, const t_filenm fnm[])
{
    const t_filenm* f = find_opt(opt, nfile, fnm);
    return (f && (f->flag & ffSET)) ? TRUE : FALSE;
}
gmx_bool opt2parg_bSet(const char* opt, int nargs, const t_pargs pa[])
{
    for (int i = 0; i < nargs; ++i)
        if (!std::strcmp(opt, pa[i].option)) return pa[i].bSet;
    return FALSE;
}

int gmx_run_cmain(int argc, char* argv[], int (*mainFunction)(int, char*[])) { return mainFunction(argc, argv); }

void gmx_fatal(int /*fatal_errno*/, const char* file, int line, const char* fmt, ...)
{
    std::fflush(stdout);
    std::fprintf(stderr, "\n-------------------------------------------------------\nProgram:     %s\nSource file: %s (line %d)\n\nFatal error:\n",
                 g_program.c_str(), file, line);
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(stderr, fmt, ap);
    va_end(ap);
    std::fprintf(stderr, "\n-------------------------------------------------------\n");
    std::exit(1);
}

FILE* gmx_ffopen(const char* file, const char* mode)
{
    FILE* fp = std::fopen(file, mode);
    if (!fp) gmx_fatal(FARGS, "Cannot open file %s (mode %s)", file, mode);
    return fp;
}
int gmx_ffclose(FILE* fp) { return std::fclose(fp); }

/* ---- topology (.b2p) ---------------------------------------------------------------- */
namespace
{
struct TopHeader
{
    char    magic[8];
    int32_t natoms, nmols, ntypes, reserved;
};
struct TopAtom
{
    float   m, q;
    int32_t resind, type;
    char    name[8];
};
template <typename T>
void read_exact(FILE* fp, T* dst, size_t n, const char* fn)
{
    if (n && std::fread(dst, sizeof(T), n, fp) != n) gmx_fatal(FARGS, "Unexpected end of file in %s", fn);
}
} // namespace

int read_tpx_top(const char* fn, t_inputrec* ir, matrix box, int* natoms, rvec* /*x*/, rvec* /*v*/, t_topology* top)
{
    FILE*     fp = gmx_ffopen(fn, "rb");
    TopHeader h;
    read_exact(fp, &h, 1, fn);
    if (std::memcmp(h.magic, "B2ATOP01", 8) != 0) gmx_fatal(FARGS, "%s is not a gmxstub topology (bad magic)", fn);
    std::vector<TopAtom> at(h.natoms);
    read_exact(fp, at.data(), at.size(), fn);
    std::memset(top, 0, sizeof(*top));
    top->atoms.nr = h.natoms;
    snew(top->atoms.atom, h.natoms);
    snew(top->atoms.atomname, h.natoms);
    int nres = 0;
    for (int i = 0; i < h.natoms; ++i)
    {
        t_atom& a = top->atoms.atom[i];
        a.m = a.mB = at[i].m;
        a.q = a.qB = at[i].q;
        a.type = a.typeB = static_cast<unsigned short>(at[i].type);
        a.resind         = at[i].resind;
        if (a.resind + 1 > nres) nres = a.resind + 1;
        char** slot = nullptr;
        snew(slot, 1);
        char* nm = nullptr;
        snew(nm, 9);
        std::memcpy(nm, at[i].name, 8);
        *slot                  = nm;
        top->atoms.atomname[i] = slot;
    }
    top->atoms.nres      = nres;
    top->atoms.haveMass  = TRUE;
    top->atoms.haveCharge = TRUE;
    top->atoms.haveType  = TRUE;
    top->mols.nr         = h.nmols;
    snew(top->mols.index, h.nmols + 1);
    {
        std::vector<int32_t> mi(h.nmols + 1);
        read_exact(fp, mi.data(), mi.size(), fn);
        for (int i = 0; i <= h.nmols; ++i) top->mols.index[i] = mi[i];
    }
    top->mols.nalloc_index = h.nmols + 1;
    top->atomtypes.nr      = h.ntypes;
    top->idef.ntypes       = h.ntypes * h.ntypes;
    top->idef.atnr         = h.ntypes;
    snew(top->idef.iparams, static_cast<size_t>(h.ntypes) * h.ntypes);
    {
        std::vector<float> lj(static_cast<size_t>(h.ntypes) * h.ntypes * 2);
        read_exact(fp, lj.data(), lj.size(), fn);
        for (size_t i = 0; i < lj.size() / 2; ++i)
        {
            top->idef.iparams[i].lj.c6  = lj[2 * i];
            top->idef.iparams[i].lj.c12 = lj[2 * i + 1];
        }
    }
    std::fclose(fp);
    if (natoms) *natoms = h.natoms;
    if (ir) { std::memset(ir, 0, sizeof(*ir)); ir->ePBC = 0; }
    if (box) std::memset(box, 0, sizeof(matrix));
    return 0; /* epbcXYZ */
}

/* ---- index groups (.ndx) ------------------------------------------------------------ */
namespace
{
struct NdxGroup
{
    std::string      name;
    std::vector<int> atoms;
};

std::vector<NdxGroup> load_ndx(const char* fn)
{
    std::vector<NdxGroup> groups;
    FILE*                 fp = gmx_ffopen(fn, "r");
    std::string           text;
    char                  buf[1 << 16];
    size_t                n;
    while ((n = std::fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
    std::fclose(fp);
    const char* p   = text.c_str();
    const char* end = p + text.size();
    while (p < end)
    {
        while (p < end && std::isspace(static_cast<unsigned char>(*p))) ++p;
        if (p >= end) break;
        if (*p == '[')
        {
            const char* q = p + 1;
            while (q < end && *q != ']') ++q;
            std::string name(p + 1, q);
            size_t      a = name.find_first_not_of(" \t");
            size_t      b = name.find_last_not_of(" \t");
            groups.push_back({ a == std::string::npos ? std::string() : name.substr(a, b - a + 1), {} });
            p = (q < end) ? q + 1 : end;
        }
        else if (*p == ';')
        {
            while (p < end && *p != '\n') ++p;
        }
        else
        {
            char* e = nullptr;
            long  v = std::strtol(p, &e, 10);
            if (e == p) gmx_fatal(FARGS, "Invalid token in index file %s", fn);
            if (groups.empty()) gmx_fatal(FARGS, "Index file %s: atom number before first group header", fn);
            groups.back().atoms.push_back(static_cast<int>(v - 1)); /* .ndx is 1-based */
            p = e;
        }
    }
    return groups;
}

void select_groups(const std::vector<NdxGroup>& groups, int ngrps, int isize[], int* index[], char* grpnames[])
{
    for (size_t g = 0; g < groups.size(); ++g)
        std::printf("Group %5zu (%15s) has %5zu elements\n", g, groups[g].name.c_str(), groups[g].atoms.size());
    for (int s = 0; s < ngrps; ++s)
    {
        int choice = -1;
        if (groups.size() == 1) choice = 0;
        else
        {
            std::printf("Select a group: ");
            std::fflush(stdout);
            char tok[256];
            if (std::scanf("%255s", tok) == 1)
            {
                char* e = nullptr;
                long  v = std::strtol(tok, &e, 10);
                if (e != tok && *e == '\0') choice = static_cast<int>(v);
                else
                    for (size_t g = 0; g < groups.size(); ++g)
                        if (groups[g].name == tok) choice = static_cast<int>(g);
            }
            else
            {
                choice = s; /* no stdin: take groups in file order */
            }
        }
        if (choice < 0 || choice >= static_cast<int>(groups.size())) gmx_fatal(FARGS, "Invalid group selection");
        const NdxGroup& G = groups[choice];
        std::printf("Selected %d: '%s'\n", choice, G.name.c_str());
        isize[s] = static_cast<int>(G.atoms.size());
        snew(index[s], G.atoms.size());
        for (size_t i = 0; i < G.atoms.size(); ++i) index[s][i] = G.atoms[i];
        char* nm = nullptr;
        snew(nm, G.name.size() + 1);
        std::memcpy(nm, G.name.c_str(), G.name.size());
        grpnames[s] = nm;
    }
}
} // namespace

void rd_index(const char* statfile, int ngrps, int isize[], int* index[], char* grpnames[])
{
    if (!statfile) gmx_fatal(FARGS, "No index file specified");
    select_groups(load_ndx(statfile), ngrps, isize, index, grpnames);
}

void get_index(const t_atoms* atoms, const char* fnm, int ngrps, int isize[], int* index[], char* grpnames[])
{
    if (fnm)
    {
        rd_index(fnm, ngrps, isize, index, grpnames);
        return;
    }
    /* default groups when no index file is given: the whole system only */
    std::vector<NdxGroup> groups(1);
    groups[0].name = "System";
    groups[0].atoms.resize(atoms ? atoms->nr : 0);
    for (size_t i = 0; i < groups[0].atoms.size(); ++i) groups[0].atoms[i] = static_cast<int>(i);
    select_groups(groups, ngrps, isize, index, grpnames);
}

/* ---- trajectory (.b2t) -------------------------------------------------------------- */
struct t_trxstatus
{
    FILE*   fp      = nullptr;
    int     natoms  = 0;
    int     nframes = 0;
    int     fflags  = 0; /* bit0 x, bit1 v, bit2 f as stored in the file */
    int     next    = 0; /* next frame index in file */
    int     nread   = 0;
    int     want    = 0; /* TRX_* flags requested */
    bool    have_t0 = false;
    double  t0      = 0;
    std::string fn;
    std::vector<float> scratch;
    int     fmt     = 0; /* 0: B2ATRJ01, 1: XTC, 2: TRR (xdrtraj.hpp) */
    std::vector<unsigned char> blob;
    /* XTC files are mapped read-only: frames are decoded straight out of the page cache (no read() copy, no per-frame
     * buffer), and the batch reader's decode threads fault the pages in themselves.  map == nullptr: FILE* path. */
    const unsigned char* map = nullptr;
    size_t  map_size = 0, mpos = 0;
    int     repeat  = 1;  /* raw format: GMXSTUB_REPEAT=n replays the file n times (benchmarks: a long trajectory without the disk space) */
    int     pass    = 0;
    double  t_first = 0, t_last = 0;
    const unsigned char* cur = nullptr; /* coordinate block of the frame xtc_next() delivered last ... */
    size_t  cur_avail = 0;              /* ... and the readable bytes from there on */
};

namespace
{
struct TrjHeader
{
    char    magic[8];
    int32_t natoms, nframes, flags, reserved;
};

size_t frame_floats(const t_trxstatus* st)
{
    size_t per = 0;
    if (st->fflags & 1) per += 3;
    if (st->fflags & 2) per += 3;
    if (st->fflags & 4) per += 3;
    return 10 + per * static_cast<size_t>(st->natoms);
}

bool time_selected(t_trxstatus* st, double t, bool* past_end)
{
    *past_end = false;
    if (g_time.has_e && t > g_time.e + 1e-9) { *past_end = true; return false; }
    if (g_time.has_b && t < g_time.b - 1e-9) return false;
    if (g_time.has_dt && g_time.dt > 0)
    {
        if (!st->have_t0) return true;
        double k = (t - st->t0) / g_time.dt;
        if (std::fabs(k - std::round(k)) > 1e-6) return false;
    }
    return true;
}

/* Reads the next selected frame header (time, box) into hdr[10]; leaves the file positioned at x. */
bool next_selected(t_trxstatus* st, float hdr[10])
{
    for (;;)
    {
        if (st->next >= st->nframes)
        {
            if (st->pass + 1 >= st->repeat || st->nframes <= 0) return false;
            st->pass++;                                   /* GMXSTUB_REPEAT: back to the first frame, times keep growing */
            st->next = 0;
            std::fseek(st->fp, static_cast<long>(sizeof(TrjHeader)), SEEK_SET);
        }
        if (std::fread(hdr, sizeof(float), 10, st->fp) != 10) return false;
        if (st->pass == 0) { if (st->next == 0) st->t_first = hdr[0]; st->t_last = hdr[0]; }
        else hdr[0] += (float)(st->pass * (st->t_last - st->t_first + (st->nframes > 1 ? (st->t_last - st->t_first) / (st->nframes - 1) : 1.0)));
        bool past = false;
        bool sel  = time_selected(st, hdr[0], &past);
        if (past) return false;
        st->next++;
        if (sel)
        {
            if (!st->have_t0) { st->have_t0 = true; st->t0 = hdr[0]; }
            return true;
        }
        std::fseek(st->fp, static_cast<long>((frame_floats(st) - 10) * sizeof(float)), SEEK_CUR);
    }
    return false;
}

/* ---- real GROMACS wire formats (XTC / TRR), decoded by xdrtraj.hpp ------------------------------------------------ */
bool read_bytes(FILE* fp, void* dst, size_t n) { return std::fread(dst, 1, n, fp) == n; }

/* XTC: positions the file after the coordinate block of the next selected frame; the block is at st->cur (inside the
 * mapping, or in st->blob on the FILE* path).  hdr receives time and box.  Returns false at end of file.               */
bool xtc_next(t_trxstatus* st, xdrtraj::XtcHeader* h)
{
    while (st->map)
    {
        const size_t hb = xdrtraj::kXtcHeaderBytes;
        if (st->mpos + hb > st->map_size) return false;
        const unsigned char* p = st->map + st->mpos;
        if (!xdrtraj::xtc_parse_header(p, h)) gmx_fatal(FARGS, "%s: bad XTC frame header (magic)", st->fn.c_str());
        if (h->natoms != st->natoms) gmx_fatal(FARGS, "%s: number of atoms changes between frames", st->fn.c_str());
        const size_t npre = h->natoms <= 9 ? 4 : 40;
        if (st->mpos + hb + npre > st->map_size) return false;
        const size_t total = xdrtraj::xtc_coord_block_bytes(p + hb, h->natoms);
        bool         past  = false;
        const bool   sel   = time_selected(st, h->time, &past);
        if (past) return false;
        if (st->mpos + hb + total > st->map_size) gmx_fatal(FARGS, "%s: truncated XTC frame", st->fn.c_str());
        st->next++;
        st->cur       = p + hb;
        st->cur_avail = st->map_size - (st->mpos + hb);
        st->mpos += hb + total;
        if (!sel) continue;
        if (!st->have_t0) { st->have_t0 = true; st->t0 = h->time; }
        return true;
    }
    for (;;)
    {
        unsigned char hb[xdrtraj::kXtcHeaderBytes];
        if (!read_bytes(st->fp, hb, sizeof(hb))) return false;
        if (!xdrtraj::xtc_parse_header(hb, h)) gmx_fatal(FARGS, "%s: bad XTC frame header (magic)", st->fn.c_str());
        if (h->natoms != st->natoms) gmx_fatal(FARGS, "%s: number of atoms changes between frames", st->fn.c_str());
        unsigned char pre[40];
        const size_t  npre = h->natoms <= 9 ? 4 : 40;
        if (!read_bytes(st->fp, pre, npre)) return false;
        const size_t total = xdrtraj::xtc_coord_block_bytes(pre, h->natoms);
        bool         past  = false;
        const bool   sel   = time_selected(st, h->time, &past);
        if (past) return false;
        st->next++;
        if (!sel) { std::fseek(st->fp, static_cast<long>(total - npre), SEEK_CUR); continue; }
        if (!st->have_t0) { st->have_t0 = true; st->t0 = h->time; }
        st->blob.resize(total);
        std::memcpy(st->blob.data(), pre, npre);
        if (!read_bytes(st->fp, st->blob.data() + npre, total - npre)) gmx_fatal(FARGS, "%s: truncated XTC frame", st->fn.c_str());
        st->cur = st->blob.data(); st->cur_avail = st->blob.size();
        return true;
    }
}

/* TRR: the body (everything after lambda) of the next selected frame is at st->cur */
bool trr_next(t_trxstatus* st, xdrtraj::TrrHeader* h)
{
    while (st->map)
    {
        if (st->mpos + 96 > st->map_size) return false;
        const unsigned char* p = st->map + st->mpos;
        if (!xdrtraj::trr_parse_header(p, st->map_size - st->mpos, h)) gmx_fatal(FARGS, "%s: bad TRR frame header", st->fn.c_str());
        if (h->natoms != st->natoms) gmx_fatal(FARGS, "%s: number of atoms changes between frames", st->fn.c_str());
        const size_t body = xdrtraj::trr_body_bytes(*h);
        bool         past = false;
        const bool   sel  = time_selected(st, h->t, &past);
        if (past) return false;
        if (st->mpos + h->header_bytes + body > st->map_size) gmx_fatal(FARGS, "%s: truncated TRR frame", st->fn.c_str());
        st->next++;
        st->cur       = p + h->header_bytes;
        st->cur_avail = body;
        st->mpos += h->header_bytes + body;
        if (!sel) continue;
        if (!st->have_t0) { st->have_t0 = true; st->t0 = h->t; }
        return true;
    }
    for (;;)
    {
        unsigned char hb[128];
        const long    pos = std::ftell(st->fp);
        const size_t  got = std::fread(hb, 1, sizeof(hb), st->fp);
        if (got < 96) return false;
        if (!xdrtraj::trr_parse_header(hb, got, h)) gmx_fatal(FARGS, "%s: bad TRR frame header", st->fn.c_str());
        if (h->natoms != st->natoms) gmx_fatal(FARGS, "%s: number of atoms changes between frames", st->fn.c_str());
        const size_t body = xdrtraj::trr_body_bytes(*h);
        bool         past = false;
        const bool   sel  = time_selected(st, h->t, &past);
        if (past) return false;
        st->next++;
        if (!sel) { std::fseek(st->fp, pos + static_cast<long>(h->header_bytes + body), SEEK_SET); continue; }
        if (!st->have_t0) { st->have_t0 = true; st->t0 = h->t; }
        std::fseek(st->fp, pos + static_cast<long>(h->header_bytes), SEEK_SET);
        st->blob.resize(body);
        if (!read_bytes(st->fp, st->blob.data(), body)) gmx_fatal(FARGS, "%s: truncated TRR frame", st->fn.c_str());
        st->cur = st->blob.data(); st->cur_avail = body;
        return true;
    }
}

/* XTC / TRR files are read through a read-only mapping when they are regular files (GMXSTUB_NO_MMAP=1: FILE* path) */
void map_trajectory(t_trxstatus* st)
{
    struct stat sb;
    if (std::getenv("GMXSTUB_NO_MMAP") || fstat(fileno(st->fp), &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size <= 0) return;
    void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fileno(st->fp), 0);
    if (m == MAP_FAILED) return;
    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
    st->map = static_cast<const unsigned char*>(m); st->map_size = (size_t)sb.st_size;
}

t_trxstatus* open_trx(const char* fn)
{
    t_trxstatus* st = new t_trxstatus;
    st->fp          = gmx_ffopen(fn, "rb");
    st->fn          = fn;
    std::setvbuf(st->fp, nullptr, _IOFBF, 8 << 20);
    unsigned char first[128];
    const size_t  got = std::fread(first, 1, sizeof(first), st->fp);
    std::fseek(st->fp, 0, SEEK_SET);
    if (got >= 8 && xdrtraj::be32(first) == (uint32_t)xdrtraj::kXtcMagic)
    {
        st->fmt = 1; st->natoms = (int)xdrtraj::be32(first + 4); st->nframes = -1; st->fflags = 1;
        map_trajectory(st);
        return st;
    }
    if (got >= 96 && xdrtraj::be32(first) == (uint32_t)xdrtraj::kTrrMagic)
    {
        xdrtraj::TrrHeader h;
        if (!xdrtraj::trr_parse_header(first, got, &h)) gmx_fatal(FARGS, "%s: bad TRR header", fn);
        st->fmt = 2; st->natoms = h.natoms; st->nframes = -1;
        st->fflags = (h.x_size ? 1 : 0) | (h.v_size ? 2 : 0) | (h.f_size ? 4 : 0);
        map_trajectory(st);
        return st;
    }
    TrjHeader h;
    read_exact(st->fp, &h, 1, fn);
    if (std::memcmp(h.magic, "B2ATRJ01", 8) != 0) gmx_fatal(FARGS, "%s is neither an XTC / TRR file nor a gmxstub trajectory (bad magic)", fn);
    st->natoms  = h.natoms;
    st->nframes = h.nframes;
    st->fflags  = h.flags;
    if (const char* e = std::getenv("GMXSTUB_REPEAT")) st->repeat = std::max(1, std::atoi(e));
    return st;
}

bool deliver_xdr(t_trxstatus* st, t_trxframe* fr)
{
    const size_t n3 = static_cast<size_t>(st->natoms) * 3;
    fr->natoms = st->natoms;
    fr->bStep = TRUE; fr->bTime = TRUE; fr->bBox = TRUE;
    fr->bX = fr->bV = fr->bF = FALSE;
    if (st->fmt == 1)
    {
        xdrtraj::XtcHeader h;
        if (!xtc_next(st, &h)) return false;
        fr->time = h.time; fr->step = h.step;
        for (int k = 0; k < 9; ++k) fr->box[k / 3][k % 3] = h.box[k];
        if ((st->want & (TRX_READ_X | TRX_NEED_X)) && fr->x)
        {
            if (!xdrtraj::xtc_decode_coords(st->cur, st->cur_avail, st->natoms, &fr->x[0][0]))
                gmx_fatal(FARGS, "%s: corrupt XTC coordinate block in frame %d", st->fn.c_str(), st->next - 1);
            fr->bX = TRUE;
        }
    }
    else
    {
        xdrtraj::TrrHeader h;
        if (!trr_next(st, &h)) return false;
        fr->time = (float)h.t; fr->step = h.step;
        const unsigned char* p  = st->cur + h.ir_size + h.e_size;
        const size_t         rs = h.dbl ? 8 : 4;
        if (h.box_size) { float b[9]; xdrtraj::trr_read_reals(p, h.dbl, 9, b); for (int k = 0; k < 9; ++k) fr->box[k / 3][k % 3] = b[k]; }
        p += h.box_size + h.vir_size + h.pres_size + h.top_size + h.sym_size;
        auto rd = [&](int size, int want_bit, rvec* dst, gmx_bool* have) {
            if (!size) return;
            if ((st->want & want_bit) && dst) { xdrtraj::trr_read_reals(p, h.dbl, n3, &dst[0][0]); *have = TRUE; }
            p += n3 * rs;
        };
        rd(h.x_size, TRX_READ_X | TRX_NEED_X, fr->x, &fr->bX);
        rd(h.v_size, TRX_READ_V | TRX_NEED_V, fr->v, &fr->bV);
        rd(h.f_size, TRX_READ_F | TRX_NEED_F, fr->f, &fr->bF);
    }
    st->nread++;
    return true;
}

bool deliver(t_trxstatus* st, t_trxframe* fr)
{
    if (st->fmt != 0) return deliver_xdr(st, fr);
    float hdr[10];
    if (!next_selected(st, hdr)) return false;
    fr->time  = hdr[0];
    fr->bTime = TRUE;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) fr->box[a][b] = hdr[1 + 3 * a + b];
    fr->bBox    = TRUE;
    fr->natoms  = st->natoms;
    fr->step    = st->next - 1;
    fr->bStep   = TRUE;
    const size_t n3 = static_cast<size_t>(st->natoms) * 3;
    auto         rd = [&](int bit, int want_bit, rvec* dst, gmx_bool* have) {
        if (!(st->fflags & bit)) { *have = FALSE; return; }
        if ((st->want & want_bit) && dst) { read_exact(st->fp, &dst[0][0], n3, st->fn.c_str()); *have = TRUE; }
        else { std::fseek(st->fp, static_cast<long>(n3 * sizeof(float)), SEEK_CUR); *have = FALSE; }
    };
    rd(1, TRX_READ_X | TRX_NEED_X, fr->x, &fr->bX);
    rd(2, TRX_READ_V | TRX_NEED_V, fr->v, &fr->bV);
    rd(4, TRX_READ_F | TRX_NEED_F, fr->f, &fr->bF);
    st->nread++;
    return true;
}
} // namespace

namespace
{
/* statuses handed out and not yet closed: a caller that passes one of them to read_first_frame again (two-pass programs restart
 * the trajectory that way) gets the old one closed instead of leaked -- file handle and mapping; an uninitialised pointer
 * (GROMACS callers declare `t_trxstatus* status;`) is simply not found here */
std::vector<t_trxstatus*>& open_statuses()
{
    static std::vector<t_trxstatus*> v;
    return v;
}
} // namespace

bool read_first_frame(const gmx_output_env_t* /*oenv*/, t_trxstatus** status, const char* fn, t_trxframe* fr, int flags)
{
    {
        auto& v = open_statuses();
        for (size_t k = 0; k < v.size(); ++k)
            if (status && v[k] == *status) { close_trx(*status); break; }   // close_trx takes it out of the list
    }
    t_trxstatus* st = open_trx(fn);
    open_statuses().push_back(st);
    st->want        = flags;
    *status         = st;
    rvec* keep_x = fr->x; rvec* keep_v = fr->v; rvec* keep_f = fr->f;
    std::memset(fr, 0, sizeof(*fr));
    fr->x = keep_x; fr->v = keep_v; fr->f = keep_f;
    if ((flags & (TRX_READ_X | TRX_NEED_X)) && (st->fflags & 1)) { if (fr->x) sfree(fr->x); snew(fr->x, st->natoms); }
    if ((flags & (TRX_READ_V | TRX_NEED_V)) && (st->fflags & 2)) { if (fr->v) sfree(fr->v); snew(fr->v, st->natoms); }
    if ((flags & (TRX_READ_F | TRX_NEED_F)) && (st->fflags & 4)) { if (fr->f) sfree(fr->f); snew(fr->f, st->natoms); }
    if ((flags & TRX_NEED_X) && !(st->fflags & 1)) gmx_fatal(FARGS, "Trajectory %s has no coordinates", fn);
    if ((flags & TRX_NEED_V) && !(st->fflags & 2)) gmx_fatal(FARGS, "Trajectory %s has no velocities", fn);
    if ((flags & TRX_NEED_F) && !(st->fflags & 4)) gmx_fatal(FARGS, "Trajectory %s has no forces", fn);
    return deliver(st, fr);
}

bool read_next_frame(const gmx_output_env_t* /*oenv*/, t_trxstatus* status, t_trxframe* fr) { return deliver(status, fr); }

void close_trx(t_trxstatus* status)
{
    if (!status) return;
    {
        auto& v = open_statuses();
        for (size_t k = 0; k < v.size(); ++k)
            if (v[k] == status) { v.erase(v.begin() + (long)k); break; }
    }
    if (status->map) munmap(const_cast<unsigned char*>(status->map), status->map_size);
    if (status->fp) std::fclose(status->fp);
    delete status;
}

int nframes_read(t_trxstatus* status) { return status ? status->nread : 0; }

int gmxstub_natoms(const t_trxstatus* status) { return status ? status->natoms : 0; }

int gmxstub_read_frames(t_trxstatus* st, int max_frames, float* x_out, float* box_out, float* time_out)
{
    if (!st || !(st->fflags & 1)) return 0;
    const size_t n3   = static_cast<size_t>(st->natoms) * 3;
    if (st->fmt == 1)
    {
        /* XTC frames are independent: slice K compressed frames out of the file, then decode them on K host threads
         * straight into the caller's (pinned) batch */
        // The reader (this thread) slices the frames off the file one after the other; the decode threads pick a frame up as
        // soon as it has been read, so the serial read of the batch overlaps its decode.
        std::vector<std::vector<unsigned char>> blobs(st->map ? 0 : max_frames);   // FILE* path only: owned copies
        std::vector<const unsigned char*>       src(std::max(max_frames, 1), nullptr);
        std::vector<size_t>                     src_avail(std::max(max_frames, 1), 0);
        std::vector<int>                        bad(std::max(max_frames, 1), 0);
        int nthreads = std::min<int>(max_frames, std::max(1u, std::min(64u, std::thread::hardware_concurrency())));
        if (const char* e = std::getenv("GMXSTUB_DECODE_THREADS")) nthreads = std::max(1, std::min(max_frames, std::atoi(e)));
        std::mutex              mu;
        std::condition_variable cv;
        int                     ready = 0, claimed = 0;   // frames read so far / frames handed to a decode thread
        bool                    done = false;             // the reader has stopped (batch full or end of file)
        auto work = [&]() {
            for (;;)
            {
                int k;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return claimed < ready || done; });
                    if (claimed >= ready) return;         // done, and nothing left
                    k = claimed++;
                }
                if (!xdrtraj::xtc_decode_coords(src[k], src_avail[k], st->natoms, x_out + n3 * k)) bad[k] = 1;
            }
        };
        std::vector<std::thread> pool;
        if (nthreads > 1)
            for (int t = 0; t < nthreads; ++t) pool.emplace_back(work);
        int n = 0;
        while (n < max_frames)
        {
            xdrtraj::XtcHeader h;
            if (!xtc_next(st, &h)) break;
            time_out[n] = h.time;
            std::memcpy(box_out + 9 * (size_t)n, h.box, 9 * sizeof(float));
            if (st->map) { src[n] = st->cur; src_avail[n] = st->cur_avail; }
            else { blobs[n] = std::move(st->blob); src[n] = blobs[n].data(); src_avail[n] = blobs[n].size(); }
            st->nread++;
            {
                std::lock_guard<std::mutex> lk(mu);
                ready = ++n;
            }
            cv.notify_one();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            done = true;
        }
        cv.notify_all();
        if (nthreads > 1)
            for (auto& th : pool) th.join();
        else work();
        for (int k = 0; k < n; ++k)
            if (bad[k]) gmx_fatal(FARGS, "%s: corrupt XTC coordinate block", st->fn.c_str());
        return n;
    }
    if (st->fmt == 2 && st->map)
    {
        /* TRR frames in a mapped file: collect the headers of the batch (cheap), then convert the big-endian coordinate
         * blocks of all frames on several threads straight out of the page cache */
        struct Item { const unsigned char* x; bool dbl; };
        std::vector<Item> items;
        while ((int)items.size() < max_frames)
        {
            xdrtraj::TrrHeader h;
            if (!trr_next(st, &h)) break;
            const int k = (int)items.size();
            const unsigned char* p = st->cur + h.ir_size + h.e_size;
            float b[9] = { 0 };
            if (h.box_size) xdrtraj::trr_read_reals(p, h.dbl, 9, b);
            std::memcpy(box_out + 9 * static_cast<size_t>(k), b, sizeof(b));
            time_out[k] = (float)h.t;
            if (!h.x_size) gmx_fatal(FARGS, "%s: TRR frame without coordinates", st->fn.c_str());
            items.push_back({ p + h.box_size + h.vir_size + h.pres_size + h.top_size + h.sym_size, h.dbl });
            st->nread++;
        }
        const int n = (int)items.size();
        int nthreads = std::min<int>(n, std::max(1u, std::min(64u, std::thread::hardware_concurrency())));
        if (const char* e = std::getenv("GMXSTUB_DECODE_THREADS")) nthreads = std::max(1, std::min(n, std::atoi(e)));
        std::atomic<int> next{ 0 };
        auto work = [&]() {
            for (int k = next++; k < n; k = next++) xdrtraj::trr_read_reals(items[k].x, items[k].dbl, n3, x_out + n3 * k);
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
        if (n) work();
        for (auto& th : pool) th.join();
        return n;
    }
    if (st->fmt == 2)
    {
        int got = 0;
        t_trxframe fr{};
        const int  want = st->want;
        st->want        = TRX_READ_X;
        while (got < max_frames)
        {
            fr.x = reinterpret_cast<rvec*>(x_out + n3 * got);
            if (!deliver_xdr(st, &fr)) break;
            time_out[got] = fr.time;
            std::memcpy(box_out + 9 * static_cast<size_t>(got), &fr.box[0][0], 9 * sizeof(float));
            ++got;
        }
        st->want = want;
        return got;
    }
    /* raw format: walk the (tiny) frame headers sequentially, then fetch the coordinate blocks of the batch with one pread per
     * frame on several threads -- a single fread stream tops out near 3 GB/s, below what a few GPUs consume */
    const int          fd = fileno(st->fp);
    const off_t        fb = static_cast<off_t>(frame_floats(st) * sizeof(float));
    std::vector<off_t> at;
    while ((int)at.size() < max_frames)
    {
        /* the same walk as next_selected(), on file offsets: the buffered FILE* would refill megabytes per 40-byte header */
        if (st->next >= st->nframes)
        {
            if (st->pass + 1 >= st->repeat || st->nframes <= 0) break;
            st->pass++;
            st->next = 0;
        }
        float       hdr[10];
        const off_t off = static_cast<off_t>(sizeof(TrjHeader)) + static_cast<off_t>(st->next) * fb;
        if (pread(fd, hdr, sizeof(hdr), off) != (ssize_t)sizeof(hdr)) break;
        if (st->pass == 0) { if (st->next == 0) st->t_first = hdr[0]; st->t_last = hdr[0]; }
        else hdr[0] += (float)(st->pass * (st->t_last - st->t_first + (st->nframes > 1 ? (st->t_last - st->t_first) / (st->nframes - 1) : 1.0)));
        bool       past = false;
        const bool sel  = time_selected(st, hdr[0], &past);
        if (past) break;
        st->next++;
        if (!sel) continue;
        if (!st->have_t0) { st->have_t0 = true; st->t0 = hdr[0]; }
        const int k = (int)at.size();
        time_out[k] = hdr[0];
        std::memcpy(box_out + 9 * static_cast<size_t>(k), hdr + 1, 9 * sizeof(float));
        at.push_back(off + (off_t)sizeof(hdr));
        st->nread++;
    }
    std::fseek(st->fp, static_cast<long>(sizeof(TrjHeader) + static_cast<off_t>(st->next) * fb), SEEK_SET);   /* keep the stream in step */
    const int n  = (int)at.size();
    int nthreads = std::min<int>(n, std::max(1u, std::min(32u, std::thread::hardware_concurrency())));
    if (const char* e = std::getenv("GMXSTUB_DECODE_THREADS")) nthreads = std::max(1, std::min(n, std::atoi(e)));
    std::atomic<int> next{ 0 }, bad{ 0 };
    auto work = [&]() {
        for (int k = next++; k < n; k = next++)
        {
            char*  dst  = reinterpret_cast<char*>(x_out + n3 * k);
            size_t left = n3 * sizeof(float);
            off_t  off  = at[k];
            while (left)
            {
                const ssize_t r = pread(fd, dst, left, off);
                if (r <= 0) { bad = 1; break; }
                dst += r; off += r; left -= (size_t)r;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    if (n) work();
    for (auto& th : pool) th.join();
    if (bad) gmx_fatal(FARGS, "%s: short read", st->fn.c_str());
    return n;
}

/* ---- writers for the real formats (tests, tools) --------------------------------------------------------------- */
extern "C" int gmxstub_write_xtc(const char* path, int natoms, int nframes, const float* x, const float* box9, int nbox,
                                 const float* times, float precision)
{
    FILE* fp = std::fopen(path, "wb");
    if (!fp) return -1;
    for (int k = 0; k < nframes; ++k)
    {
        xdrtraj::Writer w;
        xdrtraj::xtc_write_frame(w, natoms, k, times ? times[k] : (float)k, box9 + 9 * static_cast<size_t>(std::min(k, nbox - 1)),
                                 x + static_cast<size_t>(natoms) * 3 * k, precision);
        std::fwrite(w.buf.data(), 1, w.buf.size(), fp);
    }
    std::fclose(fp);
    return 0;
}

extern "C" int gmxstub_write_trr(const char* path, int natoms, int nframes, const float* x, const float* v, const float* f,
                                 const float* box9, int nbox, const float* times)
{
    FILE* fp = std::fopen(path, "wb");
    if (!fp) return -1;
    const size_t n3 = static_cast<size_t>(natoms) * 3;
    for (int k = 0; k < nframes; ++k)
    {
        xdrtraj::Writer w;
        xdrtraj::trr_write_frame(w, natoms, k, times ? times[k] : (float)k, box9 + 9 * static_cast<size_t>(std::min(k, nbox - 1)),
                                 x ? x + n3 * k : nullptr, v ? v + n3 * k : nullptr, f ? f + n3 * k : nullptr);
        std::fwrite(w.buf.data(), 1, w.buf.size(), fp);
    }
    std::fclose(fp);
    return 0;
}

/* decode-only entry for tests: all frames of an XTC / TRR / stub file through read_first_frame / read_next_frame */
extern "C" int gmxstub_read_all(const char* path, int max_frames, int want_flags, float* x_out, float* v_out, float* f_out,
                                float* box_out, float* time_out, int* natoms_out)
{
    t_trxstatus* st = nullptr;
    t_trxframe   fr{};
    if (!read_first_frame(nullptr, &st, path, &fr, want_flags)) return 0;
    const size_t n3 = static_cast<size_t>(fr.natoms) * 3;
    if (natoms_out) *natoms_out = fr.natoms;
    int n = 0;
    do
    {
        if (n >= max_frames) break;
        if (x_out && fr.bX) std::memcpy(x_out + n3 * n, fr.x, n3 * sizeof(float));
        if (v_out && fr.bV) std::memcpy(v_out + n3 * n, fr.v, n3 * sizeof(float));
        if (f_out && fr.bF) std::memcpy(f_out + n3 * n, fr.f, n3 * sizeof(float));
        if (box_out) std::memcpy(box_out + 9 * static_cast<size_t>(n), &fr.box[0][0], 9 * sizeof(float));
        if (time_out) time_out[n] = fr.time;
        ++n;
    } while (read_next_frame(nullptr, st, &fr));
    close_trx(st);
    return n;
}

extern "C" int gmxstub_read_batches(const char* path, int batch, int max_frames, float* x_out, float* box_out, float* time_out)
{
    t_trxstatus* st = open_trx(path);
    st->want        = TRX_READ_X;
    const size_t n3 = static_cast<size_t>(st->natoms) * 3;
    int          n  = 0;
    while (n < max_frames)
    {
        const int got = gmxstub_read_frames(st, std::min(batch, max_frames - n), x_out + n3 * n, box_out + 9 * static_cast<size_t>(n), time_out + n);
        if (got <= 0) break;
        n += got;
    }
    close_trx(st);
    return n;
}
