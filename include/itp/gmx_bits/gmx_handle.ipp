// itp::GmxHandle -- B200 drop-in implementation.  Behaviour contract: reference
// include/itp/gmx_bits/gmx_handle.ipp (cited per method); arithmetic of the position loaders runs in libb2a.so.
#pragma once

#include <cstdlib>
#include <cstring>
#include <filesystem>

#include "gmx_handle.hpp"

namespace itp
{
    inline GmxHandle::GmxHandle(int argc, char** argv) :
        flags(TRX_READ_X), ngrps(1), nframe(0), argc(argc), argv(argv), status(nullptr), oenv(nullptr), ePBC(0)
    {
        grpname = nullptr; ngx = nullptr; top = nullptr; ir = nullptr; index = nullptr; fr = nullptr;
        dt = time = preTime = 0.0;
        Lbox[0] = Lbox[1] = Lbox[2] = 0.0;
        nthreads = 0;
        if (const char* e = std::getenv("B2A_BATCH")) batch_k_ = std::max(1, std::atoi(e));
    }

    inline GmxHandle::~GmxHandle()
    {
        if (ctx_) b2a_destroy(ctx_);
        if (bx_) b2a_pinned_free(bx_);
        // top / ir / index / fr are deliberately left alone: the reference never frees them either and programs
        // may hold pointers into them until exit (SURVEY.md 8b "Ownership").
    }

    inline void GmxHandle::b2aCheck(int rc)
    {
        // the C ABI returns codes; the reference aborts through gmx_fatal -- same observable behaviour
        if (rc != B2A_OK) gmx_fatal(FARGS, "%s", b2a_last_error());
    }

    // ---- init: command line, topology, index groups, molecule discovery (reference ipp:11-66) -------------------
    inline void GmxHandle::init(bool bFullMolecule)
    {
        fnm.push_back({ efTRX, "-f", nullptr, ffREAD });
        fnm.push_back({ efTPR, "-s", nullptr, ffREAD });
        fnm.push_back({ efNDX, "-n", nullptr, ffOPTRD });

        if (!parse_common_args(&argc, argv, PCA_CAN_VIEW | PCA_CAN_TIME, (int)fnm.size(), fnm.data(), (int)pa.size(),
                               pa.data(), (int)desc.size(), desc.data(), 0, nullptr, &oenv))
        {
            exit(0);
        }

        top = new t_topology;
        ir  = new t_inputrec;
        int natoms = 0;
        ePBC    = read_tpx_top(get_ftp2fn(efTPR), ir, nullptr, &natoms, nullptr, nullptr, top);
        natoms_ = natoms;

        snew(grpname, ngrps);
        snew(index, ngrps);
        snew(ngx, ngrps);
        snew(fr, 1);

        fmt::print("\n\033[31mSpecify {} group{} to analysis:\33[0m\n", ngrps, (ngrps > 1) ? "s" : "");
        get_index(&top->atoms, get_ftp2fn_null(efNDX), ngrps, ngx, index, grpname);

        if (!bFullMolecule) return;

        napm.resize(ngrps);
        nmol.resize(ngrps);
        charge.resize(ngrps);
        mass.resize(ngrps);
        for (int g = 0; g != ngrps; ++g)
        {
            fmt::print("\n\033[31m### Group {}: '{}': \033[0m\n", g, grpname[g]);
            napm[g] = get_natom_per_mol(g);
            nmol[g] = ngx[g] / napm[g];
            fmt::print("# Nr. of atoms in each molecule : {}\n", napm[g]);
            fmt::print("# Nr. of molecules in group: {}\n", nmol[g]);
            mass[g]   = get_mass(g);
            charge[g] = get_charge(g);
            fmt::print("# Print infomation of molecule:\n");
            fmt::print("\033[33m# nr.   name     mass   charge\033[0m\n");
            fmt::print("------------------------------\n");
            for (int j = 0; j != napm[g]; ++j)
                fmt::print("{:>5d}{:>7s}{:>9.3f}{:>9.3f}\n", j, *top->atoms.atomname[index[g][j]], mass[g][j], charge[g][j]);
            fmt::print("------------------------------\n");
            fmt::print("{:<12s}{:>9.3f}{:>9.3f}\n", "Total", mass[g].sum(), charge[g].sum());
        }
        fmt::print("\n");
    }

    // A group must be a run of whole, equally sized molecules.  Molecules are looked up through the residue index of
    // each molecule's first atom, exactly like the reference (ipp:419-448), including its two failure modes.
    inline int GmxHandle::get_natom_per_mol(int grp)
    {
        std::vector<int> sizes;
        int              pos = 0;
        while (pos < ngx[grp])
        {
            const int res   = top->atoms.atom[index[grp][pos]].resind;
            const int first = top->mols.index[res], last = top->mols.index[res + 1];
            for (int a = first; a != last; ++a, ++pos)
                if (pos >= ngx[grp] || index[grp][pos] != a)
                    gmx_fatal(FARGS, "The index group does not consist of whole molecules");
            sizes.push_back(last - first);
        }
        for (int s : sizes)
            if (s != sizes[0])
            {
                fmt::print(stderr, "\nError: atoms per molecule is not the same in the group!");
                std::exit(-1);
            }
        return sizes[0];
    }

    inline vecd GmxHandle::get_mass(int grp)   // template = first molecule of the group (ipp:450-458)
    {
        vecd m(napm[grp]);
        for (int j = 0; j != napm[grp]; ++j) m[j] = top->atoms.atom[index[grp][j]].m;
        return m;
    }

    inline vecd GmxHandle::get_charge(int grp)   // ipp:460-468
    {
        vecd q(napm[grp]);
        for (int j = 0; j != napm[grp]; ++j) q[j] = top->atoms.atom[index[grp][j]].q;
        return q;
    }

    inline vecd GmxHandle::weights(int grp, int com)   // ipp:160-174
    {
        if (com == 0) return mass[grp];
        if (com == 2) return charge[grp];
        vecd w(napm[grp]);
        w.fill(1);
        return w;
    }

    // ---- device context, groups, batches -------------------------------------------------------------------------
    inline void GmxHandle::b2aEnsure()
    {
        if (!ctx_)
        {
            int dev = 0;
            if (const char* e = std::getenv("B2A_DEVICE")) dev = std::atoi(e);
            b2aCheck(b2a_create(&ctx_, dev, natoms_, batch_k_));
        }
        if (!groups_up_ && napm.size() == ngrps)
        {
            for (int g = 0; g != ngrps; ++g)
            {
                b2aCheck(b2a_set_group(ctx_, g, index[g], ngx[g], napm[g], mass[g].data(), charge[g].data()));
                // hand over Eigen's own evaluation of atomCenter.sum() (ipp:195) so centres are bit-identical
                for (int com = 0; com < 3; ++com) b2aCheck(b2a_set_group_wsum(ctx_, g, com, weights(g, com).sum()));
            }
            groups_up_ = true;
        }
    }

    inline void GmxHandle::b2aPointFrame(int k)
    {
        const size_t n3 = (size_t)natoms_ * 3;
        *fr        = io_;
        fr->x      = io_.bX ? reinterpret_cast<rvec*>(bx_ + (size_t)k * n3) : nullptr;
        fr->v      = io_.bV ? reinterpret_cast<rvec*>(bv_.data() + (size_t)k * n3) : nullptr;
        fr->f      = io_.bF ? reinterpret_cast<rvec*>(bf_.data() + (size_t)k * n3) : nullptr;
        fr->time   = btime_[k];
        fr->natoms = natoms_;
        std::memcpy(&fr->box[0][0], bbox_.data() + 9 * (size_t)k, 9 * sizeof(float));
        cur_ = k;
    }

    // Decode up to K frames into the pinned batch and ship them with one asynchronous copy.
    inline bool GmxHandle::b2aFillBatch(bool first)
    {
        const size_t n3 = (size_t)natoms_ * 3;
        int          n  = 0;
        for (; n < batch_k_; ++n)
        {
            if (!(first && n == 0))
            {
                if (eof_ || !read_next_frame(oenv, status, &io_)) { eof_ = true; break; }
            }
            if (io_.bX) std::memcpy(bx_ + (size_t)n * n3, io_.x, n3 * sizeof(float));
            if (io_.bV) std::memcpy(bv_.data() + (size_t)n * n3, io_.v, n3 * sizeof(float));
            if (io_.bF) std::memcpy(bf_.data() + (size_t)n * n3, io_.f, n3 * sizeof(float));
            std::memcpy(bbox_.data() + 9 * (size_t)n, &io_.box[0][0], 9 * sizeof(float));
            btime_[n] = io_.time;
        }
        if (n == 0) return false;
        batch_n_ = n;
        ++serial_;
        if (io_.bX)
        {
            b2aEnsure();
            b2aCheck(b2a_submit_batch(ctx_, n, bx_, bbox_.data()));
        }
        return true;
    }

    // ---- frame stepping (reference ipp:68-95) ----------------------------------------------------------------------
    inline bool GmxHandle::readFirstFrame()
    {
        // a second call restarts the trajectory, like read_first_frame does (two-pass programs rely on it)
        io_ = t_trxframe{};
        eof_ = false;
        bool b = read_first_frame(oenv, &status, get_ftp2fn(efTRX), &io_, flags);
        if (b)
        {
            natoms_ = io_.natoms;
            const size_t n3 = (size_t)natoms_ * 3;
            if (!bx_)
            {
                bx_ = static_cast<float*>(b2a_pinned_alloc(n3 * sizeof(float) * batch_k_));
                if (!bx_) gmx_fatal(FARGS, "%s", b2a_last_error());
                bbox_.resize(9 * (size_t)batch_k_);
                btime_.resize(batch_k_);
            }
            if (io_.bV) bv_.resize(n3 * batch_k_);
            if (io_.bF) bf_.resize(n3 * batch_k_);
            b = b2aFillBatch(true);
            if (b) b2aPointFrame(0);
        }
        b && (nframe = 1);
        Lbox[XX] = fr->box[XX][XX];
        Lbox[YY] = fr->box[YY][YY];
        Lbox[ZZ] = fr->box[ZZ][ZZ];
        time = preTime = fr->time;
        dt = 0;
        return b;
    }

    inline bool GmxHandle::readNextFrame()
    {
        bool b;
        if (cur_ + 1 < batch_n_) { b2aPointFrame(cur_ + 1); b = true; }
        else
        {
            b = !eof_ && b2aFillBatch(false);
            if (b) b2aPointFrame(0);
        }
        b && (++nframe);
        Lbox[XX] = fr->box[XX][XX];   // updated even at end of file, as in the reference (ipp:85-87)
        Lbox[YY] = fr->box[YY][YY];
        Lbox[ZZ] = fr->box[ZZ][ZZ];
        if (b)
        {
            dt      = time - preTime;   // one frame late by construction (ipp:90-92)
            preTime = time;
            time    = fr->time;
        }
        return b;
    }

    inline bool GmxHandle::b2aReadNextBatch()
    {
        nframe += batch_n_ - 1 - cur_;   // the frames of this batch that were consumed as a block
        cur_ = batch_n_ - 1;
        return readNextFrame();
    }

    // ---- containers (ipp:97-113) -----------------------------------------------------------------------------------
    inline boxd GmxHandle::initPos(int grp)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        return boxd(nmol[grp], napm[grp]);
    }

    inline matd GmxHandle::initPosc(int grp)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        return matd(nmol[grp], 3);
    }

    // ---- position loaders: served by the GPU (K1b / K1), layouts are Eigen's ------------------------------------------
    inline void GmxHandle::loadPosition(boxd& pos, int grp)   // ipp:115-148
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        b2aEnsure();
        // boxd = column-major array of Array3d: element (i,j)[k] at (i + j*nmol)*3 + k -- written in place
        b2aCheck(b2a_positions(ctx_, grp, cur_, reinterpret_cast<double*>(pos.data())));
    }

    inline void GmxHandle::loadPositionCenter(matd& posc, int grp, int center)   // ipp:150-199
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        if (center < 0 || center > 2) gmx_fatal(FARGS, "center(%d) should be 0 (mass), 1 (geometry) or 2 (charge)", center);
        b2aEnsure();
        // one kernel launch + one D2H per (group, centre kind) per BATCH; frames are then handed out from the host copy
        CenCache&    c  = cen_[(grp << 2) | center];
        const size_t n3 = 3 * (size_t)nmol[grp];
        if (c.serial != serial_)
        {
            c.buf.resize(n3 * batch_k_);
            b2aCheck(b2a_centers_batch(ctx_, grp, center, c.buf.data()));
            c.serial = serial_;
        }
        std::memcpy(posc.data(), c.buf.data() + n3 * cur_, n3 * sizeof(double));   // matd(nmol,3) column-major
    }

    // ---- velocity / force loaders: no unwrapping, host side (ipp:201-339) ------------------------------------------------
    inline void GmxHandle::loadVelocity(boxd& vel, int grp)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        const int nm = nmol[grp], na = napm[grp];
        for (int i = 0, n = 0; i != nm; ++i)
            for (int j = 0; j != na; ++j, ++n)
                for (int k = 0; k != 3; ++k) vel(i, j)[k] = fr->v[index[grp][n]][k];
    }

    inline void GmxHandle::loadForce(boxd& force, int grp)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        const int nm = nmol[grp], na = napm[grp];
        for (int i = 0, n = 0; i != nm; ++i)
            for (int j = 0; j != na; ++j, ++n)
                for (int k = 0; k != 3; ++k) force(i, j)[k] = fr->f[index[grp][n]][k];
    }

    inline void GmxHandle::loadVelocityCenter(matd& velc, int grp, int com)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        const vecd   w = weights(grp, com);
        const double s = w.sum();
        for (int i = 0; i != nmol[grp]; ++i)
        {
            const int first = index[grp][i * napm[grp]];
            for (int k = 0; k != 3; ++k)
            {
                double acc = 0.0;
                for (int j = 0; j != napm[grp]; ++j) acc += double(fr->v[first + j][k]) * w[j];
                velc(i, k) = acc / s;
            }
        }
    }

    inline void GmxHandle::loadForceCenter(matd& forcec, int grp, int com)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        const vecd   w = weights(grp, com);
        const double s = w.sum();
        for (int i = 0; i != nmol[grp]; ++i)
        {
            const int first = index[grp][i * napm[grp]];
            for (int k = 0; k != 3; ++k)
            {
                double acc = 0.0;
                for (int j = 0; j != napm[grp]; ++j) acc += double(fr->f[first + j][k]) * w[j];
                forcec(i, k) = acc / s;
            }
        }
    }

    // ---- LJ table of the two template molecules (ipp:341-376) --------------------------------------------------------------
    inline void GmxHandle::loadLJParameter(int group1, int group2, matd& c6, matd& c12)
    {
        if ((group1 >= ngrps) || (group2 >= ngrps))
            gmx_fatal(FARGS, "grp(%d and %d) should less than ngrps(%d)!", group1, group2, ngrps);
        const int ntypes = top->atomtypes.nr;
        c6.resize(napm[group1], napm[group2]);
        c12.resize(napm[group1], napm[group2]);
        fmt::print("\n\033[31m### LJ Parameters(Group {} and {}):\033[0m\n", group1, group2);
        fmt::print("[  at1  at2]: {:12s} {:12s}\n", "c6", "c12");
        fmt::print("---------------------------------------\n");
        int shown = 0;
        for (int i = 0; i != napm[group1]; ++i)
            for (int j = 0; j != napm[group2]; ++j)
            {
                const int a = index[group1][i], b = index[group2][j];
                const int t = ntypes * top->atoms.atom[a].type + top->atoms.atom[b].type;
                c6(i, j)  = top->idef.iparams[t].lj.c6;
                c12(i, j) = top->idef.iparams[t].lj.c12;
                if (shown++ <= 12)
                    fmt::print("[{:5s}{:5s}]: {:12.4e} {:12.4e}\n", *top->atoms.atomname[a], *top->atoms.atomname[b], c6(i, j), c12(i, j));
            }
        if (napm[group1] * napm[group2] > 12) fmt::print("(.........)\n");
        fmt::print("\n");
    }

    // ---- output file with the provenance header (ipp:378-393): kept verbatim so outputs stay drop-in ------------------------
    inline FILE* GmxHandle::openWrite(std::string fnm, bool writeInfo)
    {
        FILE* fp = fopen(fnm.c_str(), "w");
        if (writeInfo)
        {
            fmt::print(fp, "# This file was created at {}\n", itp::localTime());
            fmt::print(fp, "# Working dir : {}\n", std::filesystem::current_path());
            fmt::print(fp, "# Command line:");
            for (int i = 0; i != argc; ++i) fmt::print(fp, " {}", argv[i]);
            fmt::print(fp, "\n");
        }
        return fp;
    }

    inline const char* GmxHandle::get_ftp2fn(int ftp) { return ftp2fn(ftp, (int)fnm.size(), fnm.data()); }
    inline const char* GmxHandle::get_ftp2fn_null(int ftp) { return ftp2fn_null(ftp, (int)fnm.size(), fnm.data()); }
    inline const char* GmxHandle::get_opt2fn(const char* opt) { return opt2fn(opt, (int)fnm.size(), fnm.data()); }

    inline double GmxHandle::periodicity(double dx, double box)   // ipp:410-417: strict inequalities
    {
        const double half = box / 2.0;
        while (dx > half) dx -= box;
        while (dx < -half) dx += box;
        return dx;
    }

    // ---- extensions: fused device accumulators ---------------------------------------------------------------------------------
    inline int GmxHandle::b2aDensity(const b2a_density_spec& s)
    {
        b2aEnsure();
        int acc = -1;
        b2aCheck(b2a_density_create(ctx_, &s, &acc));
        return acc;
    }

    inline int GmxHandle::b2aRdf(const b2a_rdf_spec& s, bool symmetric)
    {
        b2aEnsure();
        int acc = -1;
        b2aCheck(b2a_rdf_create(ctx_, &s, &acc));
        b2aCheck(b2a_rdf_tune(ctx_, acc, symmetric ? 1 : 0, 0));
        return acc;
    }

    inline int GmxHandle::b2aOrient(const b2a_orient_spec& s)
    {
        b2aEnsure();
        int acc = -1;
        b2aCheck(b2a_orient_create(ctx_, &s, &acc));
        return acc;
    }

    inline void GmxHandle::b2aAccumulate(int acc) { b2aCheck(b2a_accumulate(ctx_, acc)); }

    inline void GmxHandle::b2aRead(int acc, std::vector<uint64_t>& counts, uint64_t* stats)
    {
        size_t n = 0;
        b2aCheck(b2a_acc_size(ctx_, acc, &n));
        counts.assign(n, 0);
        b2aCheck(b2a_acc_read(ctx_, acc, counts.data(), stats));
    }
} // namespace itp
