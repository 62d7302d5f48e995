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
        // top / ir / index / fr are never freed: the reference never frees them either and programs may hold pointers
        // into them until exit (SURVEY.md 8b "Ownership"); the device context and pinned batch die with batch_.
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
        b2a_ctx* c = batch_.ctx(natoms_);
        if (!groups_up_ && napm.size() == ngrps)
        {
            for (int g = 0; g != ngrps; ++g)
            {
                b2a_detail::check(b2a_set_group(c, g, index[g], ngx[g], napm[g], mass[g].data(), charge[g].data()));
                // hand over Eigen's own evaluation of atomCenter.sum() (ipp:195) so centres are bit-identical
                for (int com = 0; com < 3; ++com) b2a_detail::check(b2a_set_group_wsum(c, g, com, weights(g, com).sum()));
            }
            groups_up_ = true;
        }
    }

    // ---- frame stepping (reference ipp:68-95) ----------------------------------------------------------------------
    inline bool GmxHandle::readFirstFrame()
    {
        // a second call restarts the trajectory, like read_first_frame does (two-pass programs rely on it)
        t_trxframe io{};
        bool       b = read_first_frame(oenv, &status, get_ftp2fn(efTRX), &io, flags);
        if (b)
        {
            natoms_ = io.natoms;
            b2aEnsure();
            b = batch_.start(oenv, status, io, fr);
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
        const bool b = batch_.next(fr);
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
        nframe += batch_.skip_to_batch_end();   // the frames of this batch that were consumed as a block
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
        b2a_detail::check(b2a_positions(batch_.ctx(), grp, batch_.current(), reinterpret_cast<double*>(pos.data())));
    }

    inline void GmxHandle::loadPositionCenter(matd& posc, int grp, int center)   // ipp:150-199
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        if (center < 0 || center > 2) gmx_fatal(FARGS, "center(%d) should be 0 (mass), 1 (geometry) or 2 (charge)", center);
        b2aEnsure();
        // one kernel launch + one D2H per (group, centre kind) per BATCH; frames are then handed out from the host copy
        CenCache&    c  = cen_[(grp << 2) | center];
        const size_t n3 = 3 * (size_t)nmol[grp];
        if (c.serial != batch_.serial())
        {
            c.buf.resize(n3 * batch_.capacity());
            b2a_detail::check(b2a_centers_batch(batch_.ctx(), grp, center, c.buf.data()));
            c.serial = batch_.serial();
        }
        std::memcpy(posc.data(), c.buf.data() + n3 * batch_.current(), n3 * sizeof(double));   // matd(nmol,3) column-major
    }

    // ---- velocity / force loaders (ipp:201-339): same weighted mean / gather without unwrapping, on the device ---------------
    inline void GmxHandle::loadField(boxd& out, int grp, int field)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        b2aEnsure();
        b2a_detail::check(b2a_field_values(batch_.ctx(), field, grp, batch_.current(), reinterpret_cast<double*>(out.data())));
    }

    inline void GmxHandle::loadFieldCenter(matd& out, int grp, int com, int field)
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        if (com < 0 || com > 2) gmx_fatal(FARGS, "center(%d) should be 0 (mass), 1 (geometry) or 2 (charge)", com);
        b2aEnsure();
        CenCache&    c  = cen_[(field << 8) | (grp << 2) | com];
        const size_t n3 = 3 * (size_t)nmol[grp];
        if (c.serial != batch_.serial())
        {
            c.buf.resize(n3 * batch_.capacity());
            b2a_detail::check(b2a_field_centers_batch(batch_.ctx(), field, grp, com, c.buf.data()));
            c.serial = batch_.serial();
        }
        std::memcpy(out.data(), c.buf.data() + n3 * batch_.current(), n3 * sizeof(double));
    }

    inline void GmxHandle::loadVelocity(boxd& vel, int grp) { loadField(vel, grp, B2A_FIELD_V); }
    inline void GmxHandle::loadForce(boxd& force, int grp) { loadField(force, grp, B2A_FIELD_F); }
    inline void GmxHandle::loadVelocityCenter(matd& velc, int grp, int com) { loadFieldCenter(velc, grp, com, B2A_FIELD_V); }
    inline void GmxHandle::loadForceCenter(matd& forcec, int grp, int com) { loadFieldCenter(forcec, grp, com, B2A_FIELD_F); }

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
        b2a_detail::check(b2a_density_create(batch_.ctx(), &s, &acc));
        return acc;
    }

    inline int GmxHandle::b2aRegion(const b2a_region_spec& s)
    {
        b2aEnsure();
        int acc = -1;
        b2a_detail::check(b2a_region_create(batch_.ctx(), &s, &acc));
        return acc;
    }

    inline void GmxHandle::b2aRegionFrames(int acc, std::vector<uint64_t>& out)
    {
        size_t n = 0;
        b2a_detail::check(b2a_acc_size(batch_.ctx(), acc, &n));
        out.assign(n * (size_t)batch_.frames(), 0);
        b2a_detail::check(b2a_region_read_frames(batch_.ctx(), acc, reinterpret_cast<unsigned long long*>(out.data())));
    }

    inline int GmxHandle::b2aRdf(const b2a_rdf_spec& s, bool symmetric)
    {
        b2aEnsure();
        int acc = -1;
        b2a_detail::check(b2a_rdf_create(batch_.ctx(), &s, &acc));
        b2a_detail::check(b2a_rdf_tune(batch_.ctx(), acc, symmetric ? 1 : 0, 0));
        return acc;
    }

    inline int GmxHandle::b2aOrient(const b2a_orient_spec& s)
    {
        b2aEnsure();
        int acc = -1;
        b2a_detail::check(b2a_orient_create(batch_.ctx(), &s, &acc));
        return acc;
    }

    inline int GmxHandle::b2aCoord(const b2a_coord_spec& s)
    {
        b2aEnsure();
        int acc = -1;
        b2a_detail::check(b2a_coord_create(batch_.ctx(), &s, &acc));
        return acc;
    }

    inline void GmxHandle::b2aCoordFold(int acc, int max_count, std::vector<double>& totPair)
    {
        const int nf = batch_.frames();
        std::vector<uint32_t> table((size_t)nf * max_count), numA(nf);
        totPair.resize(max_count, 0.0);
        b2a_detail::check(b2a_coord_read_frames(batch_.ctx(), acc, table.data(), numA.data()));
        b2a_detail::check(b2a_coord_fold(table.data(), numA.data(), nf, max_count, totPair.data()));
    }

    inline size_t GmxHandle::b2aSeries(int acc, std::vector<uint64_t>& out, size_t* words, std::vector<float>* times)
    {
        size_t n = 0, w = 0;
        b2a_detail::check(b2a_series_size(batch_.ctx(), acc, &n, &w));
        out.assign(n * w, 0);
        std::vector<long long> idx(n);
        b2a_detail::check(b2a_series_read(batch_.ctx(), acc, idx.data(), out.data()));
        if (words) *words = w;
        if (times)
        {
            const std::vector<float>& all = batch_.all_times();
            times->resize(n);
            for (size_t f = 0; f < n; ++f) (*times)[f] = (idx[f] >= 0 && (size_t)idx[f] < all.size()) ? all[idx[f]] : 0.f;
        }
        return n;
    }

    inline void GmxHandle::b2aCoordFoldSeries(int acc, int max_count, std::vector<double>& totPair)
    {
        std::vector<uint64_t> rec;
        size_t                w = 0;
        const size_t          n = b2aSeries(acc, rec, &w);
        std::vector<uint32_t> table(n * (size_t)max_count), numA(n);
        for (size_t f = 0; f < n; ++f)
        {
            for (int k = 0; k < max_count; ++k) table[f * max_count + k] = (uint32_t)rec[f * w + k];
            numA[f] = (uint32_t)rec[f * w + max_count];
        }
        totPair.resize(max_count, 0.0);
        b2a_detail::check(b2a_coord_fold(table.data(), numA.data(), (int)n, max_count, totPair.data()));
    }

    inline void GmxHandle::b2aAccumulate(int acc) { b2a_detail::check(b2a_accumulate(batch_.ctx(), acc)); }

    inline void GmxHandle::b2aRead(int acc, std::vector<uint64_t>& counts, uint64_t* stats)
    {
        size_t n = 0;
        b2a_detail::check(b2a_acc_size(batch_.ctx(), acc, &n));
        counts.assign(n, 0);
        b2a_detail::check(b2a_acc_read(batch_.ctx(), acc, counts.data(), stats));
    }
} // namespace itp
