// itp::GmxHandleFull -- B200 drop-in implementation (behaviour contract: reference
// include/itp/gmx_bits/gmx_handle_full.ipp, cited per method).
#pragma once

#include <algorithm>
#include <filesystem>

#include "gmx_handle_full.hpp"

namespace itp
{
    inline GmxHandleFull::GmxHandleFull(int argc, char** argv) :
        ngrps(1), flags(TRX_READ_X), nframe(0), argc(argc), argv(argv), status(nullptr), oenv(nullptr), ePBC(0)
    {
        grpname = nullptr; ngx = nullptr; top = nullptr; ir = nullptr; index = nullptr; fr = nullptr;
        dt = time = preTime = 0.0;
        Lbox[0] = Lbox[1] = Lbox[2] = 0.0;
        nthreads = omp_get_num_procs();   // full.ipp:9
    }

    inline void GmxHandleFull::init()   // full.ipp:12-64
    {
        fnm.push_back({ efTRX, "-f", nullptr, ffREAD });
        fnm.push_back({ efTPR, "-s", nullptr, ffREAD });
        fnm.push_back({ efNDX, "-n", nullptr, ffOPTRD });
        pa.push_back({ "-threads", FALSE, etINT, { &nthreads }, "number of threads" });

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

        napm.resize(ngrps); nmol.resize(ngrps); charge.resize(ngrps); mass.resize(ngrps);
        totMass.resize(ngrps); totCharge.resize(ngrps);
        for (int g = 0; g != ngrps; ++g)
        {
            get_natom_per_mol(g);
            get_mass_charge(g);
            fmt::print("\n\033[31m### Group {}: '{}': \033[0m\n", g, grpname[g]);
            fmt::print("# Nr. of atoms in the first molecule : {}\n", napm[g][0]);
            fmt::print("# Nr. of molecules in group: {}\n", nmol[g]);
            fmt::print("# Print infomation of the first molecule:\n");
            fmt::print("\033[33m# nr.   name     mass   charge\033[0m\n");
            fmt::print("------------------------------\n");
            for (int j = 0; j < napm[g][0]; j++)
                fmt::print("{:>5d}{:>7s}{:>9.3f}{:>9.3f}\n", j, *top->atoms.atomname[index[g][j]], mass[g][0][j], charge[g][0][j]);
            fmt::print("------------------------------\n");
            fmt::print("{:<12s}{:>9.3f}{:>9.3f}\n", "Total", totMass[g][0], totCharge[g][0]);
        }
    }

    // a new molecule starts wherever the residue index of consecutive group atoms changes (full.ipp:300-321)
    inline void GmxHandleFull::get_natom_per_mol(int grp)
    {
        napm[grp].assign(1, 0);
        int res = top->atoms.atom[index[grp][0]].resind;
        for (int n = 0; n < ngx[grp]; ++n)
        {
            const int r = top->atoms.atom[index[grp][n]].resind;
            if (r == res) napm[grp].back()++;
            else { res = r; napm[grp].push_back(1); }
        }
        nmol[grp] = (int)napm[grp].size();
    }

    inline void GmxHandleFull::get_mass_charge(int grp)   // full.ipp:323-344: totals are sequential sums
    {
        mass[grp].resize(nmol[grp]); charge[grp].resize(nmol[grp]);
        totMass[grp].assign(nmol[grp], 0.0); totCharge[grp].assign(nmol[grp], 0.0);
        int n = 0;
        for (int i = 0; i < nmol[grp]; ++i)
        {
            mass[grp][i].assign(napm[grp][i], 0.0); charge[grp][i].assign(napm[grp][i], 0.0);
            for (int j = 0; j < napm[grp][i]; ++j, ++n)
            {
                mass[grp][i][j]   = top->atoms.atom[index[grp][n]].m;
                charge[grp][i][j] = top->atoms.atom[index[grp][n]].q;
                totMass[grp][i] += mass[grp][i][j];
                totCharge[grp][i] += charge[grp][i][j];
            }
        }
    }

    inline void GmxHandleFull::b2aEnsure()
    {
        b2a_ctx* c = batch_.ctx(natoms_);
        if (groups_up_) return;
        for (int g = 0; g != ngrps; ++g)
        {
            std::vector<double> m, q;
            for (int i = 0; i < nmol[g]; ++i)
            {
                m.insert(m.end(), mass[g][i].begin(), mass[g][i].end());
                q.insert(q.end(), charge[g][i].begin(), charge[g][i].end());
            }
            b2a_detail::check(b2a_set_group_full(c, g, index[g], ngx[g], nmol[g], napm[g].data(), m.data(), q.data()));
        }
        groups_up_ = true;
    }

    inline bool GmxHandleFull::readFirstFrame()   // full.ipp:66-77
    {
        t_trxframe io{};
        bool       b = read_first_frame(oenv, &status, get_ftp2fn(efTRX), &io, flags);
        if (b)
        {
            natoms_ = io.natoms;
            b2aEnsure();
            b = batch_.start(oenv, status, io, fr);
        }
        b && (nframe = 1);
        Lbox[XX] = fr->box[XX][XX]; Lbox[YY] = fr->box[YY][YY]; Lbox[ZZ] = fr->box[ZZ][ZZ];
        time = preTime = fr->time;
        dt = 0;
        return b;
    }

    inline bool GmxHandleFull::readNextFrame()   // full.ipp:79-93
    {
        const bool b = batch_.next(fr);
        b && (++nframe);
        Lbox[XX] = fr->box[XX][XX]; Lbox[YY] = fr->box[YY][YY]; Lbox[ZZ] = fr->box[ZZ][ZZ];
        if (b) { dt = time - preTime; preTime = time; time = fr->time; }
        return b;
    }

    inline void GmxHandleFull::initPos(boxd& pos, int grp)   // full.ipp:95-103
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        pos.resize(nmol[grp], *std::max_element(napm[grp].begin(), napm[grp].end()));
    }

    inline void GmxHandleFull::initPosc(matd& posc, int grp)   // full.ipp:105-113
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        posc.resize(nmol[grp], 3);
        posc.fill(0);
    }

    inline void GmxHandleFull::loadPosition(boxd& pos, int grp)   // full.ipp:115-148
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        b2aEnsure();
        // pos must be boxd(nmol, max napm) as initPos makes it; slots beyond a molecule's atoms come back as zero
        b2a_detail::check(b2a_positions(batch_.ctx(), grp, batch_.current(), reinterpret_cast<double*>(pos.data())));
    }

    inline void GmxHandleFull::loadPositionCenter(matd& posc, int grp, int center)   // full.ipp:150-199
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        if (center < 0 || center > 2) gmx_fatal(FARGS, "center(%d) should be 0 (mass), 1 (geometry) or 2 (charge)", center);
        b2aEnsure();
        CenCache&    c  = cen_[(grp << 2) | center];
        const size_t n3 = 3 * (size_t)nmol[grp];
        if (c.serial != batch_.serial())
        {
            c.buf.resize(n3 * batch_.capacity());
            b2a_detail::check(b2a_centers_batch(batch_.ctx(), grp, center, c.buf.data()));
            c.serial = batch_.serial();
        }
        std::memcpy(posc.data(), c.buf.data() + n3 * batch_.current(), n3 * sizeof(double));
    }

    inline void GmxHandleFull::loadVelocityCenter(matd& velc, int grp, int com)   // full.ipp:201-250: no unwrap, on the device
    {
        if (grp >= ngrps) gmx_fatal(FARGS, "grp(%d) should less than ngrps(%d)!", grp, ngrps);
        if (com < 0 || com > 2) gmx_fatal(FARGS, "center(%d) should be 0 (mass), 1 (geometry) or 2 (charge)", com);
        b2aEnsure();
        CenCache&    c  = cen_[(1 << 8) | (grp << 2) | com];
        const size_t n3 = 3 * (size_t)nmol[grp];
        if (c.serial != batch_.serial())
        {
            c.buf.resize(n3 * batch_.capacity());
            b2a_detail::check(b2a_field_centers_batch(batch_.ctx(), B2A_FIELD_V, grp, com, c.buf.data()));
            c.serial = batch_.serial();
        }
        std::memcpy(velc.data(), c.buf.data() + n3 * batch_.current(), n3 * sizeof(double));
    }

    inline FILE* GmxHandleFull::openWrite(std::string fnm, bool writeInfo)   // full.ipp:252-267
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

    inline const char* GmxHandleFull::get_ftp2fn(int ftp) { return ftp2fn(ftp, (int)fnm.size(), fnm.data()); }
    inline const char* GmxHandleFull::get_ftp2fn_null(int ftp) { return ftp2fn_null(ftp, (int)fnm.size(), fnm.data()); }
    inline const char* GmxHandleFull::get_opt2fn(const char* opt) { return opt2fn(opt, (int)fnm.size(), fnm.data()); }

    inline double GmxHandleFull::periodicity(double dx, double box)
    {
        const double half = box / 2.0;
        while (dx > half) dx -= box;
        while (dx < -half) dx += box;
        return dx;
    }
} // namespace itp
