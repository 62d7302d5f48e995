// itp::GmxHandleFull -- B200 drop-in for the heterogeneous-molecule handle (reference
// include/itp/gmx_bits/gmx_handle_full.hpp:7-158): same public members and methods; molecules are split by residue
// (full.ipp:300-321), each keeps its own atom count, masses, charges and totals; position loaders run in libb2a.so.
#pragma once

#include "b2a_batcher.hpp"

#include <map>

namespace itp
{
    class GmxHandleFull
    {
    public:
        template <typename T>
        using stdMatrix = std::vector<std::vector<T>>;

        GmxHandleFull(int argc, char** argv);
        GmxHandleFull(const GmxHandleFull&) = delete;
        GmxHandleFull(GmxHandleFull&&) = delete;
        GmxHandleFull& operator=(const GmxHandleFull&) = delete;
        GmxHandleFull& operator=(GmxHandleFull&&) = delete;

        void init();
        bool readFirstFrame();
        bool readNextFrame();
        void initPos(boxd& pos, int grp);
        void initPosc(matd& posc, int grp);
        void loadPosition(boxd& pos, int grp);
        void loadPositionCenter(matd& posc, int grp, int com = 0);
        void loadVelocityCenter(matd& velc, int grp, int com = 0);
        FILE* openWrite(std::string fnm, bool writeInfo = true);
        const char* get_ftp2fn(int ftp);
        const char* get_ftp2fn_null(int ftp);
        const char* get_opt2fn(const char* opt);
        static double periodicity(double dx, double box);

        b2a_ctx* b2aContext() { return batch_.ctx(); }   // extension: access to the fused accumulators of include/b2a.h

    private:
        void get_natom_per_mol(int grp);
        void get_mass_charge(int grp);
        void b2aEnsure();

    public:
        // ---- what a program sets before init() --------------------------------------------------------------------------
        std::vector<const char*> desc = { "do some analysis for trajectort." };
        std::vector<t_pargs>  pa;                 // user-defined options
        std::vector<t_filenm> fnm;                // user-defined files
        int ngrps;                                // number of index groups to ask for (default 1)
        int flags;                                // TRX_READ_* (default TRX_READ_X)
        int nthreads;                             // "-threads" option of the reference; unused by the device path

        // ---- filled by init(): topology, groups, molecules ----------------------------------------------------------------
        t_topology* top;
        t_inputrec* ir;
        char** grpname;
        int*   ngx;
        int**  index;
        std::vector<int> nmol;                    // molecules per group
        stdMatrix<int>   napm;                    // [grp][mol] atoms per molecule
        stdMatrix<std::vector<double>> mass;      // [grp][mol][atom]
        stdMatrix<std::vector<double>> charge;    // [grp][mol][atom]
        stdMatrix<double> totMass;                // [grp][mol]
        stdMatrix<double> totCharge;              // [grp][mol]

        // ---- per frame ----------------------------------------------------------------------------------------------------------
        t_trxframe* fr;                           // CURRENT frame; x / v / f point into the pinned batch
        double Lbox[3];                           // box diagonal
        double time, preTime, dt;                 // dt lags one frame, like the reference
        int    nframe;                            // frames read so far

    protected:
        int    argc;
        char** argv;
        t_trxstatus*      status;
        gmx_output_env_t* oenv;
        int    ePBC;

    private:
        b2a_detail::FrameBatcher batch_;
        int                      natoms_ = 0;
        bool                     groups_up_ = false;
        struct CenCache { std::vector<double> buf; long serial = -1; };
        std::map<int, CenCache>  cen_;
    };
} // namespace itp
