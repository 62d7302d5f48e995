// itp::GmxHandle -- B200 drop-in.  Same public surface as the reference class (reference
// include/itp/gmx_bits/gmx_handle.hpp:7-192): constructor, init, readFirstFrame/readNextFrame, initPos/initPosc,
// load{Position,Velocity,Force}[Center], loadLJParameter, openWrite, file-name helpers, periodicity, and every
// public data member programs touch directly (hd.fr->x, hd.Lbox, hd.nmol, hd.nframe, ...).
//
// What changed underneath: frames are decoded K at a time into a pinned host batch and shipped to
// the GPU with one asynchronous copy; loadPosition / loadPositionCenter are served by libb2a's sm_100a kernels for
// the whole batch and handed out frame by frame; `fr` always points at the CURRENT frame inside the batch, so
// direct reads of hd.fr->x / time / box between readNextFrame() calls keep working.  Methods and members whose
// names start with `b2a` are extensions that ported programs use to reach the fused device accumulators.
#pragma once

#include "b2a_batcher.hpp"

#include <map>

namespace itp
{
    class GmxHandle
    {
    public:
        GmxHandle(int argc, char** argv);
        GmxHandle(const GmxHandle&) = delete;
        GmxHandle(GmxHandle&&) = delete;
        GmxHandle& operator=(const GmxHandle&) = delete;
        GmxHandle& operator=(GmxHandle&&) = delete;

        void init(bool bFullMolecule = true);
        bool readFirstFrame();
        bool readNextFrame();

        boxd initPos(int grp);
        matd initPosc(int grp);

        void loadPosition(boxd& pos, int grp);
        void loadPositionCenter(matd& posc, int grp, int com = 0);
        void loadVelocity(boxd& vel, int grp);
        void loadVelocityCenter(matd& velc, int grp, int com = 0);
        void loadForce(boxd& force, int grp);
        void loadForceCenter(matd& forcec, int grp, int com = 0);
        void loadLJParameter(int group1, int group2, matd& c6, matd& c12);

        FILE* openWrite(std::string fnm, bool writeInfo = true);
        const char* get_ftp2fn(int ftp);
        const char* get_ftp2fn_null(int ftp);
        const char* get_opt2fn(const char* opt);
        static double periodicity(double dx, double box);

        // ---- extensions (not in the reference): fused device accumulators over whole batches ----------------
        b2a_ctx* b2aContext() { return batch_.ctx(); }
        int  b2aBatchFrames() const { return batch_.frames(); }          // frames in the resident batch
        float b2aBatchTime(int k) const { return batch_.time_of(k); }    // fr->time of frame k of the resident batch
        bool b2aReadNextBatch();                                   // skip to the first frame of the next batch
        int  b2aDensity(const b2a_density_spec& s);
        int  b2aRegion(const b2a_region_spec& s);                  // filters / cylinder / wrap + 0..2 binned dimensions
        // per-frame counters of the batch just accumulated (spec.per_frame): out[frame][counter]
        void b2aRegionFrames(int acc, std::vector<uint64_t>& out);
        int  b2aRdf(const b2a_rdf_spec& s, bool symmetric = true);
        int  b2aOrient(const b2a_orient_spec& s);
        int  b2aCoord(const b2a_coord_spec& s);                    // coordination numbers (mofs/calc_pair_dist*.cpp)
        // after b2aAccumulate(acc) of a coordination accumulator: totPair[k] += tmpPair[k] / numA for every frame of the batch, in
        // frame order (calc_pair_dist_bulk.cpp:89-92); totPair must hold spec.max_count doubles
        void b2aCoordFold(int acc, int max_count, std::vector<double>& totPair);
        void b2aAccumulate(int acc);                               // whole resident batch, once per batch
        void b2aRead(int acc, std::vector<uint64_t>& counts, uint64_t* stats = nullptr);
        // Several GPUs (environment B2A_DEVICES=0,1,... or "all"): the frame loop above is unchanged -- batches are dealt to the
        // devices in turn; b2aRead merges the devices (one NCCL all-reduce).  Programs with per-frame output must not read inside
        // the loop (that would serialise the devices): they fetch the whole series afterwards, ordered by frame.
        int  b2aDevices() { return b2a_num_devices(batch_.ctx()); }
        // per-frame records of a region (spec.per_frame) or coordination accumulator over every frame accumulated so far, in
        // trajectory order: out[frame][word]; returns the number of frames; times (optional) receives fr->time of each
        size_t b2aSeries(int acc, std::vector<uint64_t>& out, size_t* words = nullptr, std::vector<float>* times = nullptr);
        // totPair[k] += tmpPair[k] / numA over the whole series, in frame order (calc_pair_dist_bulk.cpp:89-92)
        void b2aCoordFoldSeries(int acc, int max_count, std::vector<double>& totPair);

    private:
        int  get_natom_per_mol(int grp);
        vecd get_mass(int grp);
        vecd get_charge(int grp);
        void b2aEnsure();
        vecd weights(int grp, int com);
        void loadField(boxd& out, int grp, int field);
        void loadFieldCenter(matd& out, int grp, int com, int field);

    public:
        std::vector<const char*> desc = { "do some analysis for trajectort." };

        double               Lbox[3];      // box lengths (diagonal of fr->box)
        char**               grpname;      // group names
        int*                     ngx;      // group sizes
        t_topology*              top;      // topology
        t_inputrec*               ir;      // input record
        double                    dt;      // ps (one frame late, like the reference)
        double                  time;      // current time
        double               preTime;      // previous time
        int**                  index;      // atom indices of every group
        t_trxframe*               fr;      // CURRENT frame (x/v/f point into the pinned batch)
        int                    flags;      // TRX_READ_X by default
        int                    ngrps;      // number of groups, default 1
        int                   nframe;      // frames read so far
        std::vector<t_pargs>      pa;      // user-defined options
        std::vector<t_filenm>    fnm;      // files

        veci                    napm;      // atoms per molecule in each group
        veci                    nmol;      // molecules in each group
        Vec<vecd>               mass;      // per-atom-in-molecule masses
        Vec<vecd>             charge;      // per-atom-in-molecule charges
        int                 nthreads;      // declared by the reference, unused

    protected:
        int                     argc;
        char**                  argv;
        t_trxstatus*          status;
        gmx_output_env_t*       oenv;
        int                     ePBC;

    private:
        b2a_detail::FrameBatcher batch_;          // pinned K-frame batches (env B2A_BATCH), device context
        int                      natoms_ = 0;
        bool                     groups_up_ = false;
        // host cache of batch centres: key (grp << 2 | com) -> [n][3][nmol] doubles, valid for one batch serial
        struct CenCache { std::vector<double> buf; long serial = -1; };
        std::map<int, CenCache>  cen_;
    };
} // namespace itp
