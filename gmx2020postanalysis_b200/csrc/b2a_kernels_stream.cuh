// TMA-staged streaming kernels for the HBM-bound part of the path (K1 centres, K2 density, K4 orientation).
//
// When an index group is one contiguous run of atoms (the normal case: a whole species), the atoms of the
// blockDim.x molecules a block owns are ONE contiguous byte range of every frame.  One elected thread moves that
// range global -> shared with a 1-D bulk tensor copy (cp.async.bulk, SASS UBLKCP) that signals an mbarrier; the
// block keeps kStages frames in flight and computes out of shared memory.  This replaces the strided per-thread
// 4-byte loads of the direct kernels (36 sectors touched per warp instruction, one molecule of latency hiding per
// thread) by perfectly coalesced 16-byte-granular traffic with kStages x tile bytes outstanding per block
// (profiles/r01_hbm.md).  The per-molecule arithmetic is the same FP64 code as in the direct kernels.
#pragma once

#include "b2a_kernels_fast.cuh"

namespace b2a
{

constexpr int kStages   = 4;
// the two-molecules-per-thread orientation variant stages 512-molecule tiles: three of them keep three blocks on an SM
__host__ __device__ constexpr int stream_stages(int mode, int var) { return ((mode == 2 && var == 1) || (mode == 0 && var == 2)) ? 3 : kStages; }
constexpr int kFastQueue = 1024;  // undecided (molecule, frame) entries per block, drained at the end of the block; overflow is evaluated inline

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// the producer's wait for a free stage: the ring is normally full, so this thread would spin for most of the kernel and
// take issue slots from the consumer warps of its scheduler (ncu: 20 try_wait round trips per tile); the suspend-time hint
// lets the hardware park it until the phase completes or the limit (ns) expires
__device__ __forceinline__ void mbar_wait_parked(unsigned long long* bar, unsigned parity, unsigned park_ns)
{
    if (park_ns == 0) { mbar_wait(bar, parity); return; }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(park_ns)
        : "memory");
}

struct StreamArgs
{
    const float*     x;
    size_t           frame_stride;   // floats per frame
    int              nframes, atom0, nmol, napm;
    const double*    w;
    DivConst         wsum;
    const FrameGeom* geom;
    int              item_frames;    // frames per work item
    int              mblocks;        // molecule blocks (blockDim.x molecules each)
    int              nitems;         // mblocks * ceil(nframes / item_frames)
    unsigned*        work;           // {next work item to hand out, blocks that have finished}; both zero at launch, re-zeroed by the last block
    int              tile_floats;    // floats reserved per stage (blockDim.x * napm * 3 + 8)
    unsigned         park_ns;        // suspend-time hint of the producer's wait for a free stage (0: plain try_wait loop)
    unsigned         cwait_ns;       // the same for the consumers' wait for a full stage
    // MODE 0: centres
    double*          cen;
    int4*            fix;
    // MODE 1: density, MODE 2: orientation
    DensityDev       den;
    OrientDev        ori;
    const double*    cosedge;
    const float*     fedge;
    const double*    stab;
    unsigned long long* hist;
    unsigned long long* stats;
    int              smem_bins;
    // FP32-bracketed fast path (b2a_kernels_fast.cuh)
    FastParams           fp;
    const float*         wf;     // weights rounded to float
    DensityFast          df;
    OrientFast           of;
    const unsigned char* lut;    // MODE 2: kCosLut first-candidate angle bins
    float                ori_sign;   // MODE 2, VAR 1: sign that turns orient_fast3x2's vector into the reference's vecOut
    int                  rot_shift;  // MODE 2, VAR 1: log2(32 / gcd(3 napm, 32)); 5 = odd stride, no rotation
    int                  ori_reuse;  // MODE 2, VAR 1, three-site molecules: the vector atoms are (reference atom, the other two)
    const float2*        fmh;        // MODE 2: per angle bin {midpoint, halfwidth - table roundings} of its cosine interval
    // MODE 2 with a fused centre output (b2a_acc_also_centers): the consumer that evaluates a molecule also writes its exact FP64
    // centre of a second kind (e.g. the dipole vector, com 3) into `cen` -- one pass over the coordinates instead of K1 + K4
    const double*        w2;         // weights of the fused centre kind (napm doubles), or nullptr
    DivConst             wsum2;
    unsigned long long*  dbg;        // optional per-block time stamps (B2A_STREAM_DEBUG; tools/bench_k4.py --timeline): 8 words per block
};

__device__ __forceinline__ void cp_async4(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct StageDesc { int m, f, poff, pad; };   // molecule block, frame, float offset of the block's first atom inside the stage; m < 0: end

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MODE 0 = K1 centres (+fixed point), 1 = K2 density, 2 = K4 orientation.  NAPM_T: atoms per molecule at compile time
// (loops unroll completely), 0 = run-time value.
//
// Persistent, warp-specialised blocks.  The grid is sized to the machine (resident blocks x SMs); every block pulls work
// items (molecule block, item_frames consecutive frames) from a global counter.  One PRODUCER warp streams the
// (molecule block, frame) tiles plus that frame's geometry through a kStages-deep ring with cp.async.bulk (UBLKCP),
// signalling a `full` mbarrier per stage; the CONSUMER warps (blockDim.x/32 - 1, one molecule per thread) wait on `full`,
// compute out of shared memory and release the stage by arriving on its `empty` mbarrier, warp by warp.  There is no
// block-wide barrier inside the stream: a slow warp never stalls the others, the ring never drains between items, and
// tables / histogram zero+flush happen once per block for the whole launch.
// VAR (MODE 2): 1 = two molecules per consumer thread in packed FP32 (a tile holds 2 x consumer-thread-count molecules):
// orient_fastNx2() for any molecule, orient_fast3x2() for three-site molecules whose vector atoms are (reference atom, the
// other two).
// FUSE (MODE 2): the instantiation that also writes a second kind of centre (b2a_acc_also_centers); a separate instantiation because
// the FP64 centre code costs the plain orientation kernel 4 % through register pressure even when it is switched off at run time.
// SPEC (MODE 2, VAR 1): 1 = calc_orient_mof's default axes and method folded in at compile time (orient_tail_x2<1>).
template <int MODE, int NAPM_T, int VAR, bool FUSE = false, int SPEC = 0>
__global__ void __launch_bounds__(288, (MODE == 2 && VAR == 1) ? 3 : 1) k_stream(const __grid_constant__ StreamArgs a)
{
    constexpr int NS = stream_stages(MODE, VAR);   // depth of the TMA ring
    extern __shared__ __align__(16) unsigned char sraw[];
    // layout: [tiles: NS x tile_floats floats] [frame geometry per stage] [full bars] [empty bars] [stage descriptors]
    //         [weights] [edges] [float weights] [histogram] [undecided queue] [cos look-up table]
    float*              tiles = reinterpret_cast<float*>(sraw);
    FrameGeom*          sgeom = reinterpret_cast<FrameGeom*>(tiles + (size_t)NS * a.tile_floats);     // NS entries (TMA)
    unsigned long long* full  = reinterpret_cast<unsigned long long*>(sgeom + NS);
    unsigned long long* empty = full + NS;
    StageDesc*          sdesc = reinterpret_cast<StageDesc*>(empty + NS);
    double*             sw    = reinterpret_cast<double*>(sdesc + NS);
    const int           napm  = NAPM_T ? NAPM_T : a.napm;
    double*             sedge = sw + napm;                                                     // MODE 2: nAngle + 1
    double*             stab  = sedge + (MODE == 2 ? a.ori.nAngle + 1 : 0);                    // MODE 2: nbin1 + 2
    float*              fedge = reinterpret_cast<float*>(stab + (MODE == 2 ? a.ori.nbin1 + 2 : 0));   // MODE 2: nAngle + 1
    float2*             fmh   = reinterpret_cast<float2*>(fedge + (MODE == 2 ? ((a.ori.nAngle + 4) & ~1) : 0));   // MODE 2: nAngle + 2
    float*              swf   = reinterpret_cast<float*>(fmh + (MODE == 2 ? a.ori.nAngle + 2 : 0));              // napm rounded up to even
    unsigned*           shist = reinterpret_cast<unsigned*>(swf + ((napm + 1) & ~1));
    uint2*              fqueue = reinterpret_cast<uint2*>(shist + (MODE != 0 ? ((a.smem_bins + 1) & ~1) : 0));
    unsigned char*      slut   = reinterpret_cast<unsigned char*>(fqueue + (MODE != 0 ? kFastQueue : 0));
    double*             sw2    = reinterpret_cast<double*>(slut + (MODE == 2 ? kCosLut : 0));   // MODE 2, fused centre output: napm weights
    __shared__ unsigned fq_n;

    if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 0] = gtime();
    const int T = blockDim.x;          // all threads (prologue / epilogue loops)
    const int M = blockDim.x - 32;     // consumer threads
    // molecules per tile.  MODE 2 / VAR 1: two per consumer thread.  MODE 0 / VAR 2: THREE LANES per molecule (one per
    // dimension), ten molecules per warp -- the exact FP64 centre of a large molecule is one long dependent chain per
    // dimension, and its atoms take 12 * napm bytes of every stage: one thread per molecule leaves an SM with a handful of
    // warps crawling along those chains (napm = 25: 0.19 of HBM peak, tools/bench_hbm_napm.py)
    const int MT = (MODE == 2 && VAR == 1) ? 2 * M : ((MODE == 0 && VAR == 2) ? 10 * (M >> 5) : M);   // (host: launch_stream_t)
    // Prologue.  The consumer threads first ISSUE the copies of the block's tables (weights, edge tables, cos look-up table) as
    // asynchronous global -> shared copies (cp.async, LDGSTS): nothing waits for them yet.  Then the barriers are armed and the
    // producer starts the ring; the table copies, issued ahead of the bulk traffic, land while the first stages are in flight,
    // and the consumers collect them (cp.async.wait_all + a consumer-only named barrier) just before the first stage arrives.
    if ((int)threadIdx.x < M)
    {
        for (int j = threadIdx.x; j < napm; j += M) { cp_async8(&sw[j], &a.w[j]); if (MODE != 0) cp_async4(&swf[j], &a.wf[j]); }
        if (MODE == 2 && FUSE)
            for (int j = threadIdx.x; j < napm; j += M) cp_async8(&sw2[j], &a.w2[j]);
        if (MODE == 2 && a.fp.enabled)
            for (int j = threadIdx.x; j < kCosLut / 4; j += M) cp_async4(reinterpret_cast<unsigned*>(slut) + j, reinterpret_cast<const unsigned*>(a.lut) + j);
        if (MODE == 2)
        {
            for (int j = threadIdx.x; j <= a.ori.nAngle; j += M) cp_async8(&sedge[j], &a.cosedge[j]);
            for (int j = threadIdx.x; j < a.ori.nAngle + 3; j += M) cp_async4(&fedge[j], &a.fedge[j]);   // two -3 sentinels past the last edge
            for (int j = threadIdx.x; j < a.ori.nAngle + 2; j += M) cp_async8(&fmh[j], &a.fmh[j]);
            for (int j = threadIdx.x; j < a.ori.nbin1 + 2; j += M) cp_async8(&stab[j], &a.stab[j]);
        }
        if (MODE != 0)
            for (int b = threadIdx.x; b < a.smem_bins; b += M) shist[b] = 0;
    }
    if (threadIdx.x == 0)
    {
        fq_n = 0;
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (unsigned)(M >> 5)); }
        mbar_fence_init();
    }
    __syncthreads();
    if ((int)threadIdx.x < M)
    {
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"r"(M) : "memory");
    }
    const OrientTabs tb{ sedge, fedge, stab };
    if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 1] = gtime();

    unsigned   ndrop = 0, nsel = 0, nexact = 0, nbad = 0;
    const bool use_smem = MODE == 1 ? a.smem_bins >= a.den.nbin
                                    : (MODE == 2 ? a.smem_bins >= a.ori.nbin1 * a.ori.nAngle + a.ori.nbin1 : false);

    if ((int)threadIdx.x >= M)
    {
        // ---- producer warp: one elected lane walks the items and keeps the ring full ------------------------------------
        if (threadIdx.x == M)
        {
            // the next item is claimed one item ahead, so the round trip of the global atomic is off the refill path
            unsigned it = atomicAdd(a.work, 1u), next = atomicAdd(a.work, 1u);
            int      f = 0, fend = 0, m = -1;
            for (unsigned n = 0;; ++n)
            {
                const int s = (int)(n % NS);
                if (n >= (unsigned)NS) mbar_wait_parked(&empty[s], ((n / NS) - 1u) & 1u, a.park_ns);
                if (m >= 0 && f >= fend) { m = -1; it = next; next = atomicAdd(a.work, 1u); }
                if (m < 0 && it < (unsigned)a.nitems)
                {
                    m    = (int)(it % (unsigned)a.mblocks);
                    f    = (int)(it / (unsigned)a.mblocks) * a.item_frames;
                    fend = min(a.nframes, f + a.item_frames);
                }
                if (m < 0)
                {
                    sdesc[s].m = -1;       // end of stream: consumers stop at the first such stage
                    mbar_arrive(&full[s]);
                    break;
                }
                const int            mvalid = min(MT, a.nmol - m * MT);
                const unsigned       nbytes = (unsigned)mvalid * napm * 12u;
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + (size_t)f * a.frame_stride + ((size_t)a.atom0 + (size_t)m * MT * napm) * 3);
                const unsigned       sh  = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
                const unsigned       nb  = (sh + nbytes + 15u) & ~15u;   // the byte range widened to 16-byte boundaries
                sdesc[s] = StageDesc{ m, f, (int)(s * a.tile_floats + (sh >> 2)), 0 };
                mbar_expect_tx(&full[s], nb + (unsigned)sizeof(FrameGeom));
                tma_load_1d(tiles + (size_t)s * a.tile_floats, src - sh, nb, &full[s]);
                tma_load_1d(&sgeom[s], a.geom + f, (unsigned)sizeof(FrameGeom), &full[s]);
                ++f;
            }
        }
    }
    else
    {
        // ---- consumer warps ------------------------------------------------------------------------------------------------
        const unsigned toff = threadIdx.x * (unsigned)napm * 3u;
        // orient_fast3x2: the two off-reference weights and the sign of the computed vector, kept in registers
        const float w3a = (MODE == 2 && VAR == 1 && NAPM_T == 3) ? a.wf[1] : 0.f, w3b = (MODE == 2 && VAR == 1 && NAPM_T == 3) ? a.wf[2] : 0.f;
        const float w3s = (MODE == 2 && VAR == 1) ? a.ori_sign : 0.f;
        // bank-conflict rotation of the atom walk (orient_fastNx2): lane block index, 0 for odd strides
        const int lane_rot = (MODE == 2 && VAR == 1 && a.rot_shift < 5) ? (int)((threadIdx.x & 31u) >> a.rot_shift) : 0;
        for (unsigned n = 0;; ++n)
        {
            const int s = (int)(n % NS);
            mbar_wait_parked(&full[s], (n / NS) & 1u, a.cwait_ns);
            const StageDesc sd = sdesc[s];
            if (a.dbg && threadIdx.x == 0 && n == 0) a.dbg[blockIdx.x * 8 + 2] = gtime();   // first stage has landed
            if (sd.m < 0) break;   // the stream is in order: nothing follows the first end marker
            const int f = sd.f;
            const int i = sd.m * MT + (int)threadIdx.x;
            // (VAR 1 masks molecules past the end of the group instead of branching around them)
            const bool vA = i < a.nmol;
            if constexpr (MODE == 0 && VAR == 2)
            {
                const int  lane = (int)threadIdx.x & 31, ml = lane / 3, k = lane - 3 * ml;
                const int  il = (int)(threadIdx.x >> 5) * 10 + ml;            // molecule inside the tile
                const int  im = sd.m * MT + il;
                const bool ok = lane < 30 && im < a.nmol;
                const FrameGeom& g = sgeom[s];
                double           ck = 0.0;
                if (ok)
                {
                    ck = center_component_pipelined<true>(tiles + sd.poff + (unsigned)il * (unsigned)napm * 3u, k, napm, sw, a.wsum, g.L[k], g.half[k], g.hsafe[k]);
                    a.cen[(size_t)f * 3 * a.nmol + (size_t)k * a.nmol + im] = ck;
                }
                if (a.fix)
                {   // the fixed-point triple of a molecule sits in three neighbouring lanes
                    const int fk = ok ? to_fixed(ck, g.inv_unit[k]) : 0;
                    const int fy = __shfl_down_sync(0xffffffffu, fk, 1), fz = __shfl_down_sync(0xffffffffu, fk, 2);
                    if (ok && k == 0) a.fix[(size_t)f * a.nmol + im] = make_int4(fk, fy, fz, im);
                }
            }
            else if ((MODE == 2 && VAR == 1) || vA)
            {
                const float*     p = tiles + sd.poff + (vA ? toff : 0u);   // !vA: the tile's first molecule, evaluated and ignored
                const FrameGeom& g = sgeom[s];
                if (MODE == 0)
                {
                    double c0, c1, c2;
                    mol_centers3<true>(p, napm, sw, a.wsum, g, c0, c1, c2);
                    double* o = a.cen + (size_t)f * 3 * a.nmol;
                    o[i]                      = c0;
                    o[(size_t)a.nmol + i]     = c1;
                    o[2 * (size_t)a.nmol + i] = c2;
                    if (a.fix)
                        a.fix[(size_t)f * a.nmol + i] = make_int4(to_fixed(c0, g.inv_unit[0]), to_fixed(c1, g.inv_unit[1]), to_fixed(c2, g.inv_unit[2]), i);
                }
                else if (MODE == 1)
                {
                    int me;
                    if (a.fp.enabled == 1)
                    {
                        me = density_fast<true>(p, napm, swf, a.fp, g.ff, a.den, a.df, nsel, ndrop);
                        if (me == -2)
                        {
                            const unsigned slot = atomicAdd(&fq_n, 1u);
                            if (slot < (unsigned)kFastQueue) fqueue[slot] = make_uint2((unsigned)i, (unsigned)f);
                            else { me = density_molecule<true>(p, napm, sw, a.wsum, g, a.den, nsel, ndrop); nexact++; }
                        }
                    }
                    else
                    {
                        me = density_molecule<true>(p, napm, sw, a.wsum, g, a.den, nsel, ndrop);
                        if (a.fp.enabled == 2)   // validation: a proven FP32 answer must equal the FP64 one
                        {
                            unsigned fs = 0, fd = 0;
                            const int mf = density_fast<true>(p, napm, swf, a.fp, g.ff, a.den, a.df, fs, fd);
                            if (mf == -2) nexact++;
                            else if (mf != me) nbad++;
                        }
                    }
                    if (me >= 0)
                    {
                        if (use_smem) atomicAdd(&shist[me], 1u);
                        else atomicAdd(&a.hist[me], 1ull);
                    }
                }
                else
                {
                    if constexpr (FUSE)
                    {   // fused centre output: the same FP64 code as K1 (mol_centers3), on the coordinates already in shared memory
                        auto put = [&](const float* pm, int im) {
                            double c0, c1, c2;
                            mol_centers3<true>(pm, napm, sw2, a.wsum2, g, c0, c1, c2);
                            double* o = a.cen + (size_t)f * 3 * a.nmol;
                            o[im]                      = c0;
                            o[(size_t)a.nmol + im]     = c1;
                            o[2 * (size_t)a.nmol + im] = c2;
                        };
                        if (vA) put(p, i);
                        if (VAR == 1 && i + M < a.nmol) put(p + (unsigned)M * (unsigned)napm * 3u, i + M);
                    }
                    auto count = [&](int cz, int ca) {
                        if (cz >= 0)
                        {
                            if (use_smem) atomicAdd(&shist[cz], 1u);
                            else atomicAdd(&a.hist[cz], 1ull);
                        }
                        if (ca >= 0)
                        {
                            if (use_smem) atomicAdd(&shist[ca], 1u);
                            else atomicAdd(&a.hist[ca], 1ull);
                        }
                    };
                    // undecided by the FP32 bracket: queue for the literal FP64 code (inline when the queue is full)
                    // (a full queue evaluates inline, from global memory like the drain at the end of the kernel)
                    auto undecided = [&](int im) {
                        const unsigned slot = atomicAdd(&fq_n, 1u);
                        if (slot < (unsigned)kFastQueue) { fqueue[slot] = make_uint2((unsigned)im, (unsigned)f); return; }
                        int          cz, ca;
                        const float* pg = a.x + (size_t)f * a.frame_stride + ((size_t)a.atom0 + (size_t)im * napm) * 3;
                        orient_molecule<false>(pg, napm, sw, a.wsum, a.geom[f], a.ori, tb, cz, ca, nsel, ndrop);
                        nexact++;
                        count(cz, ca);
                    };
                    if constexpr (VAR == 1)
                    {
                        const int    iB = i + M;                 // second molecule of this thread
                        const bool   vB = iB < a.nmol;
                        const float* pB = vB ? p + (unsigned)M * (unsigned)napm * 3u : p;   // past the end of the group: a copy that is ignored
                        int res[2], meZ[2], lo[2];
                        if (a.fp.enabled)
                        {
                            if (NAPM_T == 3 && a.ori_reuse)   // block-uniform: vectors = the unwrapped offsets themselves
                            {
                                Mol3x2 mol;
                                load3x2<true>(mol, p, pB, g.fk);
                                orient_fast3x2<SPEC>(mol, vA, vB, w3a, w3b, w3s, a.fp, a.ori, a.of, fedge, fmh, slut, res, meZ, lo);
                            }
                            else
                                orient_fastNx2<true, SPEC>(p, pB, vA, vB, napm, lane_rot, swf, a.fp, g.fk, a.ori, a.of, fedge, fmh, slut, res, meZ, lo);
                        }
                        // accounting of a proven molecule, statement for statement as at the end of orient_molecule(); the host
                        // selects this variant only when the histogram fits in shared memory
                        auto cells = [&](int l, int& cz, int& ca, unsigned& ns, unsigned& nd) {
                            cz = ca = -1;
                            if (res[l] != 2) return;
                            ns++;
                            if (meZ[l] >= a.ori.nbin1) { nd++; return; }
                            cz = a.ori.nbin1 * a.ori.nAngle + meZ[l];
                            if (lo[l] >= a.ori.nAngle) { nd++; return; }
                            ca = meZ[l] + lo[l] * a.ori.nbin1;
                        };
                        if (a.fp.enabled == 1)
                        {
#pragma unroll
                            for (int l = 0; l < 2; ++l)
                            {
                                if (res[l] == 0) undecided(l ? iB : i);   // rare
                                // accounting of a proven molecule as in cells(), without branches: two predicated shared atomics
                                const bool sel = res[l] == 2;
                                const bool zok = sel & (meZ[l] < a.ori.nbin1);
                                const bool aok = zok & (lo[l] < a.ori.nAngle);
                                nsel += sel ? 1u : 0u;
                                ndrop += ((sel & !zok) | (zok & !aok)) ? 1u : 0u;
                                if (zok) atomicAdd(&shist[a.ori.nbin1 * a.ori.nAngle + meZ[l]], 1u);
                                if (aok) atomicAdd(&shist[meZ[l] + lo[l] * a.ori.nbin1], 1u);
                            }
                        }
                        else
                        {
                            for (int l = 0; l < 2; ++l)
                            {
                                if (!(l ? vB : vA)) continue;
                                int ez, ea;
                                orient_molecule<true>(l ? pB : p, napm, sw, a.wsum, g, a.ori, tb, ez, ea, nsel, ndrop);
                                if (a.fp.enabled == 2)   // validation: a proven FP32 answer must equal the FP64 one
                                {
                                    unsigned fs = 0, fd = 0;
                                    int      fz, fa;
                                    cells(l, fz, fa, fs, fd);
                                    if (!res[l]) nexact++;
                                    else if (fz != ez || fa != ea) nbad++;
                                }
                                count(ez, ea);
                            }
                        }
                    }
                    else
                    {
                        int cz = -1, ca = -1;
                        if (a.fp.enabled == 1)
                        {
                            if (orient_fast<true>(p, napm, swf, a.fp, g.ff, a.ori, a.of, fedge, slut, cz, ca, nsel, ndrop)) count(cz, ca);
                            else undecided(i);
                        }
                        else
                        {
                            orient_molecule<true>(p, napm, sw, a.wsum, g, a.ori, tb, cz, ca, nsel, ndrop);
                            if (a.fp.enabled == 2)
                            {
                                unsigned fs = 0, fd = 0;
                                int      fz = -1, fa = -1;
                                if (!orient_fast<true>(p, napm, swf, a.fp, g.ff, a.ori, a.of, fedge, slut, fz, fa, fs, fd)) nexact++;
                                else if (fz != cz || fa != ca) nbad++;
                            }
                            count(cz, ca);
                        }
                    }
                }
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);   // this warp is done with stage s
        }
    }
    if (a.dbg && threadIdx.x == 0) { a.dbg[blockIdx.x * 8 + 3] = gtime(); }   // warp 0 has seen the end of the stream
    if (MODE != 0)
    {
        __syncthreads();
        if (a.dbg && threadIdx.x == 0) { a.dbg[blockIdx.x * 8 + 4] = gtime(); a.dbg[blockIdx.x * 8 + 7] = fq_n; }
        // undecided molecules: literal FP64 evaluation, read back from global memory (their stage has been refilled)
        const unsigned nq = min(fq_n, (unsigned)kFastQueue);
        for (unsigned e = threadIdx.x; e < nq; e += T)
        {
            const uint2      ent = fqueue[e];
            const int        f   = (int)ent.y;
            const float*     pg  = a.x + (size_t)f * a.frame_stride + ((size_t)a.atom0 + (size_t)ent.x * napm) * 3;
            const FrameGeom& g   = a.geom[f];
            nexact++;
            if (MODE == 1)
            {
                const int me = density_molecule<false>(pg, napm, sw, a.wsum, g, a.den, nsel, ndrop);
                if (me >= 0)
                {
                    if (use_smem) atomicAdd(&shist[me], 1u);
                    else atomicAdd(&a.hist[me], 1ull);
                }
            }
            else
            {
                int cz, ca;
                orient_molecule<false>(pg, napm, sw, a.wsum, g, a.ori, tb, cz, ca, nsel, ndrop);
                if (cz >= 0)
                {
                    if (use_smem) atomicAdd(&shist[cz], 1u);
                    else atomicAdd(&a.hist[cz], 1ull);
                }
                if (ca >= 0)
                {
                    if (use_smem) atomicAdd(&shist[ca], 1u);
                    else atomicAdd(&a.hist[ca], 1ull);
                }
            }
        }
        __syncthreads();
        if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 5] = gtime();   // queue drained
        const int nb = MODE == 1 ? a.den.nbin : a.ori.nbin1 * a.ori.nAngle + a.ori.nbin1;
        if (use_smem)
            for (int b = threadIdx.x; b < nb; b += T)
                if (shist[b]) atomicAdd(&a.hist[b], (unsigned long long)shist[b]);
        for (int o = 16; o > 0; o >>= 1)
        {
            ndrop += __shfl_xor_sync(0xffffffffu, ndrop, o);
            nsel += __shfl_xor_sync(0xffffffffu, nsel, o);
            nexact += __shfl_xor_sync(0xffffffffu, nexact, o);
            nbad += __shfl_xor_sync(0xffffffffu, nbad, o);
        }
        if ((threadIdx.x & 31) == 0)
        {
            if (ndrop) atomicAdd(&a.stats[1], (unsigned long long)ndrop);
            if (nsel) atomicAdd(&a.stats[4], (unsigned long long)nsel);
            if (nexact) atomicAdd(&a.stats[2], (unsigned long long)nexact);
            if (nbad) atomicAdd(&a.stats[7], (unsigned long long)nbad);
        }
        if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 6] = gtime();
    }
    // the block that finishes last re-arms the work counter for the next launch (every block has long stopped claiming items:
    // a block gets here only after its producer saw the end of the item list)
    if (threadIdx.x == 0)
    {
        __threadfence();
        if (atomicAdd(a.work + 1, 1u) == gridDim.x - 1u) { a.work[0] = 0u; a.work[1] = 0u; __threadfence(); }
    }
}

} // namespace b2a
