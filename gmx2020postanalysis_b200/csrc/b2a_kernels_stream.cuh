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

#include "b2a_kernels_orient.cuh"

namespace b2a
{

constexpr int kStages = 4;

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

struct StreamArgs
{
    const float*     x;
    size_t           frame_stride;   // floats per frame
    int              nframes, atom0, nmol, napm;
    const double*    w;
    DivConst         wsum;
    const FrameGeom* geom;
    int              frames_per_block;
    int              tile_floats;    // floats reserved per stage (blockDim.x * napm * 3 + 8)
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
};

// MODE 0 = K1 centres (+fixed point), 1 = K2 density, 2 = K4 orientation
template <int MODE>
__global__ void __launch_bounds__(256) k_stream(const __grid_constant__ StreamArgs a)
{
    extern __shared__ __align__(16) unsigned char sraw[];
    // layout: [tiles: kStages x tile_floats floats, 16-B aligned] [bars] [frame geometry] [weights] [edges] [histogram]
    float*              tiles = reinterpret_cast<float*>(sraw);
    unsigned long long* bars  = reinterpret_cast<unsigned long long*>(tiles + (size_t)kStages * a.tile_floats);
    FrameGeom*          sgeom = reinterpret_cast<FrameGeom*>(bars + kStages);          // frames_per_block entries
    double*             sw    = reinterpret_cast<double*>(sgeom + a.frames_per_block);
    double*             sedge = sw + a.napm;                                                   // MODE 2: nAngle + 1
    double*             stab  = sedge + (MODE == 2 ? a.ori.nAngle + 1 : 0);                    // MODE 2: nbin1 + 2
    float*              fedge = reinterpret_cast<float*>(stab + (MODE == 2 ? a.ori.nbin1 + 2 : 0));   // MODE 2: nAngle + 1
    unsigned*           shist = reinterpret_cast<unsigned*>(fedge + (MODE == 2 ? ((a.ori.nAngle + 2) & ~1) : 0));

    const int M      = blockDim.x;
    const int m0     = blockIdx.x * M;
    const int mvalid = min(M, a.nmol - m0);
    const int f0     = blockIdx.y * a.frames_per_block;
    const int f1     = min(a.nframes, f0 + a.frames_per_block);
    const int nf     = f1 - f0;
    if (mvalid <= 0 || nf <= 0) return;

    for (int j = threadIdx.x; j < a.napm; j += M) sw[j] = a.w[j];
    {   // frame geometry once per block: it is indexed by a run-time dimension, which must not end up in local memory
        const double* src = reinterpret_cast<const double*>(a.geom + f0);
        double*       dst = reinterpret_cast<double*>(sgeom);
        for (int j = threadIdx.x; j < nf * (int)(sizeof(FrameGeom) / sizeof(double)); j += M) dst[j] = src[j];
    }
    if (MODE == 2)
    {
        for (int j = threadIdx.x; j <= a.ori.nAngle; j += M) { sedge[j] = a.cosedge[j]; fedge[j] = a.fedge[j]; }
        for (int j = threadIdx.x; j < a.ori.nbin1 + 2; j += M) stab[j] = a.stab[j];
    }
    const OrientTabs tb{ sedge, fedge, stab };
    if (MODE != 0)
        for (int b = threadIdx.x; b < a.smem_bins; b += M) shist[b] = 0;
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // the byte range of this block's molecules in frame f, widened to 16-byte boundaries
    const size_t   first_float = ((size_t)a.atom0 + (size_t)m0 * a.napm) * 3;
    const unsigned nbytes      = (unsigned)mvalid * a.napm * 12u;
    auto issue = [&](int k) {   // k-th frame of this block -> stage k % kStages   (thread 0 only)
        const int            s   = k % kStages;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.x + (size_t)(f0 + k) * a.frame_stride + first_float);
        const unsigned       sh  = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        const unsigned       nb  = (sh + nbytes + 15u) & ~15u;
        mbar_expect_tx(&bars[s], nb);
        tma_load_1d(tiles + (size_t)s * a.tile_floats, src - sh, nb, &bars[s]);
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < min(kStages, nf); ++k) issue(k);

    const bool active = (int)threadIdx.x < mvalid;
    const int  i      = m0 + threadIdx.x;
    unsigned   ndrop = 0, nsel = 0;
    const bool use_smem = MODE == 1 ? a.smem_bins >= a.den.nbin
                                    : (MODE == 2 ? a.smem_bins >= a.ori.nbin1 * a.ori.nAngle + a.ori.nbin1 : false);

    for (int k = 0; k < nf; ++k)
    {
        const int s = k % kStages;
        mbar_wait(&bars[s], (unsigned)((k / kStages) & 1));
        const int f = f0 + k;
        if (active)
        {
            const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(a.x + (size_t)f * a.frame_stride + first_float) & 15u);
            const float*   p  = tiles + (size_t)s * a.tile_floats + (sh >> 2) + (size_t)threadIdx.x * a.napm * 3;
            const FrameGeom& g = sgeom[k];
            if (MODE == 0)
            {
                double c0, c1, c2;
                mol_centers3<true>(p, a.napm, sw, a.wsum, g, c0, c1, c2);
                double* o = a.cen + (size_t)f * 3 * a.nmol;
                o[i]                      = c0;
                o[(size_t)a.nmol + i]     = c1;
                o[2 * (size_t)a.nmol + i] = c2;
                if (a.fix)
                    a.fix[(size_t)f * a.nmol + i] = make_int4(to_fixed(c0, g.inv_unit[0]), to_fixed(c1, g.inv_unit[1]), to_fixed(c2, g.inv_unit[2]), i);
            }
            else if (MODE == 1)
            {
                const int me = density_molecule<true>(p, a.napm, sw, a.wsum, g, a.den, nsel, ndrop);
                if (me >= 0)
                {
                    if (use_smem) atomicAdd(&shist[me], 1u);
                    else atomicAdd(&a.hist[me], 1ull);
                }
            }
            else
            {
                int cz, ca;
                orient_molecule<true>(p, a.napm, sw, a.wsum, g, a.ori, tb, cz, ca, nsel, ndrop);
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
        __syncthreads();   // every thread is done with stage s: it may be refilled
        if (threadIdx.x == 0 && k + kStages < nf) issue(k + kStages);
    }
    if (MODE != 0)
    {
        const int nb = MODE == 1 ? a.den.nbin : a.ori.nbin1 * a.ori.nAngle + a.ori.nbin1;
        if (use_smem)
            for (int b = threadIdx.x; b < nb; b += M)
                if (shist[b]) atomicAdd(&a.hist[b], (unsigned long long)shist[b]);
        for (int o = 16; o > 0; o >>= 1)
        {
            ndrop += __shfl_xor_sync(0xffffffffu, ndrop, o);
            nsel += __shfl_xor_sync(0xffffffffu, nsel, o);
        }
        if ((threadIdx.x & 31) == 0)
        {
            if (ndrop) atomicAdd(&a.stats[1], (unsigned long long)ndrop);
            if (nsel) atomicAdd(&a.stats[4], (unsigned long long)nsel);
        }
    }
}

} // namespace b2a
