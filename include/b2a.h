/*
 * b2a.h -- C ABI of libb2a.so: the B200-native per-frame analysis path behind itp::GmxHandle.
 *
 * The reference has no FFI layer: its boundary is the C++ class itp::GmxHandle
 * (reference include/itp/gmx_bits/gmx_handle.hpp:7-192) plus the per-frame loops that analysis
 * programs write by hand on the arrays it returns.  This header is the boundary inserted INSIDE that
 * class: include/itp/gmx forwards the loaders to it, and ported programs /
 * bench.py call the fused accumulators.  Everything is extern "C", plain pointers and sizes, status
 * codes instead of gmx_fatal (the C++ wrapper turns them back into gmx_fatal with the reference's
 * messages).  One context = one host control thread driving one GPU (b2a_create) or several (b2a_create_multi:
 * K-frame batches are dealt to the devices in turn, histograms are merged by one NCCL all-reduce in b2a_finalize);
 * calls on a context are not thread-safe, distinct contexts are independent.  All device work is queued on the
 * context's stream(s); calls that return data to host memory synchronise, everything else -- b2a_submit_batch and
 * b2a_accumulate included -- is asynchronous (b2a_submit_batch waits only for the PREVIOUS upload to leave the
 * caller's buffer; b2a_accumulate only if the host is four batches ahead of the device).
 *
 * Reference interface replaced by each entry point (paths relative to /root/reference):
 *   b2a_set_group      <- GmxHandle::init: get_natom_per_mol / get_mass / get_charge results
 *                         (include/itp/gmx_bits/gmx_handle.ipp:36-65, 419-468)
 *   b2a_submit_batch   <- readFirstFrame / readNextFrame: fr->x, fr->box for K frames (ipp:68-95)
 *   b2a_positions      <- GmxHandle::loadPosition          (ipp:115-148)
 *   b2a_centers        <- GmxHandle::loadPositionCenter    (ipp:150-199), com 0/1/2; com 3 = dipole
 *                         numerator with the neutral-molecule guard of
 *                         src/general/9g_dist_bin_1803.cpp:215-222
 *   b2a_submit_field / b2a_field_* <- fr->v, fr->f; GmxHandle::loadVelocity[Center] / loadForce[Center] (ipp:201-339)
 *   b2a_density_*      <- README.md:129-154 template loop; src/general/calc_density_region.cpp:59-71;
 *                         src/general/density_slit_jhl.cpp:59-77
 *   b2a_rdf_*          <- src/general/calc_rdf_region.cpp:52-64 (setup), 73-95 (pair loop),
 *                         100-115 (normalisation); GmxHandle::periodicity (ipp:410-417)
 *   b2a_orient_*       <- src/mofs/calc_orient_mof.cpp:14-63, 120-200
 *   b2a_coord_*        <- src/mofs/calc_pair_dist_bulk.cpp:14-23, 66-100; src/mofs/calc_pair_dist.cpp:80-106
 *   b2a_region_*       <- src/mofs/calc_density_distribution_in_pore.cpp:47-60; calc_areaDensityDistributionPore.cpp:56-68;
 *                         calc_numdens_pore.cpp:12-27, 72-80; calc_radius_distribution_in_pore.cpp:45-58;
 *                         calc_radius_distribution_in_pore_time.cpp:81-107; calc_density_time.cpp:41-60 (per-frame counters)
 *   b2a_set_group_full <- GmxHandleFull::init (include/itp/gmx_bits/gmx_handle_full.ipp:300-344); its loaders share
 *                         b2a_centers / b2a_positions / b2a_field_centers (full.ipp:115-250)
 *   b2a_set_group_wsum <- callers that divide by their own weight sum (e.g. the dipole programs' charge sums)
 *   b2a_acc_read       <- the post-loop reads of the programs' histograms (`density /= hd.nframe` and friends)
 *   b2a_create_multi / b2a_finalize / b2a_series_* <- the single frame loop of a program (src/general/calc_rdf_region.cpp:68-97) sharded
 *                         over the GPUs of one box by contiguous frame blocks, and the per-frame outputs
 *                         (src/mofs/calc_density_time.cpp:41-58; calc_pair_dist_bulk.cpp:89-92) gathered in frame order
 * No reference counterpart (infrastructure of the boundary): b2a_comm_* (multi-process communicator), b2a_member / num_devices /
 * current_device / set_frame_base, b2a_create / destroy / sync / stream / version / device_count / last_error, b2a_pinned_alloc / free (the frame batcher's host buffers), b2a_accumulate / acc_reset / acc_size /
 * acc_device_ptr (multi-GPU all-reduce), b2a_rdf_tune / rdf_cells / acc_fast / invalidate / timing_* (experiments, validation).
 */
#ifndef B2A_H
#define B2A_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2a_ctx b2a_ctx;

enum {
    B2A_OK              = 0,
    B2A_ERR_ARG         = 1, /* bad argument (message says which; mirrors the reference's gmx_fatal checks) */
    B2A_ERR_CUDA        = 2, /* CUDA runtime failure, or no usable device: there is NO CPU fallback */
    B2A_ERR_STATE       = 3, /* call out of order (e.g. loader before any batch was submitted) */
    B2A_ERR_NOMEM       = 4,
    B2A_ERR_UNSUPPORTED = 5
};

/* centre kinds: the reference's `com` argument (ipp:161-174) plus the guarded dipole */
enum { B2A_COM_MASS = 0, B2A_COM_GEOMETRY = 1, B2A_COM_CHARGE = 2, B2A_COM_DIPOLE = 3 };

/* ---- context ------------------------------------------------------------------------------------ */
const char* b2a_last_error(void);      /* thread-local text of the last failure */
const char* b2a_version(void);
int         b2a_device_count(void);    /* 0 if no CUDA device is visible */

/* natoms = atoms per frame; max_frames = largest batch that will be submitted (device buffers are
 * sized once: 12 B x natoms x max_frames for coordinates).                                        */
int  b2a_create(b2a_ctx** ctx, int device, int natoms, int max_frames);
void b2a_destroy(b2a_ctx* ctx);
int  b2a_sync(b2a_ctx* ctx);
void* b2a_stream(b2a_ctx* ctx);        /* the cudaStream_t all work is queued on (but see b2a_join) */
/* In a context that holds several RDF accumulators (a sweep over group pairs) each of them queues its kernels on an internal
 * stream of its own, forked from the context stream, so that the short launches of small group pairs run beside the long ones.
 * b2a_join makes the context stream wait for all of that (asynchronously: nothing blocks on the host).  Every call that
 * synchronises or starts a new batch implies it (b2a_sync, b2a_acc_read, b2a_acc_reset, b2a_submit_batch, b2a_invalidate,
 * b2a_finalize); only callers that order their OWN device work after b2a_stream() need to call it.                        */
int  b2a_join(b2a_ctx* ctx);

/* ---- several GPUs behind ONE context -------------------------------------------------------------
 * ndev member contexts, one per entry of devices[] (NULL = devices 0..ndev-1).  The returned LEADER accepts every call
 * of this header: groups, accumulator creation / tuning / reset go to all devices; b2a_submit_batch deals the batch (a
 * contiguous block of frames) to the NEXT device in turn and b2a_accumulate / the loaders act on the device holding the
 * current batch; b2a_acc_read returns the totals over all devices (it calls b2a_finalize when they are stale).  A program
 * written as `do { b2a_submit_batch; b2a_accumulate; } while (frames); b2a_acc_read;` therefore runs on N GPUs unchanged
 * (the drop-in handles create a leader when the environment names several devices: B2A_DEVICES=0,1,2,3 or "all").
 * The communicator (ncclCommInitAll) is built here, not inside the frame loop.  devices[] may repeat a device (used by
 * the one-GPU tests: the merge is then a local sum, NCCL cannot place two ranks on one device).                        */
int  b2a_create_multi(b2a_ctx** ctx, int ndev, const int* devices, int natoms, int max_frames);
int  b2a_num_devices(b2a_ctx* ctx);              /* 1 for a b2a_create context */
b2a_ctx* b2a_member(b2a_ctx* ctx, int i);        /* device i's own context (per-device streams, timing); owned by ctx */
int  b2a_current_device(b2a_ctx* ctx);           /* member index that holds the current batch (-1 before the first) */

/* How many host staging buffers the caller cycles through (default 2).  On a multi-device context b2a_submit_batch returns once
 * the upload issued n - 1 submits ago has left its buffer; with n = devices + 1 the host never waits for a copy and the uploads
 * of different devices (separate PCIe links) overlap.  A plain context keeps the two-buffer contract whatever n is.            */
int  b2a_set_host_buffers(b2a_ctx* ctx, int n);

/* Merge: ONE all-reduce (64-bit unsigned integer sum, ncclAllReduce over NVLink) per accumulator over its counters and
 * stats block, across every device of the context and -- after b2a_comm_init -- of every process.  The private counts of
 * each device stay untouched (the sum lands in a second buffer): accumulation may go on and b2a_finalize may be called
 * again.  Until the next b2a_accumulate / b2a_acc_reset, b2a_acc_read returns the merged totals (on every process).
 * Collective: with a multi-process communicator every process must call it.  Also gathers the per-frame series.        */
int  b2a_finalize(b2a_ctx* ctx);

/* Multi-process (one process per GPU or per group of GPUs, e.g. under torchrun / mpirun): rank 0 obtains an id
 * (128 bytes), the launcher's own channel broadcasts it, every process joins.  Device i of process p is rank
 * p * b2a_num_devices + i.                                                                                              */
int  b2a_comm_unique_id(void* id128);
int  b2a_comm_init(b2a_ctx* ctx, int nprocs, int proc_rank, const void* id128);

/* Global index of the first frame of the NEXT submitted batch (default: frames submitted so far on this context).
 * Processes that each read their own block of a trajectory set it so that per-frame series merge in trajectory order. */
int  b2a_set_frame_base(b2a_ctx* ctx, long long first_frame);

/* Per-frame series of an accumulator (region accumulators with spec.per_frame; coordination numbers): one record per
 * accumulated frame since creation / reset, ORDERED BY GLOBAL FRAME INDEX over all devices (and, after b2a_finalize,
 * all processes).  words_per_frame: region = its counters; coordination = max_count table entries (tmpPair) followed by
 * numA.  frame_index may be NULL.  This is what lets time-series programs (mofs/calc_density_time.cpp:41-58) keep every
 * device busy: nothing is read back inside the frame loop.                                                             */
int  b2a_series_size(b2a_ctx* ctx, int acc, size_t* nframes, size_t* words_per_frame);
int  b2a_series_read(b2a_ctx* ctx, int acc, long long* frame_index, uint64_t* out);

/* pinned host staging for async H2D (optional; pageable memory also works, synchronously) */
void* b2a_pinned_alloc(size_t nbytes);
void  b2a_pinned_free(void* p);

/* ---- index groups (A2) ------------------------------------------------------------------------- */
/* index[ngx] atom ids; napm atoms per molecule; molecules are the consecutive napm-runs of index, and
 * atoms of a molecule are taken as index[i*napm] + j exactly as the reference does (ipp:187).
 * mass/charge: napm doubles (template = first molecule, ipp:450-468).  Sum of weights is evaluated in
 * the order Eigen's vectorised Array::sum() uses (what ipp:195 calls); b2a_set_group_wsum overrides it
 * with the caller's own value (the C++ wrapper passes atomCenter.sum() itself).                     */
int b2a_set_group(b2a_ctx* ctx, int grp, const int* index, int ngx, int napm, const double* mass, const double* charge);
int b2a_set_group_wsum(b2a_ctx* ctx, int grp, int com, double wsum);
/* Heterogeneous molecule sizes: replaces the results of GmxHandleFull::init (reference
 * include/itp/gmx_bits/gmx_handle_full.ipp:300-344): napm[nmol] atoms per molecule (consecutive runs of index),
 * mass/charge per index position [ngx]; per-molecule totals are summed sequentially as totMass/totCharge are.
 * com = 1 uses weight 1 and divisor napm[i] (the intent of full.ipp:166-171, which indexes an empty vector).
 * Loaders, density and RDF accumulators accept such a group; b2a_positions fills an nmol x max(napm) boxd buffer
 * (slots j >= napm[i] are returned as 0).                                                                   */
int b2a_set_group_full(b2a_ctx* ctx, int grp, const int* index, int ngx, int nmol, const int* napm, const double* mass,
                       const double* charge);
int b2a_group_nmol(b2a_ctx* ctx, int grp);

/* ---- frames (A1) ------------------------------------------------------------------------------- */
/* x: [nframes][natoms][3] float (rvec layout), box9: [nframes][9] float row-major matrix; only the
 * diagonal is used, as in the reference (ipp:72-74).  Replaces the current batch; cached centres of
 * the previous batch are invalidated.  The upload is asynchronous and double-buffered on the device: it runs on its
 * own stream into the buffer the kernels of the PREVIOUS batch are not reading, so a caller that submits batch k+1
 * right after queueing the accumulators of batch k (and only then reads results of batch k) overlaps the H2D copy
 * with the kernels.  Everything launched after the call waits for the upload.  `x` (pinned memory for a truly
 * asynchronous copy) must stay unmodified until the NEXT b2a_submit_batch has returned or the context has been
 * synchronised; callers alternate between two host buffers (include/itp/gmx_bits/b2a_batcher.hpp does).         */
int b2a_submit_batch(b2a_ctx* ctx, int nframes, const float* x, const float* box9);
int b2a_batch_frames(b2a_ctx* ctx);

/* ---- compatibility loaders (A3, A4): results in the reference's Eigen memory layouts ------------ */
/* out: nmol x 3 doubles, column-major (itp::matd): element (i,k) at i + k*nmol */
int b2a_centers(b2a_ctx* ctx, int grp, int com, int frame, double* out);
/* all frames of the batch at once: out[frame][k][i] */
int b2a_centers_batch(b2a_ctx* ctx, int grp, int com, double* out);
/* The same centres left ON THE DEVICE: queues the centre kernel for the whole current batch (asynchronous, cached per batch) and
 * returns the device address of out[frame][k][i] (valid until the next b2a_submit_batch on this device; consumers order their
 * work after b2a_stream).  For device-side consumers of e.g. the dipole vectors (com 3) that do not want a D2H copy per frame.   */
int b2a_centers_device(b2a_ctx* ctx, int grp, int com, const double** dptr);
/* out: nmol x napm x 3 doubles in itp::boxd memory order: element (i,j)[k] at (i + j*nmol)*3 + k */
int b2a_positions(b2a_ctx* ctx, int grp, int frame, double* out);

/* ---- velocities / forces (fr->v, fr->f of TRR trajectories): the loaders that do not unwrap ----- */
enum { B2A_FIELD_V = 1, B2A_FIELD_F = 2 };
/* data: [nframes][natoms][3] float, the v or f arrays of the frames of the CURRENT batch (call after
 * b2a_submit_batch; nframes must match).                                                          */
int b2a_submit_field(b2a_ctx* ctx, int field, int nframes, const float* data);
/* GmxHandle::loadVelocityCenter / loadForceCenter (ipp:228-269, 298-339), com 0/1/2; layouts as
 * b2a_centers / b2a_centers_batch                                                                 */
int b2a_field_centers(b2a_ctx* ctx, int field, int grp, int com, int frame, double* out);
int b2a_field_centers_batch(b2a_ctx* ctx, int field, int grp, int com, double* out);
/* GmxHandle::loadVelocity / loadForce (ipp:201-226, 271-296); layout as b2a_positions             */
int b2a_field_values(b2a_ctx* ctx, int field, int grp, int frame, double* out);

/* ---- fused accumulators: process every frame of the current batch on the device ---------------- */
typedef struct b2a_density_spec
{
    int    grp, com;
    int    dim;              /* 0,1,2 */
    float  low, up;          /* `real` in the reference programs */
    double dbin;             /* bin width as the program computes it (README.md:127) */
    int    nbin;
    int    upper_inclusive;  /* 0: low <= p < up (README.md:145); 1: low <= p <= up (density_slit_jhl.cpp:62) */
    int    dim2;             /* second filter dimension, or -1 (calc_density_region.cpp:64) */
    float  low2, up2;
} b2a_density_spec;

/* General region counter / histogram over molecule centres: the remaining members of the density family.
 *   mofs/calc_density_distribution_in_pore.cpp:47-60  z profile inside a cylinder: one half-open filter, `cyl`, one binned dimension
 *   mofs/calc_areaDensityDistributionPore.cpp:56-68   2-D map: three half-open filters, two binned dimensions (dens(me1, me2))
 *   mofs/calc_numdens_pore.cpp:12-27, 72-80           molecules inside a box after wrapping into [0, L]: `wrap`, three filters, no bins
 * All arithmetic is the programs': doubles for the centre, the program's `real` bounds widened in the comparison, truncation of
 * (p - low) / dbin.  Counters: me0 + me1 * nb[0] (Eigen column-major), or one counter when nbdim == 0.                            */
typedef struct b2a_region_spec
{
    int    grp, com;
    int    nfilter;            /* 0..3 region tests, all must pass: low <= p < up, or p <= up where fincl */
    int    fdim[3];
    float  flow[3], fup[3];
    int    fincl[3];
    int    wrap;               /* calc_numdens_pore.cpp:15-20: p is first cast to float and moved into [0, L] by repeated -+L (float
                                  variable, double box length), and the bounds are compared in float                               */
    int    cyl;                /* calc_density_distribution_in_pore.cpp:52-53: pow(x - cx, 2) + pow(y - cy, 2) <= R2 (x = dim 0, y = dim 1) */
    float  cx, cy;
    double R2;                 /* as the program computes it: double(Rmof) * Rmof                                                    */
    int    nbdim;              /* 0, 1 or 2 binned dimensions */
    int    bdim[2];            /* 0, 1, 2: a centre component; 3: the distance from the cylinder axis, sqrt(pow(x - cx, 2) + pow(y - cy, 2))
                                  (mofs/calc_radius_distribution_in_pore.cpp:47-49)                                                  */
    float  blow[2];
    double dbin[2];            /* bin widths as the program holds them (a `real` widened, or a double)                               */
    int    nb[2];
    int    per_frame;          /* also keep the counters of every frame of the last accumulated batch (time series programs:
                                  mofs/calc_density_time.cpp:38-55), readable with b2a_region_read_frames                             */
} b2a_region_spec;

typedef struct b2a_rdf_spec
{
    int   grp0, grp1;        /* i runs over grp0, j over grp1 (calc_rdf_region.cpp:73,79) */
    int   com0, com1;        /* centre kind for each (the program uses mass for both, :70-71) */
    int   nbin;              /* slabs along dim (:23) */
    float low, up;           /* strict: low < p < up (:75) */
    int   dim, dim2;
    float low2, up2;
    float Rm;                /* cut-off, already defaulted by the caller (:52-57) */
    int   Rbin;              /* int(Rm / 0.002) (:58) */
} b2a_rdf_spec;

typedef struct b2a_orient_spec
{
    int   grp, com;          /* NCM (:153) */
    int   DIMN;              /* axis of the cylinder / slab filter (:156-172) */
    float low, up;           /* inclusive both ends (:158) */
    float c1, c2;            /* cylinder centre */
    float dr, CR, Cr;        /* radial bins: nbin1 = int((CR - Cr) / dr) in float (:133) */
    int   nAngle;            /* dAngle = 180 / nAngle in INTEGER arithmetic (:143) */
    int   tp1, tp2, tp3;     /* 1-based atoms in molecule (:14-29) */
    int   DIMNOrient;        /* wall vector axis (:144) */
    int   method;            /* 1 sum, 2 cross product (:184-191) */
} b2a_orient_spec;

typedef struct b2a_coord_spec
{
    int   grp0, grp1;        /* i runs over grp0, j over grp1 (calc_pair_dist_bulk.cpp:69,78) */
    int   com0, com1;        /* centre kind (the programs use -ncm for both, :66-67) */
    float Rmin;              /* counted when 1e-3 < distance <= Rmin (:81) */
    int   dim;
    float low, up;           /* lowPos <= posc1(i,dim) <= upPos, inclusive (:71) */
    int   use_cyl;           /* calc_pair_dist.cpp:85-86: also Rlow <= sqrt((x-cx)^2 + (y-cy)^2) <= Rup */
    float cx, cy, Rlow, Rup;
    int   max_count;         /* size of the histogram of counts; pairNumber >= max_count is counted in B2A_STAT_DROPPED */
} b2a_coord_spec;

/* stats slots common to all accumulators (u64 each) */
enum {
    B2A_STAT_FRAMES   = 0,  /* frames accumulated */
    B2A_STAT_DROPPED  = 1,  /* samples whose bin lies outside the array (memory corruption in the reference) */
    B2A_STAT_EXACT    = 2,  /* samples (rdf: pairs; density/orient: molecules) re-evaluated in FP64 because the FP32
                               bracket could not decide them */
    B2A_STAT_MOVED    = 3,  /* rdf: of those, pairs whose bin changed */
    B2A_STAT_SELECTED = 4,  /* molecules passing the region filter (summed over frames) */
    B2A_STAT_PAIRS    = 5,  /* rdf: pair distances physically evaluated */
    B2A_STAT_OVERFLOW = 6,  /* rdf: exact-queue overflows handled inline */
    B2A_STAT_MISPROVEN = 7, /* density/orient validation mode (b2a_acc_fast 2): samples whose FP32-proven result differs
                               from the FP64 evaluation; must be 0 */
    B2A_STAT_NSLOTS   = 8
};

int b2a_density_create(b2a_ctx* ctx, const b2a_density_spec* spec, int* acc);
int b2a_region_create(b2a_ctx* ctx, const b2a_region_spec* spec, int* acc);
/* out: [frames of the last accumulated batch][counters] (spec.per_frame must have been set) */
int b2a_region_read_frames(b2a_ctx* ctx, int acc, unsigned long long* out);
int b2a_rdf_create(b2a_ctx* ctx, const b2a_rdf_spec* spec, int* acc);
int b2a_orient_create(b2a_ctx* ctx, const b2a_orient_spec* spec, int* acc);

/* coordination numbers (mofs/calc_pair_dist_bulk.cpp:66-92, mofs/calc_pair_dist.cpp:80-106): counters[k] = number of
 * selected molecules with pairNumber == k, summed over frames; B2A_STAT_SELECTED = sum of numA.  The reference normalises
 * per frame (totPair[k] += tmpPair[k] / numA, :89-92), so the per-frame tables of the last accumulated batch are
 * readable too, and b2a_coord_fold performs that accumulation in frame order.                                       */
int b2a_coord_create(b2a_ctx* ctx, const b2a_coord_spec* spec, int* acc);
/* table: [frames of the last batch][max_count] tmpPair; numA: [frames] */
int b2a_coord_read_frames(b2a_ctx* ctx, int acc, uint32_t* table, uint32_t* numA);
/* pairNumber of every molecule of grp0 in one frame of the last batch (-1: filtered out) */
int b2a_coord_counts(b2a_ctx* ctx, int acc, int frame, int32_t* count);
int b2a_coord_fold(const uint32_t* table, const uint32_t* numA, int nframes, int max_count, double* totPair);

/* queue the accumulation of the whole current batch (asynchronous) */
int b2a_accumulate(b2a_ctx* ctx, int acc);
int b2a_acc_reset(b2a_ctx* ctx, int acc);

/* number of u64 counters (histogram cells) and their device address, for the frame-sharded merge:
 * one all-reduce (sum, 64-bit integer) over [dptr, dptr + ncounters + B2A_STAT_NSLOTS) per accumulator. */
int b2a_acc_size(b2a_ctx* ctx, int acc, size_t* ncounters);
/* number of u64 words b2a_acc_read writes into `counts` (the reference layout: rdf nbin*Rbin, orient nbin1*nAngle + nbin1, ...) */
int b2a_acc_public_size(b2a_ctx* ctx, int acc, size_t* ncounters);
int b2a_acc_device_ptr(b2a_ctx* ctx, int acc, void** dptr);

/* synchronise and copy out.  counts layouts are the reference's:
 *   density: counts[me]                         (itp::vecd)
 *   rdf:     counts[meZ + meR*nbin]             (itp::matd(nbin, Rbin), column-major)
 *   orient:  counts[meZ + meA*nbin1], then ncm[nbin1] appended (orient matd + NCMDens)
 * stats: B2A_STAT_NSLOTS values (may be NULL).                                                    */
int b2a_acc_read(b2a_ctx* ctx, int acc, uint64_t* counts, uint64_t* stats);

/* host post-processing exactly as calc_rdf_region.cpp:100-115 (shell volumes with 0.002, tail mean) */
int b2a_rdf_normalise(const uint64_t* counts, int nbin, int Rbin, double* g);

/* rdf tuning: 0 = auto.  symmetric=1 lets grp0==grp1 evaluate i<j only when every molecule passes the
 * region filter into one slab (counts are identical); paranoid=1 re-evaluates EVERY pair in FP64
 * (validation of the FP32 error bound; slow).                                                      */
int b2a_rdf_tune(b2a_ctx* ctx, int acc, int symmetric, int paranoid);
/* density / orientation accumulators: 1 (default) = every molecule is first evaluated in FP32 with a rigorous error
 * bound and only the undecided ones are re-evaluated by the literal FP64 code (counts are identical, the kernel reaches
 * the HBM roofline); 0 = literal FP64 for every molecule; 2 = validation: both, B2A_STAT_MISPROVEN counts disagreements.
 * The environment variable B2A_FAST overrides the mode.                                                             */
int b2a_acc_fast(b2a_ctx* ctx, int acc, int mode);
/* Orientation accumulators: let the same pass over the coordinates also materialise the exact centres of kind `com` of the
 * accumulator's group (bit-identical to b2a_centers: the reference's loadPositionCenter arithmetic), e.g. the per-molecule dipole
 * vectors (com 3) beside the orientation histogram (BASELINE config 4: one read of the frame instead of two).  After
 * b2a_accumulate the centres of the batch are cached: b2a_centers / b2a_centers_batch / b2a_centers_device return them without
 * another kernel.  com = -1 switches it off.                                                                                 */
int b2a_acc_also_centers(b2a_ctx* ctx, int acc, int com);
/* Cell-binned evaluation for cut-offs much shorter than the box (the brute-force loop of calc_rdf_region.cpp:73-95
 * wastes > 99 % of its pairs when Rm = 3 nm in a 34 nm box; pattern and its exactness check in the reference:
 * src/test/nb_cellSearch_anyBoxSize.cpp:194-299).  Counts are identical to the all-pairs evaluation.
 * mode: -1 automatic (default: used when it culls enough to win), 0 never, 1 whenever any cell can be skipped.    */
int b2a_rdf_cells(b2a_ctx* ctx, int acc, int mode);

/* ---- measurement support ------------------------------------------------------------------------ */
/* Forget the cached centres of the resident batch: the next b2a_accumulate / loader call recomputes the
 * whole path (used by bench.py so that a timed step never reuses a previous step's intermediate results). */
int b2a_invalidate(b2a_ctx* ctx);
/* Bracket every kernel family with CUDA events on the context stream.  b2a_timing_read synchronises, returns
 * total milliseconds and launch counts for slots {0: K1 centres, 1: K3 slab prep, 2: K3 pair kernel,
 * 3: K2 density, 4: K4 orientation, 5: H2D copies, 6: K5 coordination pair kernel} (arrays of 8) and resets the totals.               */
int b2a_timing_enable(b2a_ctx* ctx, int enable);
int b2a_timing_read(b2a_ctx* ctx, double* ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* B2A_H */
