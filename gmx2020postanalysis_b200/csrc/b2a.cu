// libb2a.so -- C ABI (include/b2a.h) over the sm_100a kernels.  No CPU fallback: without a CUDA device
// b2a_create fails with B2A_ERR_CUDA.
#include "b2a.h"

#include "b2a_kernels_rdf.cuh"
#include "b2a_kernels_rdf_cell.cuh"
#include "b2a_kernels_coord.cuh"
#include "b2a_kernels_stream.cuh"   // pulls in the centre / orientation kernels

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace b2a;

namespace
{
thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                               \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) return fail(B2A_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Sum in the order Eigen 3.3.9's vectorised Array<double>::sum() uses on SSE2 (two packet accumulators of
// two doubles, then the scalar tail) -- this is what `atomCenter.sum()` evaluates in the reference
// (include/itp/gmx_bits/gmx_handle.ipp:195), and for napm > 3 it differs from a sequential sum in the last bit.
// host side of div_const (b2a_kernels_center.cuh): y = RN(1/d); the fast path needs d and y normal, and Markstein's
// theorem excludes divisors whose significand is all ones
DivConst make_div(double d)
{
    DivConst c{};
    c.d   = d;
    c.inv = 1.0 / d;
    uint64_t bits;
    std::memcpy(&bits, &d, sizeof(bits));
    const bool all_ones = (bits & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull;
    c.fast = (std::isnormal(d) && std::isnormal(c.inv) && !all_ones) ? 1 : 0;
    return c;
}

double eigen_order_sum(const double* w, int n)
{
    const int a2 = (n / 4) * 4, a1 = (n / 2) * 2;
    if (!a1)
    {
        double r = w[0];
        for (int i = 1; i < n; ++i) r += w[i];
        return r;
    }
    double p00 = w[0], p01 = w[1];
    if (a1 > 2)
    {
        double p10 = w[2], p11 = w[3];
        for (int i = 4; i < a2; i += 4) { p00 += w[i]; p01 += w[i + 1]; p10 += w[i + 2]; p11 += w[i + 3]; }
        p00 += p10; p01 += p11;
        if (a1 > a2) { p00 += w[a2]; p01 += w[a2 + 1]; }
    }
    double r = p00 + p01;
    for (int i = a1; i < n; ++i) r += w[i];
    return r;
}

struct Group
{
    int     ngx = 0, napm = 0, nmol = 0;
    int*    d_index = nullptr;
    std::vector<int> h_index;
    // heterogeneous molecule sizes (GmxHandleFull): CSR offsets, per-atom weights, per-molecule totals
    bool      csr = false;
    int       max_napm = 0, min_napm = 0;
    int*      d_off = nullptr; int* d_mol_of = nullptr;
    double*   d_wfull[4] = { nullptr, nullptr, nullptr, nullptr };
    DivConst* d_div[4]   = { nullptr, nullptr, nullptr, nullptr };
    bool    contiguous = false;   // index[n] == index[0] + n: a block's molecules are one byte range per frame (TMA path)
    double* d_w[4]  = { nullptr, nullptr, nullptr, nullptr };   // mass, ones, charge, charge(dipole)
    float*  d_wf[4] = { nullptr, nullptr, nullptr, nullptr };   // the same, rounded to float (FP32-bracketed fast paths)
    std::vector<double> h_w[4];
    double  wsum[4] = { 0, 0, 0, 0 };
    // centre cache per com for the current batch
    double*  d_cen[4]   = { nullptr, nullptr, nullptr, nullptr };
    int4*    d_fix[4]   = { nullptr, nullptr, nullptr, nullptr };
    uint64_t valid[4]   = { 0, 0, 0, 0 };   // batch serial the cache belongs to
    uint64_t fix_valid[4] = { 0, 0, 0, 0 }; // ... and the fixed-point copy (written only for consumers that read it: K3 / K5)
};

// One group's molecules counting-sorted by cell (K3 cell mode), shared by every accumulator of the same batch that needs the
// same sort: a sweep over all pairs of G groups sorts each group once (per slab signature), not once per pair.
struct CellSort
{
    int      grp = -1, com = 0, bits = 0, with_slab = 0, nbin = 0, dim = 0, dim2 = 0;
    double   low = 0, up = 0, low2 = 0, up2 = 0;     // slab signature (only when with_slab)
    uint64_t serial = 0;                             // batch serial the arrays were built from (0: never)
    int*      keys   = nullptr;                      // [maxK][n]
    unsigned* cells  = nullptr;                      // [3][maxK][nkeys + 1]: counts, starts, cursors
    int4*     sorted = nullptr;                      // [maxK][n]
    size_t    cap_n = 0;
    unsigned  cap_keys = 0, nkeys = 0;
    cudaEvent_t  ready = nullptr;                    // recorded on the stream that built the arrays; other streams wait for it
    cudaStream_t built_on = nullptr;
    unsigned* starts(int maxK) const { return cells + (size_t)(nkeys + 1) * maxK; }
};

enum AccKind { ACC_DENSITY, ACC_RDF, ACC_ORIENT, ACC_COORD, ACC_REGION };

// Per-frame constants computed on the host travel through a small ring of pinned slots: slot s may be rewritten once the
// upload that last read it has finished (an event per slot, normally long passed), so queueing an accumulation never
// synchronises the stream and the host can run kFrameRing batches ahead of the device per accumulator.
constexpr int kFrameRing = 4;
struct FrameRing
{
    cudaEvent_t ev[kFrameRing] = { nullptr, nullptr, nullptr, nullptr };
    int         next = 0;
};

struct Acc
{
    AccKind            kind;
    b2a_density_spec   den{};
    b2a_rdf_spec       rdf{};
    b2a_orient_spec    ori{};
    b2a_coord_spec     coord{};
    b2a_region_spec    region{};
    unsigned long long* d_perframe = nullptr;   // region accumulators with per_frame: the per-frame log [series_cap][ncounters]
    int                perframe_frames = 0;     // frames of the last accumulated batch (the tail of the log)
    size_t             ncounters = 0;       // public counters (reference layout size)
    size_t             ndev      = 0;       // device u64 words before the stats block
    unsigned long long* d_hist   = nullptr; // [ndev + NSLOTS]: this device's PRIVATE counts
    // b2a_finalize (b2a_multi.cuh): sum of d_hist over every device of the communicator, valid while merged_epoch == epoch
    unsigned long long* d_merged = nullptr;
    uint64_t           epoch = 1, merged_epoch = 0;   // epoch moves with every accumulate / reset
    // per-frame series (region accumulators with per_frame, coordination numbers): the per-frame tables of EVERY accumulated
    // batch, appended on the device in accumulation order; `batches` holds (global index of the batch's first frame, frames)
    size_t             series_cap = 0, series_len = 0;          // frames
    unsigned*          d_series_numA = nullptr;                 // coord: numA of every logged frame
    std::vector<std::pair<long long, int>> batches;
    std::vector<long long>          gathered_idx;               // b2a_finalize across processes: the merged, ordered series
    std::vector<unsigned long long> gathered;
    uint64_t                        gathered_epoch = 0;
    // rdf scratch
    int*      d_slab_id = nullptr;
    unsigned* d_slab_cnt = nullptr;   // [maxK][nslab] counts then [maxK][nslab] cursors
    int4*     d_sortedA = nullptr;
    int4*     sortedA_view = nullptr;   // what the next k3_rdf launch reads: d_sortedA (slab sort) or a shared CellSort array
    RdfFrame* d_frames = nullptr;
    RdfFrame* h_frames = nullptr;     // pinned, [kFrameRing][maxK]: b2a_accumulate never waits for the device (FrameRing)
    FrameRing ring;
    int       alloc_n0 = 0, alloc_n1 = 0;   // molecule counts the per-molecule scratch below was sized for (b2a_*_create)
    int       symmetric = 0, paranoid = 0, cells = -1;   // cells: -1 auto, 0 never, 1 whenever it culls anything
    // cell mode scratch (allocated on first use)
    int*      d_keysA = nullptr; int* d_keysB = nullptr;
    unsigned* d_cellA = nullptr; unsigned* d_cellB = nullptr;   // [3][maxK][nkeys + 1]: counts, starts, cursors
    int4*     d_sortedB = nullptr;
    unsigned  cap_keysA = 0, cap_keysB = 0;
    int       ipt = 0, threads = 0;   // 0 = auto
    int       fast = 1;               // density / orient: 0 literal FP64, 1 FP32 bracket + FP64 queue, 2 validate
    // RDF accumulators of a context that holds several of them (a sweep over group pairs) queue their kernels on a stream of their
    // own, forked from the context stream after the centres and joined before anything that depends on them: the short launches
    // of small group pairs then run beside the long ones instead of in front of them
    cudaStream_t side = nullptr;
    cudaEvent_t  ev_done = nullptr;
    bool         side_pending = false;
    cudaStream_t last_stream = nullptr;   // stream the last accumulation was queued on
    int       also_com = -1;          // orient: centre kind the streaming kernel materialises on the side (b2a_acc_also_centers)
    unsigned char* d_lut = nullptr;   // orient: cos cell -> first candidate angle bin
    float2*        d_fmh = nullptr;   // orient: per angle bin {midpoint, shrunk halfwidth} of its cosine interval
    // coord
    int*        d_cnt = nullptr;      // [maxK][n0] pair counts of the last batch
    unsigned*   d_table = nullptr;    // per-frame log [series_cap][max_count]: histogram of counts of every accumulated frame
    CoordFrame* d_cframes = nullptr;
    CoordFrame* h_cframes = nullptr;  // pinned, [kFrameRing][maxK]
    int         table_frames = 0;     // frames of the batch d_table describes
    // orient
    double* d_cosedge = nullptr;
    float*  d_fedge   = nullptr;
    double* d_stab    = nullptr;
    int     nbin1 = 0;
};
} // namespace

struct b2a_ctx
{
    int          device = 0;
    cudaStream_t stream = nullptr;
    int          natoms = 0, maxK = 0;
    // The frame batch is double-buffered: b2a_submit_batch copies into the buffer the kernels of the previous batch are NOT
    // reading, on its own stream, so the H2D copy of batch k+1 overlaps the kernels of batch k.  d_x / d_geom / h_geom always
    // point at the CURRENT batch (what every kernel launched after the submit reads).
    float*       d_x    = nullptr;
    FrameGeom*   d_geom = nullptr;
    FrameGeom*   h_geom = nullptr;   // pinned
    float*       d_xbuf[2]    = { nullptr, nullptr };
    FrameGeom*   d_geombuf[2] = { nullptr, nullptr };
    FrameGeom*   h_geombuf[2] = { nullptr, nullptr };
    cudaStream_t copy_stream  = nullptr;
    cudaEvent_t  ev_copied[2] = { nullptr, nullptr };   // copy stream: the upload into buffer b has finished
    cudaEvent_t  ev_free[2]   = { nullptr, nullptr };   // main stream: every kernel launched while b was current has finished
    int          cur = 0;
    cudaEvent_t  ev_field = nullptr;                    // main stream: the last velocity / force upload has finished
    cudaEvent_t  ev_fork = nullptr;                     // main stream: fork point of an accumulator's side stream
    bool         field_pending = false;
    std::vector<float> h_box;        // [K][9]
    int          nframes = 0;
    uint64_t     serial  = 0;
    int          sm_count = 148;
    std::map<int, Group> groups;
    std::vector<std::unique_ptr<Acc>> accs;
    std::vector<std::unique_ptr<CellSort>> cell_sorts;
    double*      d_pos_scratch = nullptr;
    unsigned*    d_work = nullptr;   // work-item counter of the persistent streaming kernels
    // velocities / forces of the current batch (fr->v, fr->f), uploaded on demand; [0] is unused (x lives in d_x)
    float*       d_field[3] = { nullptr, nullptr, nullptr };
    uint64_t     field_serial[3] = { 0, 0, 0 };   // batch serial the upload belongs to
    double*      d_field_cen = nullptr;
    size_t       field_cen_n = 0;
    size_t       pos_scratch_n = 0;
    // optional CUDA-event timing per kernel family (bench.py): pairs are resolved lazily in b2a_timing_read
    bool         timing = false;
    struct Span { int slot; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double       t_ms[8] = { 0 };
    long long    t_n[8]  = { 0 };
    unsigned long long* d_dbg = nullptr;   // B2A_STREAM_DEBUG: per-block time stamps of the last streaming launch (8 words x 4096 blocks)
    // ---- multi-GPU (b2a_multi.cuh) ----------------------------------------------------------------------------
    // A context made by b2a_create_multi is a LEADER: it owns one member context per device and no device resources
    // of its own; every entry point forwards (to all members, to the member holding the current batch, or to member 0).
    std::vector<b2a_ctx*> members;
    int          cur_member = -1;     // leader: member that received the last batch
    long long    frame_base = 0;      // global index of frame 0 of the current batch (per-frame series are ordered by it)
    long long    next_frame = 0;      // global index the next submitted batch starts at
    void*        comm = nullptr;      // ncclComm_t of this device (members and plain contexts), attached by b2a_comm_init / b2a_create_multi
    int          comm_rank = 0, comm_size = 1, comm_procs = 1;
    int          host_buffers = 2;    // leader: host staging buffers the caller cycles through (b2a_set_host_buffers)
    std::vector<std::pair<int, int>> inflight;   // leader: uploads not yet known to have completed, oldest first: (member, device buffer)
    bool         same_device_team = false;   // leader: members share a device (tests on one GPU): merged by a local kernel, NCCL cannot do it
    bool is_leader() const { return !members.empty(); }
};

// forwarding of the C ABI on a leader context
#define TEAM_ALL(c, call)                                                                    \
    if ((c) && (c)->is_leader())                                                             \
    {                                                                                        \
        for (b2a_ctx* m : (c)->members) { if (int rc_ = (call)) return rc_; }                \
        return B2A_OK;                                                                       \
    }
#define TEAM_CUR(c, call)                                                                    \
    if ((c) && (c)->is_leader())                                                             \
    {                                                                                        \
        if ((c)->cur_member < 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet"); \
        b2a_ctx* m = (c)->members[(c)->cur_member];                                          \
        return (call);                                                                       \
    }
#define TEAM_FIRST(c, call)                                                                  \
    if ((c) && (c)->is_leader())                                                             \
    {                                                                                        \
        b2a_ctx* m = (c)->members[0];                                                        \
        return (call);                                                                       \
    }

namespace
{
struct Timed   // RAII: records an event pair around a launch sequence when timing is on
{
    b2a_ctx* c; int slot; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    Timed(b2a_ctx* c_, int slot_, cudaStream_t on = nullptr) : c(c_), slot(slot_), st(on ? on : c_->stream)
    {
        if (!c->timing) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~Timed()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        c->spans.push_back({ slot, a, b });
    }
};
} // namespace

__global__ void k_fetch_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src_pinned, int nwords);

// The context stream waits for every accumulator that has work on a side stream: called before anything that must see (or may
// overwrite what is read by) that work -- a new batch, an invalidation, a read, a reset, the merge.
static int join_side(b2a_ctx* c)
{
    for (auto& a : c->accs)
        if (a && a->side_pending)
        {
            CU(cudaStreamWaitEvent(c->stream, a->ev_done, 0));
            a->side_pending = false;
        }
    return B2A_OK;
}

extern "C" const char* b2a_last_error(void) { return g_err.c_str(); }
extern "C" const char* b2a_version(void) { return "b2a 0.1 (sm_100a)"; }

extern "C" int b2a_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void b2a_destroy(b2a_ctx* c);
static void multi_release(b2a_ctx* c);                    // b2a_multi.cuh: communicator of a context
static int  team_submit(b2a_ctx* c, int nframes, const float* x, const float* box9);
static int  team_read(b2a_ctx* c, int id, uint64_t* counts, uint64_t* stats);

extern "C" int b2a_create(b2a_ctx** out, int device, int natoms, int max_frames)
{
    if (!out || natoms <= 0 || max_frames <= 0) return fail(B2A_ERR_ARG, "b2a_create: natoms and max_frames must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return fail(B2A_ERR_CUDA, "b2a_create: no CUDA device visible; libb2a has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(B2A_ERR_ARG, "b2a_create: device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    // a context that fails half-way through (e.g. the batch does not fit in device memory) is torn down completely
    struct Undo { void operator()(b2a_ctx* p) const { b2a_destroy(p); } };
    std::unique_ptr<b2a_ctx, Undo> c(new b2a_ctx);
    c->device = device;
    c->natoms = natoms;
    c->maxK   = max_frames;
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b)
    {
        CU(cudaMalloc(&c->d_xbuf[b], (size_t)natoms * 3 * sizeof(float) * max_frames + 64));   // +64: bulk copies are widened to 16-byte boundaries
        CU(cudaMalloc(&c->d_geombuf[b], sizeof(FrameGeom) * max_frames));
        CU(cudaMallocHost(&c->h_geombuf[b], sizeof(FrameGeom) * max_frames));
        CU(cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_free[b], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_field, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    c->cur = 0; c->d_x = c->d_xbuf[0]; c->d_geom = c->d_geombuf[0]; c->h_geom = c->h_geombuf[0];
    c->h_box.assign((size_t)max_frames * 9, 0.f);
    *out = c.release();
    return B2A_OK;
}

static void free_group(Group& g)
{
    cudaFree(g.d_index);
    for (int k = 0; k < 4; ++k) { cudaFree(g.d_w[k]); cudaFree(g.d_wf[k]); cudaFree(g.d_cen[k]); cudaFree(g.d_fix[k]); cudaFree(g.d_wfull[k]); cudaFree(g.d_div[k]); }
    cudaFree(g.d_off); cudaFree(g.d_mol_of);
    g = Group();
}

static void free_acc(Acc& a)
{
    cudaFree(a.d_merged); cudaFree(a.d_series_numA);
    cudaFree(a.d_hist); cudaFree(a.d_slab_id); cudaFree(a.d_slab_cnt); cudaFree(a.d_sortedA); cudaFree(a.d_frames);
    cudaFree(a.d_cosedge); cudaFree(a.d_fedge); cudaFree(a.d_stab); cudaFree(a.d_lut); cudaFree(a.d_fmh);
    cudaFree(a.d_cnt); cudaFree(a.d_table); cudaFree(a.d_cframes); cudaFree(a.d_perframe);
    if (a.h_cframes) cudaFreeHost(a.h_cframes);
    cudaFree(a.d_keysA); cudaFree(a.d_keysB); cudaFree(a.d_cellA); cudaFree(a.d_cellB); cudaFree(a.d_sortedB);
    if (a.h_frames) cudaFreeHost(a.h_frames);
    for (auto& e : a.ring.ev) if (e) cudaEventDestroy(e);
    if (a.side) { cudaStreamSynchronize(a.side); cudaStreamDestroy(a.side); }
    if (a.ev_done) cudaEventDestroy(a.ev_done);
}

extern "C" void b2a_destroy(b2a_ctx* c)
{
    if (!c) return;
    if (c->is_leader())
    {
        for (b2a_ctx* m : c->members) b2a_destroy(m);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    multi_release(c);
    for (auto& kv : c->groups) free_group(kv.second);
    for (auto& a : c->accs) if (a) free_acc(*a);
    for (auto& cs : c->cell_sorts) { cudaFree(cs->keys); cudaFree(cs->cells); cudaFree(cs->sorted); if (cs->ready) cudaEventDestroy(cs->ready); }
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_pos_scratch); cudaFree(c->d_work); cudaFree(c->d_dbg);
    cudaFree(c->d_field[1]); cudaFree(c->d_field[2]); cudaFree(c->d_field_cen);
    for (int b = 0; b < 2; ++b)
    {
        cudaFree(c->d_xbuf[b]); cudaFree(c->d_geombuf[b]);
        if (c->h_geombuf[b]) cudaFreeHost(c->h_geombuf[b]);
        if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]);
        if (c->ev_free[b]) cudaEventDestroy(c->ev_free[b]);
    }
    if (c->ev_field) cudaEventDestroy(c->ev_field);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();   // a context torn down after a failed allocation must not leave a sticky error behind
    delete c;
}

extern "C" int b2a_sync(b2a_ctx* c)
{
    TEAM_ALL(c, b2a_sync(m));
    if (!c) return fail(B2A_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

extern "C" int b2a_join(b2a_ctx* c)
{
    TEAM_ALL(c, b2a_join(m));
    if (!c) return fail(B2A_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    return join_side(c);
}

extern "C" void* b2a_stream(b2a_ctx* c)
{
    if (c && c->is_leader()) return (void*)c->members[0]->stream;   // per-device streams: b2a_member(ctx, i)
    return c ? (void*)c->stream : nullptr;
}

extern "C" void* b2a_pinned_alloc(size_t n)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, n) != cudaSuccess) { cudaGetLastError(); g_err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
extern "C" void b2a_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// ---- groups --------------------------------------------------------------------------------------------------
extern "C" int b2a_set_group(b2a_ctx* c, int grp, const int* index, int ngx, int napm, const double* mass, const double* charge)
{
    TEAM_ALL(c, b2a_set_group(m, grp, index, ngx, napm, mass, charge));
    if (!c || !index || !mass || !charge) return fail(B2A_ERR_ARG, "b2a_set_group: null argument");
    if (grp < 0) return fail(B2A_ERR_ARG, "grp(%d) should not be negative", grp);
    if (napm <= 0 || ngx <= 0 || ngx % napm != 0)
        return fail(B2A_ERR_ARG, "The index group does not consist of whole molecules");   // ipp:432
    const int nmol = ngx / napm;
    for (int i = 0; i < nmol; ++i)
    {
        const int m = index[(size_t)i * napm];
        if (m < 0 || m + napm > c->natoms) return fail(B2A_ERR_ARG, "b2a_set_group: atom index out of range in molecule %d", i);
    }
    for (int n = 0; n < ngx; ++n)
        if (index[n] < 0 || index[n] >= c->natoms) return fail(B2A_ERR_ARG, "b2a_set_group: atom index %d out of range", index[n]);
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    Group& g = c->groups[grp];
    free_group(g);
    for (auto& e : c->cell_sorts) if (e->grp == grp) e->serial = 0;   // sorted views of the old group are stale
    g.ngx = ngx; g.napm = napm; g.nmol = nmol;
    g.h_index.assign(index, index + ngx);
    g.contiguous = true;
    for (int n = 1; n < ngx && g.contiguous; ++n) g.contiguous = (index[n] == index[0] + n);
    CU(cudaMalloc(&g.d_index, sizeof(int) * ngx));
    CU(cudaMemcpy(g.d_index, index, sizeof(int) * ngx, cudaMemcpyHostToDevice));
    std::vector<double> ones(napm, 1.0);
    const double*       src[4] = { mass, ones.data(), charge, charge };
    for (int k = 0; k < 4; ++k)
    {
        CU(cudaMalloc(&g.d_w[k], sizeof(double) * napm));
        CU(cudaMemcpy(g.d_w[k], src[k], sizeof(double) * napm, cudaMemcpyHostToDevice));
        g.h_w[k].assign(src[k], src[k] + napm);
        std::vector<float> wf(napm);
        for (int j = 0; j < napm; ++j) wf[j] = (float)src[k][j];
        CU(cudaMalloc(&g.d_wf[k], sizeof(float) * napm));
        CU(cudaMemcpy(g.d_wf[k], wf.data(), sizeof(float) * napm, cudaMemcpyHostToDevice));
    }
    g.wsum[0] = eigen_order_sum(mass, napm);
    g.wsum[1] = eigen_order_sum(ones.data(), napm);
    g.wsum[2] = eigen_order_sum(charge, napm);
    double s = 0.0;   // general/9g_dist_bin_1803.cpp:217-221: sequential sum, |sum| < 1e-6 -> 1
    for (int j = 0; j < napm; ++j) s += charge[j];
    g.wsum[3] = (std::fabs(s) < 1e-6) ? 1.0 : s;
    return B2A_OK;
}

extern "C" int b2a_set_group_full(b2a_ctx* c, int grp, const int* index, int ngx, int nmol, const int* napm, const double* mass,
                                  const double* charge)
{
    TEAM_ALL(c, b2a_set_group_full(m, grp, index, ngx, nmol, napm, mass, charge));
    if (!c || !index || !napm || !mass || !charge) return fail(B2A_ERR_ARG, "b2a_set_group_full: null argument");
    if (grp < 0 || ngx <= 0 || nmol <= 0) return fail(B2A_ERR_ARG, "b2a_set_group_full: bad sizes");
    std::vector<int> off(nmol + 1, 0), mol_of(ngx);
    int              maxn = 0;
    for (int i = 0; i < nmol; ++i)
    {
        if (napm[i] <= 0) return fail(B2A_ERR_ARG, "b2a_set_group_full: molecule %d has no atoms", i);
        off[i + 1] = off[i] + napm[i];
        maxn       = std::max(maxn, napm[i]);
    }
    if (off[nmol] != ngx) return fail(B2A_ERR_ARG, "b2a_set_group_full: sum of napm (%d) differs from the group size (%d)", off[nmol], ngx);
    for (int i = 0; i < nmol; ++i)
    {
        const int m = index[off[i]];
        if (m < 0 || m + napm[i] > c->natoms) return fail(B2A_ERR_ARG, "b2a_set_group_full: atom index out of range in molecule %d", i);
        for (int n = off[i]; n < off[i + 1]; ++n) mol_of[n] = i;
    }
    for (int n = 0; n < ngx; ++n)
        if (index[n] < 0 || index[n] >= c->natoms) return fail(B2A_ERR_ARG, "b2a_set_group_full: atom index %d out of range", index[n]);
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    Group& g = c->groups[grp];
    free_group(g);
    for (auto& e : c->cell_sorts) if (e->grp == grp) e->serial = 0;   // sorted views of the old group are stale
    g.ngx = ngx; g.napm = 0; g.nmol = nmol; g.csr = true; g.max_napm = maxn;
    g.min_napm = maxn;
    for (int i = 0; i < nmol; ++i) g.min_napm = std::min(g.min_napm, napm[i]);
    g.h_index.assign(index, index + ngx);
    CU(cudaMalloc(&g.d_index, sizeof(int) * ngx));
    CU(cudaMemcpy(g.d_index, index, sizeof(int) * ngx, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&g.d_off, sizeof(int) * (nmol + 1)));
    CU(cudaMemcpy(g.d_off, off.data(), sizeof(int) * (nmol + 1), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&g.d_mol_of, sizeof(int) * ngx));
    CU(cudaMemcpy(g.d_mol_of, mol_of.data(), sizeof(int) * ngx, cudaMemcpyHostToDevice));
    std::vector<double> ones(ngx, 1.0);
    const double*       src[4] = { mass, ones.data(), charge, charge };
    for (int k = 0; k < 4; ++k)
    {
        // per-molecule totals, summed sequentially like totMass / totCharge (gmx_handle_full.ipp:337-340); geometry = napm
        std::vector<DivConst> div(nmol);
        for (int i = 0; i < nmol; ++i)
        {
            double t = 0.0;
            for (int n = off[i]; n < off[i + 1]; ++n) t += src[k][n];
            if (k == 1) t = (double)napm[i];
            if (k == 3 && std::fabs(t) < 1e-6) t = 1.0;   // 9g_dist_bin_1803.cpp:221
            div[i] = make_div(t);
        }
        CU(cudaMalloc(&g.d_wfull[k], sizeof(double) * ngx));
        CU(cudaMemcpy(g.d_wfull[k], src[k], sizeof(double) * ngx, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&g.d_div[k], sizeof(DivConst) * nmol));
        CU(cudaMemcpy(g.d_div[k], div.data(), sizeof(DivConst) * nmol, cudaMemcpyHostToDevice));
    }
    return B2A_OK;
}

static int get_group(b2a_ctx* c, int grp, Group** out)
{
    auto it = c->groups.find(grp);
    if (it == c->groups.end()) return fail(B2A_ERR_ARG, "grp(%d) should less than ngrps(%d)!", grp, (int)c->groups.size());   // ipp:101
    *out = &it->second;
    return B2A_OK;
}

extern "C" int b2a_set_group_wsum(b2a_ctx* c, int grp, int com, double wsum)
{
    TEAM_ALL(c, b2a_set_group_wsum(m, grp, com, wsum));
    if (!c || com < 0 || com > 3) return fail(B2A_ERR_ARG, "b2a_set_group_wsum: bad argument");
    Group* g = nullptr;
    if (int rc = get_group(c, grp, &g)) return rc;
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    g->wsum[com]  = wsum;
    g->valid[com] = 0;
    for (auto& e : c->cell_sorts) if (e->grp == grp) e->serial = 0;
    return B2A_OK;
}

extern "C" int b2a_group_nmol(b2a_ctx* c, int grp)
{
    TEAM_FIRST(c, b2a_group_nmol(m, grp));
    if (!c) return -1;
    auto it = c->groups.find(grp);
    return it == c->groups.end() ? -1 : it->second.nmol;
}

// ---- frames --------------------------------------------------------------------------------------------------
extern "C" int b2a_submit_batch(b2a_ctx* c, int nframes, const float* x, const float* box9)
{
    if (!c || !x || !box9) return fail(B2A_ERR_ARG, "b2a_submit_batch: null argument");
    if (c->is_leader()) return team_submit(c, nframes, x, box9);
    if (nframes <= 0 || nframes > c->maxK) return fail(B2A_ERR_ARG, "b2a_submit_batch: nframes %d outside 1..%d", nframes, c->maxK);
    CU(cudaSetDevice(c->device));
    for (int f = 0; f < nframes; ++f)
        for (int k = 0; k < 3; ++k)
            if (!(box9[(size_t)f * 9 + 4 * k] > 0.f)) return fail(B2A_ERR_ARG, "b2a_submit_batch: frame %d has non-positive box length", f);
    // The previous upload must have left the caller's pinned buffer before the caller may refill it (callers alternate
    // between two host buffers); this also bounds how far the host can run ahead of the device.
    CU(cudaEventSynchronize(c->ev_copied[c->cur]));
    if (c->field_pending) { CU(cudaEventSynchronize(c->ev_field)); c->field_pending = false; }
    // Everything launched so far read the current buffer; from here on the other one becomes current.  Its last readers
    // (two submits ago) are covered by ev_free, its last upload by ev_copied (which also frees its pinned geometry staging).
    if (int rc = join_side(c)) return rc;   // side-stream kernels read the current buffer too
    CU(cudaEventRecord(c->ev_free[c->cur], c->stream));
    const int b = c->cur ^ 1;
    CU(cudaEventSynchronize(c->ev_copied[b]));
    c->cur = b; c->d_x = c->d_xbuf[b]; c->d_geom = c->d_geombuf[b]; c->h_geom = c->h_geombuf[b];
    for (int f = 0; f < nframes; ++f)
    {
        FrameGeom& g = c->h_geom[f];
        for (int k = 0; k < 3; ++k)
        {
            const float bk = box9[(size_t)f * 9 + 4 * k];
            g.L[k]        = (double)bk;            // ipp:72-74
            g.half[k]     = g.L[k] / 2;            // ipp:128
            g.inv_unit[k] = 4294967296.0 / g.L[k];
            g.hsafe[k]    = (float)(g.half[k] * (1.0 - 1e-6));
            g.ff.L[k]     = bk;
            g.ff.invL[k]  = (float)(1.0 / g.L[k]);
            g.ff.hs2[k]   = (float)(g.half[k] * (1.0 - 1e-5));
            g.fk.negL2[2 * k] = g.fk.negL2[2 * k + 1] = -bk;
            g.fk.invL2[2 * k] = g.fk.invL2[2 * k + 1] = g.ff.invL[k];
        }
        {   // one-threshold variant of the shift proof (orient_fast3): |u^| < min_k hs2 and |d| <= 2 max_k L keep the FP32 error
            // of an offset below eps * 2.5 Lmax, which must stay well inside the 1e-5 * L_k / 2 margin of hs2 for every k
            const float Lmax = std::max(g.ff.L[0], std::max(g.ff.L[1], g.ff.L[2])), Lmin = std::min(g.ff.L[0], std::min(g.ff.L[1], g.ff.L[2]));
            g.fk.hsmin = (Lmax < 30.f * Lmin) ? std::min(g.ff.hs2[0], std::min(g.ff.hs2[1], g.ff.hs2[2])) : 0.f;
            g.fk.dmax  = 2.f * Lmax;
        }
    }
    std::memcpy(c->h_box.data(), box9, sizeof(float) * 9 * (size_t)nframes);
    {
        // upload on the copy stream (overlaps the kernels of the previous batch, which read the other buffer); kernels
        // launched after this call wait for it through ev_copied.  `x` must stay untouched until then, as before.
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_free[b], 0));
        Timed t(c, 5, c->copy_stream);   // the copies run on the copy stream: that is where the event pair belongs
        CU(cudaMemcpyAsync(c->d_x, x, (size_t)c->natoms * 3 * sizeof(float) * nframes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaMemcpyAsync(c->d_geom, c->h_geom, sizeof(FrameGeom) * nframes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
    }
    c->nframes = nframes;
    c->serial++;
    c->frame_base = c->next_frame;
    c->next_frame += nframes;
    return B2A_OK;
}

extern "C" int b2a_batch_frames(b2a_ctx* c)
{
    if (c && c->is_leader()) return c->cur_member < 0 ? 0 : c->members[c->cur_member]->nframes;
    return c ? c->nframes : -1;
}

// ---- TMA-staged streaming launch (contiguous groups) -----------------------------------------------------------
static bool stream_block(const Group& g, int* M)
{
    if (!g.contiguous || g.csr) return false;
    if (const char* e = getenv("B2A_NO_TMA")) if (atoi(e)) return false;
    int m = (32768 / (g.napm * 12) / 32) * 32;   // <= 32 KB of coordinates per stage
    m     = std::min(m, 256);
    if (m < 32) return false;
    *M = m;
    return true;
}

// Constants of the FP32 bracket (b2a_kernels_fast.cuh).  The fast path assumes centre = x_0 + sum w_j u_j / W, i.e. that
// W is the sum of the weights (not the guarded divisor of a neutral molecule's dipole), and that rho = sum|w| / |W| is
// moderate (charge centres of nearly neutral molecules amplify every rounding error); otherwise it stays off.
static FastParams fast_params(const Group& g, int com, int mode)
{
    FastParams fp{};
    if (const char* e = getenv("B2A_FAST")) mode = atoi(e);
    if (mode <= 0 || g.csr || g.h_w[com].empty()) return fp;
    const double W = g.wsum[com];
    double       sum = 0.0, sabs = 0.0;
    for (double w : g.h_w[com]) { sum += w; sabs += std::fabs(w); }
    if (!std::isnormal(W) || !(std::fabs(sum / W - 1.0) < 1e-12)) return fp;
    const double rho = sabs / std::fabs(W);
    if (!(rho < 1e3)) return fp;
    const double eps = std::ldexp(1.0, -24);
    fp.enabled = mode >= 2 ? 2 : 1;
    fp.invW    = (float)(1.0 / W);
    fp.e1      = (float)(1.05 * eps);
    fp.e2      = (float)(1.05 * eps * rho);
    fp.e3      = (float)(1.05 * eps * rho * (g.napm + 4));
    return fp;
}

template <int MODE, int NAPM_T, int VAR = 0, bool FUSE = false, int SPEC = 0>
static int launch_stream_t(b2a_ctx* c, const Group& g, int com, int M, StreamArgs a, size_t extra_smem, int fast_mode)
{
    if constexpr (MODE == 2 && !FUSE && SPEC == 0)
    {
        if (a.w2) return launch_stream_t<MODE, NAPM_T, VAR, true>(c, g, com, M, a, extra_smem, fast_mode);   // b2a_acc_also_centers
        // the program's default axes and method (calc_orient_mof.cpp:86-92): the specialised instantiation of the two-molecule path
        if constexpr (VAR == 1)
            if (a.ori.DIMN == 2 && a.ori.DIMNOrient == 2 && a.ori.method == 1 && !getenv("B2A_NO_SPEC"))
                return launch_stream_t<MODE, NAPM_T, VAR, false, 1>(c, g, com, M, a, extra_smem, fast_mode);
    }
    a.x = c->d_x; a.frame_stride = (size_t)c->natoms * 3; a.nframes = c->nframes;
    a.atom0 = g.h_index[0]; a.nmol = g.nmol; a.napm = g.napm;
    a.w = g.d_w[com]; a.wsum = make_div(g.wsum[com]); a.geom = c->d_geom;
    // molecules per tile: MODE 2 / VAR 1 evaluates two per consumer thread, MODE 0 / VAR 2 spends three lanes on each
    const int MT  = (MODE == 2 && VAR == 1) ? 2 * M : ((MODE == 0 && VAR == 2) ? 10 * (M / 32) : M);
    a.tile_floats = ((MT * g.napm * 3 + 8) + 3) & ~3;
    a.wf = g.d_wf[com];
    a.fp = fast_params(g, com, MODE == 0 ? 0 : fast_mode);
    // work items: (molecule block, 4 consecutive frames); persistent blocks pull them from a counter
    a.mblocks     = (g.nmol + MT - 1) / MT;
    a.item_frames = 4;
    if (const char* e = getenv("B2A_ITEM_FRAMES")) a.item_frames = std::max(1, atoi(e));
    a.nitems      = a.mblocks * ((c->nframes + a.item_frames - 1) / a.item_frames);
    if (!c->d_work)
    {   // {next work item, blocks finished}: the last block of a launch puts both back to zero, so no memset precedes a launch
        CU(cudaMalloc(&c->d_work, 2 * sizeof(unsigned)));
        CU(cudaMemsetAsync(c->d_work, 0, 2 * sizeof(unsigned), c->stream));
    }
    a.work = c->d_work;
    a.dbg = nullptr;
    if (getenv("B2A_STREAM_DEBUG"))
    {
        if (!c->d_dbg) CU(cudaMalloc(&c->d_dbg, sizeof(unsigned long long) * 8 * 4096));
        CU(cudaMemsetAsync(c->d_dbg, 0, sizeof(unsigned long long) * 8 * 4096, c->stream));
        a.dbg = c->d_dbg;
    }
    a.park_ns = 20000u;
    if (const char* e = getenv("B2A_PARK_NS")) a.park_ns = (unsigned)std::max(0, atoi(e));
    constexpr int NS = stream_stages(MODE, VAR);
    size_t smem = (size_t)NS * a.tile_floats * 4 + NS * (sizeof(FrameGeom) + 16 + sizeof(StageDesc)) + sizeof(double) * g.napm + extra_smem;
    if (MODE != 0) smem += sizeof(float) * ((g.napm + 1) & ~1) + sizeof(unsigned) * (a.smem_bins & 1) + sizeof(uint2) * kFastQueue + (MODE == 2 ? kCosLut + sizeof(double) * g.napm : 0);
    auto kern = k_stream<MODE, NAPM_T, VAR, FUSE, SPEC>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, M + 32, smem));
    if (per_sm < 1) return fail(B2A_ERR_UNSUPPORTED, "streaming kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    const int grid = std::min(a.nitems, per_sm * c->sm_count);
    a.cwait_ns = 0u;
    if (const char* e = getenv("B2A_CWAIT_NS")) a.cwait_ns = (unsigned)std::max(0, atoi(e));
    kern<<<grid, M + 32, smem, c->stream>>>(a);   // M consumer threads (one molecule each) + one producer warp
    CU(cudaGetLastError());
    return B2A_OK;
}

template <int MODE>
static int launch_stream(b2a_ctx* c, const Group& g, int com, int M, StreamArgs a, size_t extra_smem, int fast_mode = 0, int var = 0)
{
    // three-site molecules (water) get fully unrolled atom loops; everything else runs the generic instantiation
    if constexpr (MODE == 0)
    {
        // larger molecules: one lane per (molecule, dimension), ten molecules per warp, <= 24 KB of atoms per stage
        int kvar = g.napm >= 8 ? 2 : 0;   // measured cross-over (tools/bench_hbm_napm.py): 7 atoms 0.63 vs 0.57, 8 atoms 0.50 vs 0.58, 25 atoms 0.19 vs 0.47
        if (const char* e = getenv("B2A_K1_VAR")) kvar = atoi(e);
        const int warps = std::min(8, 24576 / (120 * g.napm));
        if (kvar == 2 && g.napm != 3 && warps >= 1) return launch_stream_t<0, 0, 2>(c, g, com, 32 * warps, a, extra_smem, fast_mode);
    }
    if constexpr (MODE == 2)
        if (var == 1)
        {
            // consumer threads; each evaluates two molecules, so a tile holds 2 * mc of them: at most 24 KB per stage
            int mc = std::min(M, std::max(32, (24576 / (24 * g.napm)) / 32 * 32));
            if (const char* e = getenv("B2A_K4_THREADS")) mc = std::max(32, std::min(256, atoi(e) / 32 * 32));
            if (g.napm == 3) return launch_stream_t<2, 3, 1>(c, g, com, mc, a, extra_smem, fast_mode);
            if (2 * mc * g.napm * 12 <= 32768) return launch_stream_t<2, 0, 1>(c, g, com, mc, a, extra_smem, fast_mode);
        }
    if (g.napm == 3) return launch_stream_t<MODE, 3>(c, g, com, M, a, extra_smem, fast_mode);
    return launch_stream_t<MODE, 0>(c, g, com, M, a, extra_smem, fast_mode);
}

// ---- K1 launch + cache -----------------------------------------------------------------------------------------
static int ensure_centers(b2a_ctx* c, Group& g, int com, bool need_fix)
{
    if (com < 0 || com > 3) return fail(B2A_ERR_ARG, "com(%d) must be 0 (mass), 1 (geometry), 2 (charge) or 3 (dipole)", com);
    if (c->nframes <= 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet");
    if (!g.d_cen[com]) CU(cudaMalloc(&g.d_cen[com], sizeof(double) * 3 * (size_t)g.nmol * c->maxK));
    if (need_fix && !g.d_fix[com]) CU(cudaMalloc(&g.d_fix[com], sizeof(int4) * (size_t)g.nmol * c->maxK));
    if (g.valid[com] == c->serial && (!need_fix || g.fix_valid[com] == c->serial)) return B2A_OK;
    // the int4 fixed-point copy (16 B per molecule on top of the 24 B of the doubles) is written only when a pair kernel will read it
    int4* const fix = need_fix ? g.d_fix[com] : nullptr;
    Timed tm(c, 0);
    if (g.csr)
    {
        dim3 grid((g.nmol + 127) / 128, c->nframes);
        k1_centers_csr<<<grid, 128, 0, c->stream>>>(c->d_x, (size_t)c->natoms * 3, c->nframes, g.d_index, g.d_off, g.nmol, g.d_wfull[com],
                                                    g.d_div[com], c->d_geom, g.d_cen[com], fix);
        CU(cudaGetLastError());
        g.valid[com] = c->serial;
        if (need_fix) g.fix_valid[com] = c->serial;
        return B2A_OK;
    }
    int   M = 0;
    if (stream_block(g, &M))
    {
        StreamArgs a{};
        a.cen = g.d_cen[com]; a.fix = fix;
        if (int rc = launch_stream<0>(c, g, com, M, a, 0)) return rc;
        g.valid[com] = c->serial;
        if (need_fix) g.fix_valid[com] = c->serial;
        return B2A_OK;
    }
    const int threads = 128;
    dim3      grid((g.nmol + threads - 1) / threads, c->nframes);
    k1_centers<<<grid, threads, sizeof(double) * g.napm, c->stream>>>(
        c->d_x, (size_t)c->natoms * 3, c->nframes, g.d_index, g.nmol, g.napm, g.d_w[com], make_div(g.wsum[com]), c->d_geom,
        g.d_cen[com], fix);
    CU(cudaGetLastError());
    g.valid[com] = c->serial;
    if (need_fix) g.fix_valid[com] = c->serial;
    return B2A_OK;
}

extern "C" int b2a_centers(b2a_ctx* c, int grp, int com, int frame, double* out)
{
    TEAM_CUR(c, b2a_centers(m, grp, com, frame, out));
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_centers: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, grp, &g)) return rc;
    if (c->nframes <= 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet");
    if (frame < 0 || frame >= c->nframes) return fail(B2A_ERR_ARG, "b2a_centers: frame %d outside batch of %d", frame, c->nframes);
    CU(cudaSetDevice(c->device));
    if (int rc = ensure_centers(c, *g, com, false)) return rc;
    CU(cudaMemcpyAsync(out, g->d_cen[com] + (size_t)frame * 3 * g->nmol, sizeof(double) * 3 * g->nmol, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

extern "C" int b2a_centers_batch(b2a_ctx* c, int grp, int com, double* out)
{
    TEAM_CUR(c, b2a_centers_batch(m, grp, com, out));
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_centers_batch: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, grp, &g)) return rc;
    CU(cudaSetDevice(c->device));
    if (int rc = ensure_centers(c, *g, com, false)) return rc;
    CU(cudaMemcpyAsync(out, g->d_cen[com], sizeof(double) * 3 * g->nmol * c->nframes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

extern "C" int b2a_centers_device(b2a_ctx* c, int grp, int com, const double** dptr)
{
    TEAM_CUR(c, b2a_centers_device(m, grp, com, dptr));
    if (!c || !dptr) return fail(B2A_ERR_ARG, "b2a_centers_device: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, grp, &g)) return rc;
    CU(cudaSetDevice(c->device));
    if (int rc = ensure_centers(c, *g, com, false)) return rc;   // queued on the context stream; nothing is copied or waited for
    *dptr = g->d_cen[com];
    return B2A_OK;
}

extern "C" int b2a_positions(b2a_ctx* c, int grp, int frame, double* out)
{
    TEAM_CUR(c, b2a_positions(m, grp, frame, out));
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_positions: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, grp, &g)) return rc;
    if (c->nframes <= 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet");
    if (frame < 0 || frame >= c->nframes) return fail(B2A_ERR_ARG, "b2a_positions: frame %d outside batch of %d", frame, c->nframes);
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)g->nmol * (g->csr ? g->max_napm : g->napm);
    if (c->pos_scratch_n < n)
    {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_pos_scratch);
        c->d_pos_scratch = nullptr;
        CU(cudaMalloc(&c->d_pos_scratch, sizeof(double) * 3 * n));
        c->pos_scratch_n = n;
    }
    const int threads = 256;
    if (g->csr)
    {
        CU(cudaMemsetAsync(c->d_pos_scratch, 0, sizeof(double) * 3 * n, c->stream));   // slots j >= napm[i] read back as 0
        k1b_positions_csr<<<(unsigned)((g->ngx + threads - 1) / threads), threads, 0, c->stream>>>(
            c->d_x + (size_t)frame * c->natoms * 3, g->d_index, g->d_off, g->d_mol_of, g->ngx, g->nmol, c->h_geom[frame], c->d_pos_scratch);
    }
    else
    k1b_positions<<<(unsigned)((n + threads - 1) / threads), threads, 0, c->stream>>>(
        c->d_x + (size_t)frame * c->natoms * 3, g->d_index, g->nmol, g->napm, c->h_geom[frame], c->d_pos_scratch);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_pos_scratch, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

// ---- velocities / forces (reference ipp:201-339): no unwrapping ------------------------------------------------------
extern "C" int b2a_submit_field(b2a_ctx* c, int field, int nframes, const float* data)
{
    TEAM_CUR(c, b2a_submit_field(m, field, nframes, data));
    if (!c || !data) return fail(B2A_ERR_ARG, "b2a_submit_field: null argument");
    if (field != B2A_FIELD_V && field != B2A_FIELD_F) return fail(B2A_ERR_ARG, "b2a_submit_field: field must be B2A_FIELD_V or B2A_FIELD_F");
    if (c->nframes <= 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet");
    if (nframes != c->nframes) return fail(B2A_ERR_ARG, "b2a_submit_field: %d frames, but the current batch holds %d", nframes, c->nframes);
    CU(cudaSetDevice(c->device));
    if (!c->d_field[field]) CU(cudaMalloc(&c->d_field[field], (size_t)c->natoms * 3 * sizeof(float) * c->maxK));
    CU(cudaMemcpyAsync(c->d_field[field], data, (size_t)c->natoms * 3 * sizeof(float) * nframes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev_field, c->stream));   // the caller's buffer is free again once this has passed (b2a_submit_batch waits)
    c->field_pending = true;
    c->field_serial[field] = c->serial;
    return B2A_OK;
}

static int field_ready(b2a_ctx* c, int field, Group** g, int grp)
{
    if (field != B2A_FIELD_V && field != B2A_FIELD_F) return fail(B2A_ERR_ARG, "field must be B2A_FIELD_V or B2A_FIELD_F");
    if (int rc = get_group(c, grp, g)) return rc;
    if (c->nframes <= 0 || !c->d_field[field] || c->field_serial[field] != c->serial)
        return fail(B2A_ERR_STATE, "no %s submitted for the current batch (b2a_submit_field)", field == B2A_FIELD_V ? "velocities" : "forces");
    return B2A_OK;
}

static int run_field_centers(b2a_ctx* c, int field, Group& g, int com)
{
    if (com < 0 || com > 2) return fail(B2A_ERR_ARG, "com(%d) must be 0 (mass), 1 (geometry) or 2 (charge)", com);
    const size_t n = (size_t)3 * g.nmol * c->nframes;
    if (c->field_cen_n < n)
    {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_field_cen); c->d_field_cen = nullptr;
        CU(cudaMalloc(&c->d_field_cen, sizeof(double) * 3 * (size_t)g.nmol * c->maxK));
        c->field_cen_n = (size_t)3 * g.nmol * c->maxK;
    }
    Timed tm(c, 0);
    dim3  grid((g.nmol + 127) / 128, c->nframes);
    if (g.csr)
    {
        k1_field_centers_csr<<<grid, 128, 0, c->stream>>>(c->d_field[field], (size_t)c->natoms * 3, c->nframes, g.d_index, g.d_off, g.nmol,
                                                          g.d_wfull[com], g.d_div[com], c->d_field_cen);
        CU(cudaGetLastError());
        return B2A_OK;
    }
    k1_field_centers<<<grid, 128, sizeof(double) * g.napm, c->stream>>>(c->d_field[field], (size_t)c->natoms * 3, c->nframes, g.d_index, g.nmol,
                                                                         g.napm, g.d_w[com], make_div(g.wsum[com]), c->d_field_cen);
    CU(cudaGetLastError());
    return B2A_OK;
}

extern "C" int b2a_field_centers_batch(b2a_ctx* c, int field, int grp, int com, double* out)
{
    TEAM_CUR(c, b2a_field_centers_batch(m, field, grp, com, out));
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_field_centers_batch: null argument");
    Group* g = nullptr;
    if (int rc = field_ready(c, field, &g, grp)) return rc;
    CU(cudaSetDevice(c->device));
    if (int rc = run_field_centers(c, field, *g, com)) return rc;
    CU(cudaMemcpyAsync(out, c->d_field_cen, sizeof(double) * 3 * g->nmol * c->nframes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

extern "C" int b2a_field_centers(b2a_ctx* c, int field, int grp, int com, int frame, double* out)
{
    TEAM_CUR(c, b2a_field_centers(m, field, grp, com, frame, out));
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_field_centers: null argument");
    Group* g = nullptr;
    if (int rc = field_ready(c, field, &g, grp)) return rc;
    if (frame < 0 || frame >= c->nframes) return fail(B2A_ERR_ARG, "b2a_field_centers: frame %d outside batch of %d", frame, c->nframes);
    CU(cudaSetDevice(c->device));
    if (int rc = run_field_centers(c, field, *g, com)) return rc;
    CU(cudaMemcpyAsync(out, c->d_field_cen + (size_t)frame * 3 * g->nmol, sizeof(double) * 3 * g->nmol, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

extern "C" int b2a_field_values(b2a_ctx* c, int field, int grp, int frame, double* out)
{
    TEAM_CUR(c, b2a_field_values(m, field, grp, frame, out));
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_field_values: null argument");
    Group* g = nullptr;
    if (int rc = field_ready(c, field, &g, grp)) return rc;
    if (frame < 0 || frame >= c->nframes) return fail(B2A_ERR_ARG, "b2a_field_values: frame %d outside batch of %d", frame, c->nframes);
    if (g->csr) return fail(B2A_ERR_UNSUPPORTED, "b2a_field_values on a heterogeneous (GmxHandleFull) group: the reference class has no loadVelocity / loadForce");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)g->nmol * g->napm;
    if (c->pos_scratch_n < n)
    {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_pos_scratch);
        c->d_pos_scratch = nullptr;
        CU(cudaMalloc(&c->d_pos_scratch, sizeof(double) * 3 * n));
        c->pos_scratch_n = n;
    }
    k1b_field_values<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_field[field] + (size_t)frame * c->natoms * 3, g->d_index, g->nmol,
                                                                         g->napm, c->d_pos_scratch);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_pos_scratch, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

// ---- accumulators ----------------------------------------------------------------------------------------------
static int new_acc(b2a_ctx* c, std::unique_ptr<Acc> a, int* id)
{
    CU(cudaMalloc(&a->d_hist, sizeof(unsigned long long) * (a->ndev + B2A_STAT_NSLOTS)));
    CU(cudaMemsetAsync(a->d_hist, 0, sizeof(unsigned long long) * (a->ndev + B2A_STAT_NSLOTS), c->stream));
    c->accs.push_back(std::move(a));
    *id = (int)c->accs.size() - 1;
    return B2A_OK;
}

// next pinned slot of the per-frame constants (slot index through *slot); waits only if the host is kFrameRing batches ahead
static int ring_acquire(Acc& a, int* slot)
{
    FrameRing& r = a.ring;
    const int  s = r.next;
    r.next       = (s + 1) % kFrameRing;
    if (!r.ev[s]) CU(cudaEventCreateWithFlags(&r.ev[s], cudaEventDisableTiming));
    else CU(cudaEventSynchronize(r.ev[s]));
    *slot = s;
    return B2A_OK;
}

// Scratch sized at create time must still fit the groups (b2a_set_group may have redefined them since): grow it here
// rather than let the prep kernels write past the old allocation.
template <typename T>
static int regrow(T** p, size_t n)
{
    cudaFree(*p);
    *p = nullptr;
    CU(cudaMalloc(p, sizeof(T) * n));
    return B2A_OK;
}

template <typename T>
static int grow_log(b2a_ctx* c, T** p, size_t used, size_t n)
{
    T* q = nullptr;
    CU(cudaMalloc(&q, sizeof(T) * n));
    if (*p && used) CU(cudaMemcpyAsync(q, *p, sizeof(T) * used, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(*p);
    *p = q;
    return B2A_OK;
}

// room for `more` frames at the end of an accumulator's per-frame log; grows geometrically (rare: a stream-ordered copy)
static int series_reserve(b2a_ctx* c, Acc& a, size_t more)
{
    const size_t need = a.series_len + more;
    if (need <= a.series_cap) return B2A_OK;
    const size_t cap = std::max(need, std::max<size_t>(2 * a.series_cap, 4 * (size_t)c->maxK));
    if (a.kind == ACC_REGION)
    {
        if (int rc = grow_log(c, &a.d_perframe, a.series_len * a.ndev, cap * a.ndev)) return rc;
    }
    else
    {
        if (int rc = grow_log(c, &a.d_table, a.series_len * (size_t)a.coord.max_count, cap * (size_t)a.coord.max_count)) return rc;
        if (int rc = grow_log(c, &a.d_series_numA, a.series_len, cap)) return rc;
    }
    a.series_cap = cap;
    return B2A_OK;
}

static int get_acc(b2a_ctx* c, int id, Acc** out)
{
    if (!c) return fail(B2A_ERR_ARG, "null context");
    if (id < 0 || id >= (int)c->accs.size() || !c->accs[id]) return fail(B2A_ERR_ARG, "unknown accumulator %d", id);
    *out = c->accs[id].get();
    return B2A_OK;
}

extern "C" int b2a_density_create(b2a_ctx* c, const b2a_density_spec* s, int* id)
{
    TEAM_ALL(c, b2a_density_create(m, s, id));   // created in the same order on every member: same id everywhere
    if (!c || !s || !id) return fail(B2A_ERR_ARG, "b2a_density_create: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, s->grp, &g)) return rc;
    if (s->nbin <= 0 || s->dim < 0 || s->dim > 2 || s->dim2 > 2 || !(s->dbin > 0)) return fail(B2A_ERR_ARG, "b2a_density_create: bad spec");
    if (s->com < 0 || s->com > 3) return fail(B2A_ERR_ARG, "com(%d) out of range", s->com);
    CU(cudaSetDevice(c->device));
    std::unique_ptr<Acc> a(new Acc);
    a->kind = ACC_DENSITY;
    a->den  = *s;
    a->ncounters = a->ndev = (size_t)s->nbin;
    return new_acc(c, std::move(a), id);
}

extern "C" int b2a_region_create(b2a_ctx* c, const b2a_region_spec* s, int* id)
{
    TEAM_ALL(c, b2a_region_create(m, s, id));   // created in the same order on every member: same id everywhere
    if (!c || !s || !id) return fail(B2A_ERR_ARG, "b2a_region_create: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, s->grp, &g)) return rc;
    if (s->com < 0 || s->com > 3) return fail(B2A_ERR_ARG, "com(%d) out of range", s->com);
    if (s->nfilter < 0 || s->nfilter > 3 || s->nbdim < 0 || s->nbdim > 2) return fail(B2A_ERR_ARG, "b2a_region_create: 0..3 filters and 0..2 binned dimensions");
    for (int q = 0; q < s->nfilter; ++q)
        if (s->fdim[q] < 0 || s->fdim[q] > 2) return fail(B2A_ERR_ARG, "b2a_region_create: filter dimension out of range");
    size_t n = 1;
    for (int b = 0; b < s->nbdim; ++b)
    {
        if (s->bdim[b] < 0 || s->bdim[b] > 3 || s->nb[b] <= 0 || !(s->dbin[b] > 0)) return fail(B2A_ERR_ARG, "b2a_region_create: bad binned dimension %d", b);
        n *= (size_t)s->nb[b];
    }
    if (n > ((size_t)1 << 28)) return fail(B2A_ERR_UNSUPPORTED, "b2a_region_create: %zu counters", n);
    CU(cudaSetDevice(c->device));
    std::unique_ptr<Acc> a(new Acc);
    a->kind   = ACC_REGION;
    a->region = *s;
    a->ncounters = a->ndev = n;
    if (s->per_frame)
    {
        if (n * (size_t)c->maxK > ((size_t)1 << 28)) return fail(B2A_ERR_UNSUPPORTED, "b2a_region_create: per-frame table of %zu x %d counters", n, c->maxK);
        if (int rc = series_reserve(c, *a, (size_t)c->maxK)) return rc;
    }
    return new_acc(c, std::move(a), id);
}

extern "C" int b2a_rdf_create(b2a_ctx* c, const b2a_rdf_spec* s, int* id)
{
    TEAM_ALL(c, b2a_rdf_create(m, s, id));   // created in the same order on every member: same id everywhere
    if (!c || !s || !id) return fail(B2A_ERR_ARG, "b2a_rdf_create: null argument");
    Group *g0 = nullptr, *g1 = nullptr;
    if (int rc = get_group(c, s->grp0, &g0)) return rc;
    if (int rc = get_group(c, s->grp1, &g1)) return rc;
    if (s->nbin <= 0 || s->Rbin <= 0 || !(s->Rm > 0.f) || s->dim < 0 || s->dim > 2 || s->dim2 < 0 || s->dim2 > 2)
        return fail(B2A_ERR_ARG, "b2a_rdf_create: bad spec (nbin=%d Rbin=%d Rm=%g dim=%d dim2=%d)", s->nbin, s->Rbin, (double)s->Rm, s->dim, s->dim2);
    if (s->com0 < 0 || s->com0 > 3 || s->com1 < 0 || s->com1 > 3) return fail(B2A_ERR_ARG, "com out of range");
    if (s->nbin > 4096) return fail(B2A_ERR_UNSUPPORTED, "b2a_rdf_create: nbin > 4096 slabs not supported");
    CU(cudaSetDevice(c->device));
    std::unique_ptr<Acc> a(new Acc);
    a->kind      = ACC_RDF;
    a->rdf       = *s;
    a->ncounters = (size_t)s->nbin * s->Rbin;
    a->ndev      = (size_t)(s->nbin + 1) * s->Rbin;   // + slab nbin: i whose meZ == nbin (out of bounds in the reference)
    const int nslab = s->nbin + 1;
    CU(cudaMalloc(&a->d_slab_id, sizeof(int) * (size_t)g0->nmol * c->maxK));
    CU(cudaMalloc(&a->d_slab_cnt, sizeof(unsigned) * 2 * (size_t)nslab * c->maxK));
    CU(cudaMalloc(&a->d_sortedA, sizeof(int4) * (size_t)g0->nmol * c->maxK));
    CU(cudaMalloc(&a->d_frames, sizeof(RdfFrame) * c->maxK));
    CU(cudaMallocHost(&a->h_frames, sizeof(RdfFrame) * c->maxK * kFrameRing));
    a->alloc_n0 = g0->nmol; a->alloc_n1 = g1->nmol;
    return new_acc(c, std::move(a), id);
}

extern "C" int b2a_rdf_tune(b2a_ctx* c, int id, int symmetric, int paranoid)
{
    TEAM_ALL(c, b2a_rdf_tune(m, id, symmetric, paranoid));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind != ACC_RDF) return fail(B2A_ERR_ARG, "accumulator %d is not an RDF", id);
    a->symmetric = symmetric;
    a->paranoid  = paranoid;
    return B2A_OK;
}

extern "C" int b2a_acc_fast(b2a_ctx* c, int id, int mode)
{
    TEAM_ALL(c, b2a_acc_fast(m, id, mode));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind == ACC_RDF) return fail(B2A_ERR_ARG, "accumulator %d is an RDF: use b2a_rdf_tune", id);
    if (mode < 0 || mode > 2) return fail(B2A_ERR_ARG, "b2a_acc_fast: mode must be 0 (FP64 only), 1 (FP32 bracket + FP64 queue) or 2 (validate)");
    a->fast = mode;
    return B2A_OK;
}

extern "C" int b2a_acc_also_centers(b2a_ctx* c, int id, int com)
{
    TEAM_ALL(c, b2a_acc_also_centers(m, id, com));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind != ACC_ORIENT) return fail(B2A_ERR_UNSUPPORTED, "b2a_acc_also_centers: accumulator %d is not an orientation accumulator", id);
    if (com < -1 || com > 3) return fail(B2A_ERR_ARG, "com(%d) must be -1 (off), 0 (mass), 1 (geometry), 2 (charge) or 3 (dipole)", com);
    a->also_com = com;
    return B2A_OK;
}

extern "C" int b2a_rdf_cells(b2a_ctx* c, int id, int mode)
{
    TEAM_ALL(c, b2a_rdf_cells(m, id, mode));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind != ACC_RDF && a->kind != ACC_COORD) return fail(B2A_ERR_ARG, "accumulator %d is not an RDF or coordination-number accumulator", id);
    if (mode < -1 || mode > 1) return fail(B2A_ERR_ARG, "b2a_rdf_cells: mode must be -1 (auto), 0 (never) or 1 (whenever cells cull)");
    a->cells = mode;
    return B2A_OK;
}

static float next_down(float v) { return std::nextafterf(v, -INFINITY); }
static float next_up(float v) { return std::nextafterf(v, INFINITY); }

template <int IPT, int THREADS, bool CUBIC, bool PARANOID, int MODE, bool DBUF = false>
static int launch_rdf(b2a_ctx* c, Acc& a, Group& g0, Group& g1, const RdfDev& p, size_t smem, dim3 grid)
{
    auto kern = k3_rdf<IPT, THREADS, true, CUBIC, PARANOID, MODE, DBUF>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, THREADS, smem, a.last_stream>>>(a.sortedA_view ? a.sortedA_view : a.d_sortedA, g1.d_fix[a.rdf.com1], g0.d_cen[a.rdf.com0], g1.d_cen[a.rdf.com1],
                                            a.d_slab_cnt, a.d_frames, p, a.d_hist, a.d_hist + a.ndev);
    CU(cudaGetLastError());
    return B2A_OK;
}

// cell mode, warp-autonomous tiles (b2a_kernels_rdf_cell.cuh)
template <bool CUBIC, bool PARANOID>
static int launch_rdf_cellw(b2a_ctx* c, Acc& a, Group& g0, Group& g1, const RdfDev& p, const CellDev& cd, size_t smem, dim3 grid)
{
    auto kern = k3_rdf_cellw<kCwWarps, CUBIC, PARANOID>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 32 * kCwWarps, smem, a.last_stream>>>(a.sortedA_view ? a.sortedA_view : a.d_sortedA, g0.d_cen[a.rdf.com0], g1.d_cen[a.rdf.com1], a.d_slab_cnt,
                                                       a.d_frames, p, cd, a.d_hist, a.d_hist + a.ndev);
    CU(cudaGetLastError());
    return B2A_OK;
}

template <int IPT, int MODE, int TH = 128>
static int dispatch_rdf(b2a_ctx* c, Acc& a, Group& g0, Group& g1, const RdfDev& p, size_t smem, dim3 grid, bool cubic, bool dbuf = false)
{
    if (dbuf && !a.paranoid)   // double-buffered j-tile (8 KB more shared memory: the caller checked that residency is unchanged)
        return cubic ? launch_rdf<IPT, TH, true, false, MODE, true>(c, a, g0, g1, p, smem + kTileJ * sizeof(int4), grid)
                     : launch_rdf<IPT, TH, false, false, MODE, true>(c, a, g0, g1, p, smem + kTileJ * sizeof(int4), grid);
    if (a.paranoid)
        return cubic ? launch_rdf<IPT, TH, true, true, MODE>(c, a, g0, g1, p, smem, grid) : launch_rdf<IPT, TH, false, true, MODE>(c, a, g0, g1, p, smem, grid);
    return cubic ? launch_rdf<IPT, TH, true, false, MODE>(c, a, g0, g1, p, smem, grid) : launch_rdf<IPT, TH, false, false, MODE>(c, a, g0, g1, p, smem, grid);
}

constexpr int kCellIPT = 4, kCellThreads = 64;   // k5_coord cell tiles (64 threads x 4); the RDF's cell tiles are kCwTile = 128 molecules per warp

// Cell grid for r_max << L: cells per dimension 2^bits, reach per dimension from the cut-off plus the fixed-point slop.
// Returns false when cells would not cull enough to beat the brute-force (possibly symmetric) evaluation.
// `columns`: the caller cuts its i-tiles from whole (cx, cy) columns of cells (k3_rdf_cellw) instead of giving every cell a
// tile of its own (k5_coord): tiles are always full, so cells may be much finer than a tile, and a tile that spans several
// cells along z searches the union of their z ranges.
static bool choose_cells(b2a_ctx* c, const Acc& a, const RdfDev& p, bool sym_possible, int* bits, int reach[3], int ti = kCellIPT * kCellThreads,
                         double* cost_out = nullptr, bool columns = false)
{
    int want = a.cells;
    if (const char* e = getenv("B2A_RDF_CELLS")) want = atoi(e);
    if (want == 0) return false;
    const int    TI    = ti;
    const double nmin  = (double)std::min(p.n0, p.n1);
    double       bestc = 1e30;
    int          bestb = -1, bestr[3] = { 0, 0, 0 };
    int force_bits = 0;   // experiment switch: cells per dimension = 2^B2A_RDF_CELL_BITS for every column-tiled accumulator
    if (const char* e = getenv("B2A_RDF_CELL_BITS")) force_bits = columns ? atoi(e) : 0;
    for (int b = 2; b <= 7; ++b)
    {
        const int    nc = 1 << b;
        const double m  = nmin / ((double)nc * nc * nc);        // mean molecules per cell
        if (force_bits && b != force_bits) continue;
        if (!force_bits && m < (columns ? 0.2 : (ti >= 512 ? 48.0 : 12.0))) break;
        if ((unsigned long long)p.nslab << (3 * b) > (4ull << 20)) break;
        int  r[3];
        bool ok = false;
        for (int k = 0; k < 3; ++k)
        {
            int rk = 0;
            for (int f = 0; f < c->nframes; ++f)
            {
                const double L = c->h_geom[f].L[k], cell = L / nc, unit = L / 4294967296.0;
                // |true dx| < Rm(1+3e-8) can still be counted (float-rounded cut); fixed coordinates are within half a
                // unit of the truth, and flooring two positions to cells can add one more cell of distance
                rk = std::max(rk, (int)std::floor((p.Rm * (1.0 + 1e-6) + 2.0 * unit) / cell) + 1);
            }
            r[k] = rk;
            if (2 * rk + 1 < nc) ok = true;
        }
        if (!ok) continue;
        double frac = 1.0, util, cost;
        if (!columns)
        {
            for (int k = 0; k < 3; ++k) frac *= (double)std::min(2 * r[k] + 1, nc) / nc;
            util = m / (std::ceil(m / TI) * TI);       // lane utilisation of a one-cell tile
            cost = frac / std::min(1.0, util) * 1.15;  // 15 %: IPT 4 instead of 8, per-range tile edges
        }
        else
        {
            // a tile of TI molecules spans TI / m cells of its column (+1 for the alignment) and searches the union of their z
            // ranges; every neighbour run costs about four j-iterations of set-up (offsets, staging the first chunk) on top, and the
            // last tile of a column is evaluated on half the register slots when it holds at most TI / 2 molecules.  Calibrated on
            // config 5 (k3_rdf_cellw, 1.2 M waters x 8 k ions: 8^3 cells 0.44 ms, 16^3 0.32, 32^3 0.33; 8 k x 8 k ions: all pairs
            // 0.32 ms, 16^3 or 32^3 cells 0.11): sparse groups want cells far finer than one molecule per cell, and then share the
            // sorted view of the big group with its own pair.
            const double p0    = (double)p.n0 / ((double)nc * nc * nc);
            const double zc    = std::min((double)nc, 2.0 * r[2] + 1.0 + (double)TI / std::max(p0, 1e-9));
            const double col   = p0 * nc;                                  // molecules per i column
            util               = col / (std::ceil(col / (TI / 2)) * (TI / 2));
            const double jrun  = zc * (double)p.n1 / ((double)nc * nc * nc);   // j molecules per neighbour column
            // neighbour columns the z reach table keeps (corners outside the cut-off sphere are dropped; first frame's box)
            double       ncols = 0.0;
            const int    wx = std::min(2 * r[0] + 1, nc), wy = std::min(2 * r[1] + 1, nc);
            const bool   sphere = wx < nc && wy < nc && r[0] <= 7 && r[1] <= 7;
            const double cwx = c->h_geom[0].L[0] / nc, cwy = c->h_geom[0].L[1] / nc;
            for (int ix = 0; ix < wx; ++ix)
                for (int iy = 0; iy < wy; ++iy)
                {
                    const double mx = std::max(std::abs(ix - wx / 2) - 1, 0) * cwx, my = std::max(std::abs(iy - wy / 2) - 1, 0) * cwy;
                    if (!sphere || mx * mx + my * my <= p.Rm * p.Rm) ncols += 1.0;
                }
            frac               = ncols * zc / ((double)nc * nc * nc);
            cost               = frac / std::min(1.0, util) * (1.0 + 4.0 / std::max(jrun, 1e-3)) * 1.1;
            // one block per tile: a small i side leaves most of the machine idle, where the all-pairs kernel would split j
            const double blocks = std::ceil(col / TI) * (double)nc * nc * c->nframes;
            cost /= std::min(1.0, blocks / (4.0 * c->sm_count));
            if (sym_possible) cost *= 0.5;   // k3_rdf_cellw evaluates eligible frames symmetrically as well
        }
        if (cost < bestc) { bestc = cost; bestb = b; bestr[0] = r[0]; bestr[1] = r[1]; bestr[2] = r[2]; }
    }
    if (bestb < 0) return false;
    if (cost_out) *cost_out = bestc;
    const double brute = sym_possible ? 0.5 : 1.0;
    if (want < 0 && bestc > 0.8 * brute) return false;
    *bits = bestb;
    reach[0] = bestr[0]; reach[1] = bestr[1]; reach[2] = bestr[2];
    return true;
}

// Sorted-by-cell view of one group for the current batch, built on first use and shared afterwards (see CellSort).  `slab` is
// the accumulator whose slab ids enter the key (its k3_slab_count has been queued), or nullptr for a plain cell sort.
static int cell_sort(b2a_ctx* c, int grp, Group& g, int com, int bits, const Acc* slab, const RdfDev& p, CellSort** out, cudaStream_t stm)
{
    const int      n     = g.nmol;
    const unsigned nkeys = (slab ? (unsigned)p.nslab : 1u) << (3 * bits);
    CellSort*      cs    = nullptr;
    for (auto& e : c->cell_sorts)
    {
        if (e->grp != grp || e->com != com || e->bits != bits || e->with_slab != (slab ? 1 : 0)) continue;
        if (slab && !(e->nbin == p.nbin && e->dim == p.dim && e->dim2 == p.dim2 && e->low == p.low && e->up == p.up && e->low2 == p.low2 && e->up2 == p.up2))
            continue;
        cs = e.get();
        break;
    }
    if (!cs)
    {
        c->cell_sorts.emplace_back(new CellSort);
        cs = c->cell_sorts.back().get();
        cs->grp = grp; cs->com = com; cs->bits = bits; cs->with_slab = slab ? 1 : 0;
        if (slab) { cs->nbin = p.nbin; cs->dim = p.dim; cs->dim2 = p.dim2; cs->low = p.low; cs->up = p.up; cs->low2 = p.low2; cs->up2 = p.up2; }
    }
    *out = cs;
    if ((size_t)n > cs->cap_n || nkeys > cs->cap_keys)
    {   // about to replace arrays another stream may still read
        if (int rc = join_side(c)) return rc;
        CU(cudaStreamSynchronize(c->stream));
    }
    if ((size_t)n > cs->cap_n)
    {
        cudaFree(cs->keys); cudaFree(cs->sorted); cs->keys = nullptr; cs->sorted = nullptr; cs->cap_n = 0; cs->serial = 0;
        CU(cudaMalloc(&cs->keys, sizeof(int) * (size_t)n * c->maxK));
        CU(cudaMalloc(&cs->sorted, sizeof(int4) * (size_t)n * c->maxK));
        cs->cap_n = (size_t)n;
    }
    if (nkeys > cs->cap_keys)
    {
        cudaFree(cs->cells); cs->cells = nullptr; cs->cap_keys = 0; cs->serial = 0;
        CU(cudaMalloc(&cs->cells, sizeof(unsigned) * 3 * (size_t)(nkeys + 1) * c->maxK));
        cs->cap_keys = nkeys;
    }
    if (cs->serial == c->serial && cs->nkeys == nkeys)
    {   // built for this batch already -- possibly by another accumulator on another stream
        if (cs->built_on != stm && cs->ready) CU(cudaStreamWaitEvent(stm, cs->ready, 0));
        return B2A_OK;
    }
    cs->nkeys = nkeys;
    const size_t sz = (size_t)(nkeys + 1) * c->maxK;
    unsigned *cnt = cs->cells, *st = cs->cells + sz, *cur = cs->cells + 2 * sz;
    CU(cudaMemsetAsync(cs->cells, 0, sizeof(unsigned) * 3 * sz, stm));
    dim3 grid((n + 255) / 256, c->nframes);
    k3_cell_keys<<<grid, 256, 0, stm>>>(g.d_fix[com], slab ? slab->d_slab_id : nullptr, c->nframes, n, bits, nkeys, cs->keys, cnt);
    k3_cell_scan<<<c->nframes, 1024, 0, stm>>>(cnt, nkeys, st);
    k3_cell_scatter<<<grid, 256, 0, stm>>>(g.d_fix[com], cs->keys, c->nframes, n, nkeys, st, cur, cs->sorted);
    CU(cudaGetLastError());
    if (!cs->ready) CU(cudaEventCreateWithFlags(&cs->ready, cudaEventDisableTiming));
    CU(cudaEventRecord(cs->ready, stm));
    cs->built_on = stm;
    cs->serial = c->serial;
    return B2A_OK;
}

static int accumulate_rdf(b2a_ctx* c, Acc& a)
{
    const b2a_rdf_spec& s = a.rdf;
    Group *g0 = nullptr, *g1 = nullptr;
    if (int rc = get_group(c, s.grp0, &g0)) return rc;
    if (int rc = get_group(c, s.grp1, &g1)) return rc;
    if (int rc = ensure_centers(c, *g0, s.com0, true)) return rc;
    if (int rc = ensure_centers(c, *g1, s.com1, true)) return rc;
    // stream of this accumulation: the context stream, or -- in a context with several RDF accumulators -- the accumulator's own,
    // forked here (the centres are queued) and joined by join_side()
    cudaStream_t stm = c->stream;
    {
        int nrdf = 0;
        for (auto& o : c->accs) if (o && o->kind == ACC_RDF) ++nrdf;
        bool side = nrdf > 1;
        if (const char* e = getenv("B2A_SIDE_STREAMS")) side = side && atoi(e) != 0;
        if (side)
        {
            if (!a.side) CU(cudaStreamCreateWithFlags(&a.side, cudaStreamNonBlocking));
            if (!a.ev_done) CU(cudaEventCreateWithFlags(&a.ev_done, cudaEventDisableTiming));
            CU(cudaEventRecord(c->ev_fork, c->stream));
            CU(cudaStreamWaitEvent(a.side, c->ev_fork, 0));
            stm = a.side;
        }
    }
    a.last_stream = stm;
    struct Done   // whatever happens below, what was queued on the side stream is joined later
    {
        Acc& a; cudaStream_t stm; b2a_ctx* c;
        ~Done() { if (stm != c->stream) { cudaEventRecord(a.ev_done, stm); a.side_pending = true; } }
    };
    Done done{ a, stm, c };

    RdfDev p{};
    p.n0 = g0->nmol; p.n1 = g1->nmol; p.nbin = s.nbin; p.nslab = s.nbin + 1; p.Rbin = s.Rbin;
    p.dim = s.dim; p.dim2 = s.dim2;
    p.low = (double)s.low; p.up = (double)s.up; p.low2 = (double)s.low2; p.up2 = (double)s.up2;
    p.width  = (double)(float)(s.up - s.low);          // float subtraction, calc_rdf_region.cpp:77
    p.rm2    = (double)(float)(s.Rm * s.Rm);           // float product, :88
    p.Rm     = (double)s.Rm;
    p.Rbin_d = (double)s.Rbin;
    p.rbin_bits = kMagicBits + (unsigned)s.Rbin;
    if (g0->nmol > a.alloc_n0)
    {   // the group was redefined with more molecules after b2a_rdf_create: the per-molecule scratch must follow
        if (a.side) CU(cudaStreamSynchronize(a.side));
        CU(cudaStreamSynchronize(c->stream));
        if (int rc = regrow(&a.d_slab_id, (size_t)g0->nmol * c->maxK)) return rc;
        if (int rc = regrow(&a.d_sortedA, (size_t)g0->nmol * c->maxK)) return rc;
        a.alloc_n0 = g0->nmol;
    }
    // symmetric evaluation is legal only when both sides are the very same molecules with the same centres: same atoms, same
    // centre kind, same weights and the same divisor (b2a_set_group_wsum may have overridden one of them)
    const bool same = s.com0 == s.com1 && (s.grp0 == s.grp1 || (g0->nmol == g1->nmol && g0->napm == g1->napm && !g0->csr && !g1->csr &&
                                                                g0->h_index == g1->h_index && g0->h_w[s.com0] == g1->h_w[s.com1] &&
                                                                g0->wsum[s.com0] == g1->wsum[s.com1]));
    p.sym_enabled   = (a.symmetric && same) ? 1 : 0;

    // per-frame fast-path constants
    bool cubic = true;
    int  slot  = 0;
    if (int rc = ring_acquire(a, &slot)) return rc;
    RdfFrame* const h_frames = a.h_frames + (size_t)slot * c->maxK;
    for (int f = 0; f < c->nframes; ++f)
    {
        const FrameGeom& g = c->h_geom[f];
        if (g.L[0] != g.L[1] || g.L[0] != g.L[2]) cubic = false;
    }
    // K3 error bound: relative error of the FP32 distance <= 4.5 * 2^-24 (cubic), 6.5 * 2^-24 otherwise
    const double eps_true = (cubic ? 4.5 : 6.5) * std::ldexp(1.0, -24);
    // slack above the true error bound: it only has to absorb the ABSOLUTE grid error (delta_bins below) from bin LOWK upwards, and
    // LOWK is derived from it -- a smaller slack means fewer undecided pairs (their share is ~2 eps t) and a few more always-rechecked
    // low bins (r < ~0.2 nm: no pairs in a liquid).  Round 1 used 1.3e-7 (LOWK ~ 22); 0.4e-7 (LOWK ~ 70-115) re-evaluates 22 % fewer pairs.
    const double eps      = eps_true + 0.4e-7;
    for (int f = 0; f < c->nframes; ++f)
    {
        const FrameGeom& g    = c->h_geom[f];
        RdfFrame&        fr   = h_frames[f];
        const double     unit = g.L[0] / 4294967296.0;                  // nm per x fixed unit
        const double     sc   = unit * (double)s.Rbin / (double)s.Rm;   // bins per unit
        float lo = (float)(sc * (1.0 - eps)), hi = (float)(sc * (1.0 + eps));
        if ((double)lo > sc * (1.0 - eps)) lo = next_down(lo);
        if ((double)hi < sc * (1.0 + eps)) hi = next_up(hi);
        fr.s_lo = lo; fr.s_hi = hi;
        fr.ay = (float)(g.L[1] / g.L[0]);
        fr.az = (float)(g.L[2] / g.L[0]);
        // absolute error: coordinates are rounded to the fixed grid (<= 1 unit per component of the difference)
        const double delta_bins = std::sqrt(3.0) * 1.0001 * sc * std::max(1.0, std::max((double)fr.ay, (double)fr.az));
        const unsigned lowk     = (unsigned)std::ceil(delta_bins / (eps - eps_true)) + 2;
        fr.lowk_bits = kMagicBits + lowk;
        // r2max: r with bin coordinate Rbin + 0.5 -- both bracket bins are exactly Rbin for any eps < 1e-4
        const double r_units = ((double)s.Rbin + 0.5) / sc;
        fr.r2max = (float)(r_units * r_units);
        for (int k = 0; k < 3; ++k) fr.L[k] = g.L[k];
    }
    static_assert(sizeof(RdfFrame) % 4 == 0, "RdfFrame is fetched word by word");
    k_fetch_words<<<1, 256, 0, stm>>>(reinterpret_cast<unsigned*>(a.d_frames), reinterpret_cast<const unsigned*>(h_frames),
                                            (int)(sizeof(RdfFrame) / 4) * c->nframes);
    CU(cudaGetLastError());
    CU(cudaEventRecord(a.ring.ev[slot], stm));

    // Shared histogram: pairs beyond the cut-off are clamped into ONE trash bin (bin Rbin).  Measured on B200
    // (profiles/r01_k3_tuning.md): faster than an oversize histogram -- the ~48 % of lanes beyond Rm hit a single address
    // (one merged wavefront instead of random bank conflicts) and never enter the undecided queue.
    const size_t fixed_smem = kTileJ * sizeof(int4) + kQueueCap * sizeof(unsigned) + 16;
    p.kstride               = (unsigned)s.Rbin + 3;
    const size_t smem       = fixed_smem + (size_t)p.kstride * 4;
    if (smem > 200 * 1024) return fail(B2A_ERR_UNSUPPORTED, "rdf: Rbin=%d needs %zu bytes of shared memory per block", s.Rbin, smem);

    // prep: slab ids and counts
    const int nslab = p.nslab;
    CU(cudaMemsetAsync(a.d_slab_cnt, 0, sizeof(unsigned) * 2 * (size_t)nslab * c->maxK, stm));
    int  bits = 0, reach[3] = { 0, 0, 0 };
    const bool cells = choose_cells(c, a, p, p.sym_enabled != 0, &bits, reach, kCwTile, nullptr, true);
    CellDev cd{};
    {
        Timed tm(c, 1, stm);
        dim3 grid((p.n0 + 255) / 256, c->nframes);
        k3_slab_count<<<grid, 256, sizeof(unsigned) * nslab, stm>>>(g0->d_cen[s.com0], c->nframes, p, a.d_slab_id, a.d_slab_cnt);
        CU(cudaGetLastError());
        a.sortedA_view = nullptr;
        if (!cells)
        {
            k3_slab_scatter<<<grid, 256, sizeof(unsigned) * 3 * nslab, stm>>>(g0->d_fix[s.com0], a.d_slab_id, c->nframes, p.n0, nslab,
                                                                                 a.d_slab_cnt, a.d_slab_cnt + (size_t)nslab * c->maxK, a.d_sortedA);
            CU(cudaGetLastError());
        }
        else
        {
            // both groups sorted by cell: shared with the other accumulators of this batch (cell_sort)
            CellSort *sA = nullptr, *sB = nullptr;
            if (int rc = cell_sort(c, s.grp0, *g0, s.com0, bits, &a, p, &sA, stm)) return rc;
            if (int rc = cell_sort(c, s.grp1, *g1, s.com1, bits, nullptr, p, &sB, stm)) return rc;
            const unsigned  nkA = sA->nkeys, nkB = sB->nkeys;
            const unsigned *stA = sA->starts(c->maxK), *stB = sB->starts(c->maxK);
            a.sortedA_view = sA->sorted;
            cd.bits = bits; cd.rx = reach[0]; cd.ry = reach[1]; cd.rz = reach[2];
            cd.nkeysA = nkA; cd.nkeysB = nkB; cd.startA = stA; cd.startB = stB; cd.sortedB = sB->sorted;
            {
                // per-column z reach: two cells whose index offsets are (dx, dy, dz) have closest points sqrt(sum (max(|d| - 1, 0) cell_k)^2)
                // apart; a pair can only count if that is below the cut-off (the float-rounded Rm * Rm is below Rm^2 (1 + 1e-6)) plus the
                // distance the fixed-point rounding can move two centres (cells are cut from the fixed-point coordinates).  Smallest
                // cells and largest units over the frames of the batch: conservative for every frame.
                const int nc = 1 << bits;
                double    cell[3] = { 1e300, 1e300, 1e300 }, unit = 0.0;
                for (int f = 0; f < c->nframes; ++f)
                    for (int k = 0; k < 3; ++k)
                    {
                        cell[k] = std::min(cell[k], c->h_geom[f].L[k] / nc);
                        unit    = std::max(unit, c->h_geom[f].L[k] / 4294967296.0);
                    }
                const double reff = p.Rm * (1.0 + 1e-6) + 4.0 * std::sqrt(3.0) * unit, r2 = reff * reff;
                const bool   table = 2 * reach[0] + 1 < nc && 2 * reach[1] + 1 < nc && reach[0] <= 7 && reach[1] <= 7;
                for (int ix = 0; ix < 15; ++ix)
                    for (int iy = 0; iy < 15; ++iy)
                    {
                        int rz = reach[2];
                        if (table && ix <= 2 * reach[0] && iy <= 2 * reach[1])
                        {
                            const double mx = std::max(std::abs(ix - reach[0]) - 1, 0) * cell[0], my = std::max(std::abs(iy - reach[1]) - 1, 0) * cell[1];
                            const double left = r2 - mx * mx - my * my;
                            if (left < 0.0) rz = -1;
                            else rz = std::min(reach[2], (int)std::floor(std::sqrt(left) / cell[2]) + 1);
                        }
                        cd.rzc[ix * 15 + iy] = (signed char)rz;
                    }
                cd.use_rzc = (table && !getenv("B2A_NO_RZC")) ? 1 : 0;
            }
        }
    }

    Timed tm(c, 2, stm);
    if (cells)
    {
        RdfDev pc = p;
        const double per_col = (double)p.n0 / (double)(1u << (2 * bits));
        // every warp owns a 128-molecule tile of its column; rows of kCwWarps tiles, a block loops if a column is more crowded;
        // symmetric evaluation appends the rows of the tiles against themselves (k3_rdf_cellw)
        const int    ny      = (int)std::min(64.0, std::max(1.0, std::ceil(1.5 * per_col / (kCwTile * kCwWarps))));
        const size_t smem_cw = (size_t)kCwWarps * (32 * sizeof(int4) + kCwQueue * sizeof(unsigned)) + 16 + (size_t)pc.kstride * 4;
        // a tile is one block of one warp, and a launch of a few waves of equal blocks ends in a half-empty wave: deal the
        // neighbour columns of every tile to two or three blocks while the launch is short (config 5, 6.9 waves per launch:
        // split 1 / 2 / 3 / 4 -> 126.2 / 128.9 / 128.3 / 126.3 frames/s)
        const double waves = (double)p.n0 / kCwTile * c->nframes / ((double)c->sm_count * (B2A_CW_WARPS_PER_SM / kCwWarps) * kCwWarps);
        cd.split = waves >= 12.0 ? 1 : (waves >= 5.0 ? 2 : 3);
        if (const char* e = getenv("B2A_CW_SPLIT")) cd.split = std::max(1, std::min(8, atoi(e)));
        const dim3   grid(cd.nkeysA >> bits, (pc.sym_enabled ? cd.split + 1 : cd.split) * ny, c->nframes);
        if (a.paranoid)
            return cubic ? launch_rdf_cellw<true, true>(c, a, *g0, *g1, pc, cd, smem_cw, grid) : launch_rdf_cellw<false, true>(c, a, *g0, *g1, pc, cd, smem_cw, grid);
        return cubic ? launch_rdf_cellw<true, false>(c, a, *g0, *g1, pc, cd, smem_cw, grid) : launch_rdf_cellw<false, false>(c, a, *g0, *g1, pc, cd, smem_cw, grid);
    }
    constexpr int TI = 8 * 128;   // TI molecules of one slab per block
    bool          dbuf = false;
    {
        // full mode: j split so that the grid covers the machine several times
        const int max_tiles = (p.n0 + TI - 1) / TI + std::min(nslab, p.n0);
        const int itiles    = std::max(1, (p.n0 + TI - 1) / TI);
        // j split: the grid runs in waves of `slots` resident blocks and the under-filled last wave is lost time, so offer
        // many waves, but never chunks below two j tiles (per-block set-up, histogram zeroing and flush).  Measured on
        // config 3 (31 i-tiles x 8 frames): 4.4 waves -> 0.75 of FP32 issue, 7 -> 0.78, 10.4 -> 0.79; config 2 full
        // evaluation: 5.4 waves -> 0.85, 10.7 -> 0.87; config 3 with 64-frame batches: 10 / 14 / 20 / 30 / 45 / 90 waves ->
        // 1613 / 1629 / 1633 / 1641 / 1649 / 1649 frames/s (0.814 -> 0.833): shorter blocks also even out the SMs.
        const int resident = std::max(1, std::min(6, (int)((227 * 1024) / (smem + 1024))));   // 80 registers x 128 threads: 6
        // double-buffered j-tile only where its 8 KB do not cost a resident block
        dbuf = std::min(6, (int)((227 * 1024) / (smem + kTileJ * sizeof(int4) + 1024))) >= resident;
        if (const char* e = getenv("B2A_K3_DBUF")) dbuf = atoi(e) != 0;
        int       waves    = 45;
        if (const char* e = getenv("B2A_K3_WAVES")) waves = std::max(1, atoi(e));
        int       jsplit   = std::max(1, (c->sm_count * resident * waves + itiles * c->nframes - 1) / (itiles * c->nframes));
        jsplit             = std::min(jsplit, std::max(1, p.n1 / (2 * kTileJ)));
        p.jchunk           = (p.n1 + jsplit - 1) / jsplit;
        p.jchunk           = ((p.jchunk + kTileJ - 1) / kTileJ) * kTileJ;
        p.jsplit           = (p.n1 + p.jchunk - 1) / p.jchunk;
        if (int rc = dispatch_rdf<8, 0>(c, a, *g0, *g1, p, smem, dim3(max_tiles, p.jsplit, c->nframes), cubic, dbuf)) return rc;
    }
    if (p.sym_enabled)
    {
        // symmetric mode: work items = per i-tile one diagonal block + chunks of the molecules beyond the tile
        RdfDev ps  = p;
        ps.jchunk  = 4 * kTileJ;
        if (const char* e = getenv("B2A_K3_SYMCHUNK")) ps.jchunk = std::max(1, atoi(e)) * kTileJ;
        const int T = (p.n0 + TI - 1) / TI;
        long long items = 0;
        for (int t = 0; t < T; ++t) items += 1 + (std::max(0, p.n0 - (t + 1) * TI) + ps.jchunk - 1) / ps.jchunk;
        if (int rc = dispatch_rdf<8, 1>(c, a, *g0, *g1, ps, smem, dim3((unsigned)items, 1, c->nframes), cubic, dbuf)) return rc;
    }
    return B2A_OK;
}

static int accumulate_region(b2a_ctx* c, Acc& a)
{
    const b2a_region_spec& s = a.region;
    Group* g = nullptr;
    if (int rc = get_group(c, s.grp, &g)) return rc;
    if (int rc = ensure_centers(c, *g, s.com, false)) return rc;   // K1: exact FP64 centres, cached for the batch
    RegionDev d{};
    d.nfilter = s.nfilter; d.wrap = s.wrap; d.cyl = s.cyl; d.nbdim = s.nbdim;
    for (int q = 0; q < 3; ++q) { d.fdim[q] = s.fdim[q]; d.fincl[q] = s.fincl[q]; d.flow[q] = s.flow[q]; d.fup[q] = s.fup[q]; }
    d.cx = (double)s.cx; d.cy = (double)s.cy; d.R2 = s.R2;
    for (int b = 0; b < 2; ++b) { d.bdim[b] = s.bdim[b]; d.nb[b] = s.nb[b]; d.blow[b] = (double)s.blow[b]; d.dbin[b] = s.dbin[b]; }
    Timed tm(c, 3);
    dim3  grid((g->nmol + 255) / 256, c->nframes);
    unsigned long long* pf = nullptr;   // this batch's slice of the per-frame log
    if (s.per_frame)
    {
        if (int rc = series_reserve(c, a, (size_t)c->nframes)) return rc;
        pf = a.d_perframe + a.series_len * a.ndev;
        CU(cudaMemsetAsync(pf, 0, sizeof(unsigned long long) * a.ndev * c->nframes, c->stream));
    }
    k2g_region<<<grid, 256, 0, c->stream>>>(g->d_cen[s.com], c->nframes, g->nmol, d, c->d_geom, a.d_hist, a.d_hist + a.ndev, pf, a.ndev);
    CU(cudaGetLastError());
    if (s.per_frame)
    {
        a.batches.push_back({ c->frame_base, c->nframes });
        a.series_len += (size_t)c->nframes;
    }
    a.perframe_frames = c->nframes;
    return B2A_OK;
}

extern "C" int b2a_region_read_frames(b2a_ctx* c, int id, unsigned long long* out)
{
    TEAM_CUR(c, b2a_region_read_frames(m, id, out));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind != ACC_REGION || !a->d_perframe) return fail(B2A_ERR_ARG, "accumulator %d keeps no per-frame counters (b2a_region_spec.per_frame)", id);
    if (!out) return fail(B2A_ERR_ARG, "b2a_region_read_frames: null output");
    CU(cudaSetDevice(c->device));
    if (a->perframe_frames <= 0 || a->series_len < (size_t)a->perframe_frames) return fail(B2A_ERR_STATE, "nothing accumulated yet");
    CU(cudaMemcpyAsync(out, a->d_perframe + (a->series_len - a->perframe_frames) * a->ndev, sizeof(unsigned long long) * a->ndev * a->perframe_frames,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

static int accumulate_density(b2a_ctx* c, Acc& a)
{
    const b2a_density_spec& s = a.den;
    Group* g = nullptr;
    if (int rc = get_group(c, s.grp, &g)) return rc;
    if (c->nframes <= 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet");
    DensityDev d{};
    d.dim = s.dim; d.dim2 = s.dim2; d.nbin = s.nbin; d.upper_inclusive = s.upper_inclusive;
    d.low = (double)s.low; d.up = (double)s.up; d.low2 = (double)s.low2; d.up2 = (double)s.up2; d.dbin = make_div(s.dbin);
    if (g->csr)
    {
        if (int rc = ensure_centers(c, *g, s.com, false)) return rc;
        Timed tm(c, 3);
        dim3  grid((g->nmol + 255) / 256, c->nframes);
        k2_density_from_centers<<<grid, 256, 0, c->stream>>>(g->d_cen[s.com], c->nframes, g->nmol, d, a.d_hist, a.d_hist + a.ndev);
        CU(cudaGetLastError());
        return B2A_OK;
    }
    const int    threads   = 256;
    const int    smem_bins = s.nbin <= 24000 ? s.nbin : 0;
    int          M = 0;
    if (stream_block(*g, &M))
    {
        Timed      tm(c, 3);
        StreamArgs sa{};
        sa.den = d; sa.hist = a.d_hist; sa.stats = a.d_hist + a.ndev; sa.smem_bins = smem_bins;
        sa.df.low = s.low; sa.df.up = s.up; sa.df.low2 = s.low2; sa.df.up2 = s.up2; sa.df.inv_dbin = (float)(1.0 / s.dbin);
        return launch_stream<1>(c, *g, s.com, M, sa, sizeof(unsigned) * smem_bins, a.fast);
    }
    const size_t smem      = sizeof(double) * g->napm + sizeof(unsigned) * smem_bins;
    CU(cudaFuncSetAttribute(k2_density, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 blocks((g->nmol + threads - 1) / threads, (c->nframes + kFramesPerBlock - 1) / kFramesPerBlock);
    Timed      tm(c, 3);
    k2_density<<<blocks, threads, smem, c->stream>>>(c->d_x, (size_t)c->natoms * 3, c->nframes, g->d_index, g->nmol, g->napm,
                                                     g->d_w[s.com], make_div(g->wsum[s.com]), c->d_geom, d, a.d_hist, a.d_hist + a.ndev, smem_bins);
    CU(cudaGetLastError());
    return B2A_OK;
}

static int accumulate_orient(b2a_ctx* c, Acc& a);
static int accumulate_coord(b2a_ctx* c, Acc& a);

__global__ void k_add_frames(unsigned long long* stats, unsigned long long n) { stats[0] += n; }

// Per-frame constants travel host -> device through a KERNEL that reads the pinned slot over the bus (pinned memory is mapped into
// the device's address space), not through cudaMemcpyAsync: a copy would queue on the H2D copy engine BEHIND the upload of the
// next frame batch (hundreds of MB on the copy stream), and the pair kernel that needs these few hundred bytes would wait for
// all of it (measured: 3 ms of a 98 ms step on config 2, 7 of 27 ms on config 3).
__global__ void k_fetch_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src_pinned, int nwords)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) dst[i] = src_pinned[i];
}

extern "C" int b2a_accumulate(b2a_ctx* c, int id)
{
    TEAM_CUR(c, b2a_accumulate(m, id));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    CU(cudaSetDevice(c->device));
    if (c->nframes <= 0) return fail(B2A_ERR_STATE, "no frame batch submitted yet");
    int rc = B2A_OK;
    switch (a->kind)
    {
        case ACC_DENSITY: rc = accumulate_density(c, *a); break;
        case ACC_RDF: rc = accumulate_rdf(c, *a); break;
        case ACC_ORIENT: rc = accumulate_orient(c, *a); break;
        case ACC_COORD: rc = accumulate_coord(c, *a); break;
        case ACC_REGION: rc = accumulate_region(c, *a); break;
    }
    if (rc) return rc;
    if (a->kind == ACC_RDF && a->last_stream && a->last_stream != c->stream)
    {   // the accumulation went to the accumulator's side stream: the frame counter follows it (before the join event)
        k_add_frames<<<1, 1, 0, a->last_stream>>>(a->d_hist + a->ndev, (unsigned long long)c->nframes);
        CU(cudaGetLastError());
        CU(cudaEventRecord(a->ev_done, a->last_stream));
    }
    else
    {
        k_add_frames<<<1, 1, 0, c->stream>>>(a->d_hist + a->ndev, (unsigned long long)c->nframes);
        CU(cudaGetLastError());
    }
    a->epoch++;
    return B2A_OK;
}

extern "C" int b2a_acc_reset(b2a_ctx* c, int id)
{
    TEAM_ALL(c, b2a_acc_reset(m, id));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    CU(cudaMemsetAsync(a->d_hist, 0, sizeof(unsigned long long) * (a->ndev + B2A_STAT_NSLOTS), c->stream));
    a->epoch++;
    a->series_len = 0;
    a->batches.clear();
    a->perframe_frames = a->table_frames = 0;
    a->gathered.clear(); a->gathered_idx.clear();
    return B2A_OK;
}

extern "C" int b2a_acc_size(b2a_ctx* c, int id, size_t* n)
{
    TEAM_FIRST(c, b2a_acc_size(m, id, n));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (!n) return fail(B2A_ERR_ARG, "null argument");
    *n = a->ndev;   // device words before the stats block; the merge covers ndev + B2A_STAT_NSLOTS
    return B2A_OK;
}

extern "C" int b2a_acc_public_size(b2a_ctx* c, int id, size_t* n)
{
    TEAM_FIRST(c, b2a_acc_public_size(m, id, n));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (!n) return fail(B2A_ERR_ARG, "null argument");
    *n = a->ncounters;
    return B2A_OK;
}

extern "C" int b2a_acc_device_ptr(b2a_ctx* c, int id, void** p)
{
    TEAM_FIRST(c, b2a_acc_device_ptr(m, id, p));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (!p) return fail(B2A_ERR_ARG, "null argument");
    *p = a->d_hist;
    return B2A_OK;
}

extern "C" int b2a_acc_read(b2a_ctx* c, int id, uint64_t* counts, uint64_t* stats)
{
    if (c && c->is_leader()) return team_read(c, id, counts, stats);   // merges the devices first (b2a_finalize) when needed
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (!counts) return fail(B2A_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    // after b2a_finalize (and until the next accumulate / reset) the merged totals are what a read returns
    const unsigned long long* src = (a->d_merged && a->merged_epoch == a->epoch) ? a->d_merged : a->d_hist;
    std::vector<unsigned long long> h(a->ndev + B2A_STAT_NSLOTS);
    CU(cudaMemcpyAsync(h.data(), src, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long* st = h.data() + a->ndev;
    if (a->kind == ACC_RDF)
    {
        const int nbin = a->rdf.nbin, Rbin = a->rdf.Rbin;
        // device layout [slab][meR] -> reference layout meZ + meR*nbin (itp::matd(nbin, Rbin), column-major)
        for (int z = 0; z < nbin; ++z)
            for (int r = 0; r < Rbin; ++r) counts[(size_t)z + (size_t)r * nbin] = h[(size_t)z * Rbin + r];
        unsigned long long oob = 0;   // pairs of i whose slab index == nbin: out-of-bounds writes in the reference
        for (int r = 0; r < Rbin; ++r) oob += h[(size_t)nbin * Rbin + r];
        st[B2A_STAT_DROPPED] += oob;
    }
    else
    {
        for (size_t i = 0; i < a->ncounters; ++i) counts[i] = h[i];
    }
    if (stats) for (int i = 0; i < B2A_STAT_NSLOTS; ++i) stats[i] = st[i];
    return B2A_OK;
}

extern "C" int b2a_rdf_normalise(const uint64_t* counts, int nbin, int Rbin, double* g)
{
    if (!counts || !g || nbin <= 0 || Rbin <= 0) return fail(B2A_ERR_ARG, "b2a_rdf_normalise: bad argument");
    // calc_rdf_region.cpp:100-109: shell "volumes" from 0.002-wide shells (NOT Rm/Rbin), :111 divide
    for (int r = 0; r < Rbin; ++r)
    {
        double v = std::pow(0.002 * (r + 1), 3);
        if (r > 0) v = v - std::pow(0.002 * r, 3);
        for (int z = 0; z < nbin; ++z) g[(size_t)z + (size_t)r * nbin] = (double)counts[(size_t)z + (size_t)r * nbin] / v;
    }
    // :113-115: each slab divided by the mean of its last int(Rbin*0.1) bins (0/0 -> NaN rows, as the reference prints)
    const int tail = (int)(Rbin * 0.1);
    for (int z = 0; z < nbin; ++z)
    {
        double s = 0.0;
        for (int r = Rbin - tail; r < Rbin; ++r) s += g[(size_t)z + (size_t)r * nbin];
        const double mean = s / tail;
        for (int r = 0; r < Rbin; ++r) g[(size_t)z + (size_t)r * nbin] /= mean;
    }
    return B2A_OK;
}

// ---- measurement support ---------------------------------------------------------------------------------------
extern "C" int b2a_invalidate(b2a_ctx* c)
{
    TEAM_ALL(c, b2a_invalidate(m));
    if (!c) return fail(B2A_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;   // the centre caches and sorted views about to be rebuilt may still be read
    c->serial++;
    return B2A_OK;
}

// development aid (not in b2a.h): time stamps written by the last streaming launch under B2A_STREAM_DEBUG
extern "C" int b2a_debug_read(b2a_ctx* c, unsigned long long* out, int nwords)
{
    if (!c || !out || !c->d_dbg || nwords > 8 * 4096) return fail(B2A_ERR_STATE, "no debug buffer");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out, c->d_dbg, sizeof(unsigned long long) * nwords, cudaMemcpyDeviceToHost));
    return B2A_OK;
}

extern "C" int b2a_timing_enable(b2a_ctx* c, int enable)
{
    TEAM_ALL(c, b2a_timing_enable(m, enable));
    if (!c) return fail(B2A_ERR_ARG, "null context");
    c->timing = enable != 0;
    return B2A_OK;
}

extern "C" int b2a_timing_read(b2a_ctx* c, double* ms, long long* launches)
{
    if (!c || !ms || !launches) return fail(B2A_ERR_ARG, "null argument");
    if (c->is_leader())
    {   // totals over the devices
        for (int i = 0; i < 8; ++i) { ms[i] = 0; launches[i] = 0; }
        for (b2a_ctx* m : c->members)
        {
            double t[8]; long long n[8];
            if (int rc = b2a_timing_read(m, t, n)) return rc;
            for (int i = 0; i < 8; ++i) { ms[i] += t[i]; launches[i] += n[i]; }
        }
        return B2A_OK;
    }
    CU(cudaSetDevice(c->device));
    if (int rc = join_side(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    for (auto& s : c->spans)
    {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, s.a, s.b));
        c->t_ms[s.slot] += t;
        c->t_n[s.slot]++;
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    c->spans.clear();
    for (int i = 0; i < 8; ++i) { ms[i] = c->t_ms[i]; launches[i] = c->t_n[i]; c->t_ms[i] = 0; c->t_n[i] = 0; }
    return B2A_OK;
}

// ---- orientation -------------------------------------------------------------------------------------------------
extern "C" int b2a_orient_create(b2a_ctx* c, const b2a_orient_spec* s, int* id)
{
    TEAM_ALL(c, b2a_orient_create(m, s, id));   // created in the same order on every member: same id everywhere
    if (!c || !s || !id) return fail(B2A_ERR_ARG, "b2a_orient_create: null argument");
    Group* g = nullptr;
    if (int rc = get_group(c, s->grp, &g)) return rc;
    if (s->DIMN < 0 || s->DIMN > 2 || s->DIMNOrient < 0 || s->DIMNOrient > 2) return fail(B2A_ERR_ARG, "orient: dim out of range");
    // heterogeneous (GmxHandleFull) group: every molecule must own the three vector atoms
    const int napm_lim = g->csr ? g->min_napm : g->napm;
    if (s->tp1 < 1 || s->tp2 < 1 || s->tp3 < 1 || s->tp1 > napm_lim || s->tp2 > napm_lim || s->tp3 > napm_lim)
        return fail(B2A_ERR_ARG, "orient: tp1..3 must lie in 1..napm(%d)", napm_lim);
    if (s->nAngle <= 0 || s->nAngle > 180) return fail(B2A_ERR_ARG, "orient: nAngle must be 1..180 (dAngle = 180 / nAngle is integer division)");
    if (s->method != 1 && s->method != 2) return fail(B2A_ERR_ARG, "orient: method must be 1 or 2");
    if (s->com < 0 || s->com > 3) return fail(B2A_ERR_ARG, "com out of range");
    const int nbin1 = (int)((s->CR - s->Cr) / s->dr);   // float arithmetic, calc_orient_mof.cpp:133
    if (nbin1 <= 0) return fail(B2A_ERR_ARG, "orient: (CR - Cr) / dr gives no radial bins");
    CU(cudaSetDevice(c->device));
    std::unique_ptr<Acc> a(new Acc);
    a->kind  = ACC_ORIENT;
    a->ori   = *s;
    a->nbin1 = nbin1;
    a->ncounters = a->ndev = (size_t)nbin1 * s->nAngle + nbin1;
    // Angle-bin edges as thresholds on cos(angle): the reference evaluates
    //   angle = acos(c) / pi(long double) * 180 ; meA = int(angle / dAngle)        (calc_orient_mof.cpp:41-47, 193-194)
    // with the HOST libm.  acos is monotone, so bin(c) = #{k >= 1 : c <= edge[k]} where edge[k] is the largest
    // double c whose reference bin is >= k.  The table is found here by bisection with the host's own acos and
    // long double arithmetic, and the device only compares -- it never evaluates a transcendental.
    const int          nA     = s->nAngle;
    const double       dAngle = (double)(180 / nA);
    std::vector<double> edge(nA + 1);
    auto bin_of = [&](double cv) -> int {
        const long double pi = 3.14159265358979323846264338327950288L;
        const double      ang = (double)(std::acos(cv) / pi * 180);
        return (int)(ang / dAngle);
    };
    for (int k = 1; k <= nA; ++k)
    {
        // largest c in [-1, 1] with bin_of(c) >= k  (bin_of is non-increasing in c)
        if (bin_of(-1.0) < k) { edge[k] = -2.0; continue; }   // never reached
        double lo = -1.0, hi = 1.0;                            // invariant: bin_of(lo) >= k
        if (bin_of(hi) >= k) { edge[k] = 1.0; continue; }
        for (;;)
        {
            const double mid = lo + (hi - lo) / 2;
            if (mid <= lo || mid >= hi) break;
            if (bin_of(mid) >= k) lo = mid; else hi = mid;
        }
        // lo and hi are now adjacent doubles
        edge[k] = lo;
    }
    edge[0] = 2.0;
    CU(cudaMalloc(&a->d_cosedge, sizeof(double) * (nA + 1)));
    CU(cudaMemcpy(a->d_cosedge, edge.data(), sizeof(double) * (nA + 1), cudaMemcpyHostToDevice));
    std::vector<float> fedge(nA + 3, -3.0f);   // two sentinels below every cosine: the loop-free bin search reads fedge[lo + 2]
    for (int k = 0; k <= nA; ++k) fedge[k] = (float)edge[k];
    CU(cudaMalloc(&a->d_fedge, sizeof(float) * (nA + 3)));
    CU(cudaMemcpy(a->d_fedge, fedge.data(), sizeof(float) * (nA + 3), cudaMemcpyHostToDevice));
    {   // fast path: cos cell ci = [-1 + 2 ci / N, -1 + 2 (ci + 1) / N) -> the lowest angle bin any value below the upper end of the
        // NEXT cell can have (one cell of slack for the float rounding of the cell index); the device walks up from there
        std::vector<unsigned char> lut(kCosLut);
        for (int ci = 0; ci < kCosLut; ++ci)
        {
            const float ub = (float)(-1.0 + 2.0 * (ci + 2) / kCosLut);
            int         lo = 0;
            while (lo < nA && ub <= fedge[lo + 1]) ++lo;
            lut[ci] = (unsigned char)lo;
        }
        CU(cudaMalloc(&a->d_lut, kCosLut));
        CU(cudaMemcpy(a->d_lut, lut.data(), kCosLut, cudaMemcpyHostToDevice));
        // interval form of "bin == k": edge[k + 1] < c <= edge[k]  <=  |c - mid| < hw, with hw shrunk by more than the float
        // roundings of mid, of c - mid and of hw - E on the device (7.5 * 2^-24 for |values| <= 3) and rounded towards zero
        std::vector<float2> fmh(nA + 2);
        for (int k = 0; k <= nA + 1; ++k)
        {
            const double up = k <= nA ? edge[k] : -3.0, dn = k < nA ? edge[k + 1] : -3.0;
            const double hw = (up - dn) / 2 - 6e-7;
            fmh[k].x = (float)((up + dn) / 2);
            fmh[k].y = hw > 0 ? std::nextafter((float)hw, 0.0f) : 0.0f;
        }
        CU(cudaMalloc(&a->d_fmh, sizeof(float2) * (nA + 2)));
        CU(cudaMemcpy(a->d_fmh, fmh.data(), sizeof(float2) * (nA + 2), cudaMemcpyHostToDevice));
    }
    // Radial thresholds in s = t1^2 + t2^2 (calc_orient_mof.cpp:175-179): R = sqrt(s) (IEEE, identical on host and
    // device), selected iff Cr <= R <= CR, meZ = int((R - Cr) / dr).  Everything is monotone in s, so bisection over the
    // bit patterns of positive doubles finds, for every k, the smallest s with meZ >= k; the device only compares.
    if (nbin1 > 8000) return fail(B2A_ERR_UNSUPPORTED, "orient: more than 8000 radial bins");
    const double Cr = (double)s->Cr, CR = (double)s->CR, dr = (double)s->dr;
    auto rbin = [&](double sv) -> double {             // -1: R < Cr; else (R - Cr) / dr, whose truncation is meZ
        const double R = std::sqrt(sv);
        if (!(Cr <= R)) return -1.0;
        return (R - Cr) / dr;                          // >= 0 here; int(q) >= k  <=>  q >= k for integer k >= 0
    };
    auto from_bits = [](uint64_t b) { double v; std::memcpy(&v, &b, 8); return v; };
    auto smallest_s_with = [&](auto pred) -> double {   // pred monotone false..true over s in [0, +max]; returns +inf if never true
        uint64_t lo = 0, hi = 0x7FEFFFFFFFFFFFFFull;      // bit patterns of +0 .. DBL_MAX
        if (pred(from_bits(lo))) return 0.0;
        if (!pred(from_bits(hi))) return INFINITY;
        while (hi - lo > 1) { const uint64_t mid = lo + (hi - lo) / 2; if (pred(from_bits(mid))) hi = mid; else lo = mid; }
        return from_bits(hi);
    };
    std::vector<double> stab(nbin1 + 2);
    for (int k = 0; k <= nbin1; ++k) stab[k] = smallest_s_with([&](double sv) { return rbin(sv) >= (double)k; });
    {   // largest s with sqrt(s) <= CR  ==  predecessor of the smallest s with sqrt(s) > CR
        const double first_above = smallest_s_with([&](double sv) { return std::sqrt(sv) > CR; });
        stab[nbin1 + 1] = std::isinf(first_above) ? first_above : std::nextafter(first_above, -INFINITY);
        if (first_above == 0.0) stab[nbin1 + 1] = -1.0;   // CR < 0: nothing is ever selected
    }
    CU(cudaMalloc(&a->d_stab, sizeof(double) * (nbin1 + 2)));
    CU(cudaMemcpy(a->d_stab, stab.data(), sizeof(double) * (nbin1 + 2), cudaMemcpyHostToDevice));
    return new_acc(c, std::move(a), id);
}

static int accumulate_orient(b2a_ctx* c, Acc& a)
{
    const b2a_orient_spec& s = a.ori;
    Group* g = nullptr;
    if (int rc = get_group(c, s.grp, &g)) return rc;
    OrientDev d{};
    d.DIMN = s.DIMN; d.nAngle = s.nAngle; d.nbin1 = a.nbin1; d.tp1 = s.tp1 - 1; d.tp2 = s.tp2 - 1; d.tp3 = s.tp3 - 1;
    d.DIMNOrient = s.DIMNOrient; d.method = s.method;
    d.low = (double)s.low; d.up = (double)s.up; d.c1 = (double)s.c1; d.c2 = (double)s.c2;
    d.CR = (double)s.CR; d.Cr = (double)s.Cr; d.Crf = s.Cr; d.inv_drf = 1.0f / s.dr;
    const int    threads = 256;
    const size_t hist_words = (size_t)a.nbin1 * s.nAngle + a.nbin1;
    const int    smem_bins  = hist_words <= 24000 ? (int)hist_words : 0;
    int          M = 0;
    if (g->csr)
    {
        // heterogeneous molecule sizes: literal FP64 per molecule with its own atom count, weights and divisor
        const size_t smem = sizeof(double) * (s.nAngle + 1 + a.nbin1 + 2) + sizeof(float) * ((s.nAngle + 2) & ~1) + sizeof(unsigned) * smem_bins;
        CU(cudaFuncSetAttribute(k4_orient_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 blocks((g->nmol + threads - 1) / threads, (c->nframes + kFramesPerBlock - 1) / kFramesPerBlock);
        Timed      tm(c, 4);
        k4_orient_csr<<<blocks, threads, smem, c->stream>>>(c->d_x, (size_t)c->natoms * 3, c->nframes, g->d_index, g->d_off, g->nmol, g->d_wfull[s.com],
                                                            g->d_div[s.com], c->d_geom, d, a.d_cosedge, a.d_fedge, a.d_stab, a.d_hist, a.d_hist + a.ndev,
                                                            smem_bins);
        CU(cudaGetLastError());
        return B2A_OK;
    }
    if (stream_block(*g, &M))
    {
        Timed      tm(c, 4);
        StreamArgs sa{};
        sa.ori = d; sa.cosedge = a.d_cosedge; sa.fedge = a.d_fedge; sa.stab = a.d_stab; sa.hist = a.d_hist; sa.stats = a.d_hist + a.ndev; sa.smem_bins = smem_bins;
        sa.of.low = s.low; sa.of.up = s.up; sa.of.c1 = s.c1; sa.of.c2 = s.c2; sa.of.Cr = s.Cr; sa.of.CR = s.CR;
        sa.of.inv_dr = 1.0f / s.dr; sa.of.cr_lo = s.Cr > 0.f ? s.Cr : -3.0e38f;
        sa.lut = a.d_lut; sa.fmh = a.d_fmh;
        // the bracket needs monotone radial bins
        const int fast = (s.dr > 0.f && s.CR > s.Cr) ? a.fast : 0;
        // three-site molecule with vectors (reference atom -> the other two): packed fast path orient_fast3
        // two molecules per thread in packed FP32 whenever the histogram fits in shared memory; three-site molecules whose
        // vectors run from the reference atom to the other two additionally reuse the unwrapped offsets as the vectors
        int var = smem_bins > 0 ? 1 : 0;
        sa.ori_sign = 1.f;
        {   // shared-memory bank conflicts of the one-molecule-per-lane reads: g lanes per bank, g = gcd(3 napm, 32)
            int gcd = 1;
            while (gcd < 32 && (3 * g->napm) % (2 * gcd) == 0) gcd *= 2;
            int sh = 5;
            for (int v = gcd; v > 1; v >>= 1) --sh;
            sa.rot_shift = (gcd >= 8 && gcd <= g->napm && !getenv("B2A_NO_ROT")) ? sh : 5;   // the walk needs g distinct atoms; measured
            // (tools/bench_hbm_napm.py): 8 atoms 0.34 -> 0.43, 16 atoms 0.21 -> 0.32 of HBM peak; 2- and 4-way conflicts (6, 12 atoms) are cheaper
            // than the extra atom and index arithmetic of the rotated walk
        }
        if (g->napm == 3 && d.tp1 == 0 && ((d.tp2 == 1 && d.tp3 == 2) || (d.tp2 == 2 && d.tp3 == 1)))
        {
            sa.ori_reuse = 1;
            sa.ori_sign  = s.method == 1 ? -1.f : (d.tp2 == 1 ? 1.f : -1.f);
        }
        if (const char* e = getenv("B2A_ORIENT_VAR")) var = std::min(var, atoi(e));
        // one pass for two products: the kernel also writes the exact centres of kind also_com (unless this batch has them already)
        const int  ac    = a.also_com;
        const bool fused = ac >= 0 && g->valid[ac] != c->serial && !getenv("B2A_NO_FUSE");
        if (fused)
        {
            if (!g->d_cen[ac]) CU(cudaMalloc(&g->d_cen[ac], sizeof(double) * 3 * (size_t)g->nmol * c->maxK));
            sa.cen = g->d_cen[ac]; sa.w2 = g->d_w[ac]; sa.wsum2 = make_div(g->wsum[ac]);
        }
        struct Mark { Group* g; int ac; uint64_t serial; bool on; ~Mark() { if (on) g->valid[ac] = serial; } } mark{ g, ac, c->serial, fused };
        return launch_stream<2>(c, *g, s.com, M, sa, sizeof(double) * (s.nAngle + 1 + a.nbin1 + 2) + sizeof(float) * ((s.nAngle + 4) & ~1) + sizeof(float2) * (s.nAngle + 2) + sizeof(unsigned) * smem_bins, fast, var);
    }
    const size_t smem       = sizeof(double) * (g->napm + s.nAngle + 1 + a.nbin1 + 2) + sizeof(float) * ((s.nAngle + 2) & ~1) + sizeof(unsigned) * smem_bins;
    CU(cudaFuncSetAttribute(k4_orient, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 blocks((g->nmol + threads - 1) / threads, (c->nframes + kFramesPerBlock - 1) / kFramesPerBlock);
    Timed      tm(c, 4);
    k4_orient<<<blocks, threads, smem, c->stream>>>(c->d_x, (size_t)c->natoms * 3, c->nframes, g->d_index, g->nmol, g->napm,
                                                    g->d_w[s.com], make_div(g->wsum[s.com]), c->d_geom, d, a.d_cosedge, a.d_fedge, a.d_stab, a.d_hist,
                                                    a.d_hist + a.ndev, smem_bins);
    CU(cudaGetLastError());
    return B2A_OK;
}

// ---- coordination numbers (K5) -------------------------------------------------------------------------------------
extern "C" int b2a_coord_create(b2a_ctx* c, const b2a_coord_spec* s, int* id)
{
    TEAM_ALL(c, b2a_coord_create(m, s, id));   // created in the same order on every member: same id everywhere
    if (!c || !s || !id) return fail(B2A_ERR_ARG, "b2a_coord_create: null argument");
    Group *g0 = nullptr, *g1 = nullptr;
    if (int rc = get_group(c, s->grp0, &g0)) return rc;
    if (int rc = get_group(c, s->grp1, &g1)) return rc;
    if (!(s->Rmin > 0.f) || s->dim < 0 || s->dim > 2 || s->max_count <= 0 || s->max_count > 8192)
        return fail(B2A_ERR_ARG, "b2a_coord_create: bad spec (Rmin=%g dim=%d max_count=%d)", (double)s->Rmin, s->dim, s->max_count);
    if (s->com0 < 0 || s->com0 > 3 || s->com1 < 0 || s->com1 > 3) return fail(B2A_ERR_ARG, "com out of range");
    CU(cudaSetDevice(c->device));
    std::unique_ptr<Acc> a(new Acc);
    a->kind      = ACC_COORD;
    a->coord     = *s;
    a->ncounters = a->ndev = (size_t)s->max_count;
    CU(cudaMalloc(&a->d_slab_id, sizeof(int) * (size_t)g0->nmol * c->maxK));
    CU(cudaMalloc(&a->d_slab_cnt, sizeof(unsigned) * 2 * (size_t)c->maxK));
    CU(cudaMalloc(&a->d_sortedA, sizeof(int4) * (size_t)g0->nmol * c->maxK));
    CU(cudaMalloc(&a->d_cnt, sizeof(int) * (size_t)g0->nmol * c->maxK));
    if (int rc = series_reserve(c, *a, (size_t)c->maxK)) return rc;
    CU(cudaMalloc(&a->d_cframes, sizeof(CoordFrame) * c->maxK));
    CU(cudaMallocHost(&a->h_cframes, sizeof(CoordFrame) * c->maxK * kFrameRing));
    a->alloc_n0 = g0->nmol; a->alloc_n1 = g1->nmol;
    return new_acc(c, std::move(a), id);
}

template <int IPT, int MODE>
static int launch_coord(b2a_ctx* c, Acc& a, Group& g0, Group& g1, const CoordDev& p, const CellDev& cd, dim3 grid, bool cubic)
{
    const b2a_coord_spec& s = a.coord;
    if (cubic)
        k5_coord<IPT, 128, true, MODE><<<grid, 128, 0, c->stream>>>(a.d_sortedA, g1.d_fix[s.com1], g0.d_cen[s.com0], g1.d_cen[s.com1], a.d_slab_cnt,
                                                                  a.d_cframes, p, cd, a.d_cnt, a.d_hist + a.ndev);
    else
        k5_coord<IPT, 128, false, MODE><<<grid, 128, 0, c->stream>>>(a.d_sortedA, g1.d_fix[s.com1], g0.d_cen[s.com0], g1.d_cen[s.com1], a.d_slab_cnt,
                                                                   a.d_cframes, p, cd, a.d_cnt, a.d_hist + a.ndev);
    CU(cudaGetLastError());
    return B2A_OK;
}

static int accumulate_coord(b2a_ctx* c, Acc& a)
{
    const b2a_coord_spec& s = a.coord;
    Group *g0 = nullptr, *g1 = nullptr;
    if (int rc = get_group(c, s.grp0, &g0)) return rc;
    if (int rc = get_group(c, s.grp1, &g1)) return rc;
    if (int rc = ensure_centers(c, *g0, s.com0, true)) return rc;
    if (int rc = ensure_centers(c, *g1, s.com1, true)) return rc;
    const int n0 = g0->nmol, n1 = g1->nmol;

    if (n0 > a.alloc_n0)
    {   // group 0 redefined with more molecules after b2a_coord_create
        CU(cudaStreamSynchronize(c->stream));
        if (int rc = regrow(&a.d_slab_id, (size_t)n0 * c->maxK)) return rc;
        if (int rc = regrow(&a.d_sortedA, (size_t)n0 * c->maxK)) return rc;
        if (int rc = regrow(&a.d_cnt, (size_t)n0 * c->maxK)) return rc;
        if (a.d_keysA) { if (int rc = regrow(&a.d_keysA, (size_t)n0 * c->maxK)) return rc; }
        a.alloc_n0 = n0;
    }
    if (n1 > a.alloc_n1)
    {
        CU(cudaStreamSynchronize(c->stream));
        if (a.d_keysB) { if (int rc = regrow(&a.d_keysB, (size_t)n1 * c->maxK)) return rc; }
        if (a.d_sortedB) { if (int rc = regrow(&a.d_sortedB, (size_t)n1 * c->maxK)) return rc; }
        a.alloc_n1 = n1;
    }
    bool cubic = true;
    int  slot  = 0;
    if (int rc = ring_acquire(a, &slot)) return rc;
    CoordFrame* const h_cframes = a.h_cframes + (size_t)slot * c->maxK;
    for (int f = 0; f < c->nframes; ++f)
        if (c->h_geom[f].L[0] != c->h_geom[f].L[1] || c->h_geom[f].L[0] != c->h_geom[f].L[2]) cubic = false;
    // FP32 error of r^2 (x fixed units^2): 3 I2FP + FMUL/FFMA/FFMA -> 5 * 2^-24 relative (7 * 2^-24 with the two axis-ratio
    // multiplies of a non-cubic box), plus the 2^32 grid: every component of the difference is within 1 unit of the truth, so
    // |r - r_true| <= sqrt(3) a, a = the largest axis ratio, and |r^2 - r_true^2| <= 2 sqrt(3) a r + 3 a^2.
    const double erel = (cubic ? 5.0 : 7.0) * std::ldexp(1.0, -24) + 1e-7;
    for (int f = 0; f < c->nframes; ++f)
    {
        const FrameGeom& g  = c->h_geom[f];
        CoordFrame&      fr = h_cframes[f];
        const double     unit = g.L[0] / 4294967296.0;
        fr.ay = (float)(g.L[1] / g.L[0]);
        fr.az = (float)(g.L[2] / g.L[0]);
        const double amax = std::max(1.0, std::max((double)fr.ay, (double)fr.az)) * (1.0 + 1e-6);
        const double rT   = (double)s.Rmin / unit, T = rT * rT;
        const double slop = (2.0 * std::sqrt(3.0) * amax * rT + 3.0 * amax * amax) * 1.01;
        const double lo   = T * (1.0 - erel) - slop, hi = T * (1.0 + erel) + slop;
        float t_in = (float)lo;
        if ((double)t_in > lo) t_in = next_down(t_in);
        float t_out = (float)hi;
        if ((double)t_out < hi) t_out = next_up(t_out);
        fr.t_in  = t_in;
        fr.t_mid = 0.5f * (t_in + t_out);
        // |r2 - t_mid| <= w must cover (t_in, t_out]: half width plus the rounding of t_mid and of the subtraction
        fr.w = next_up((float)((0.5 * ((double)t_out - (double)t_in) + 4.0 * std::ldexp(1.0, -24) * (double)t_out) * (1.0 + 1e-6)));
        const double rs = 1e-3 / unit * (1.0 + 1e-6) + std::sqrt(3.0) * amax * 1.01 + 1.0;
        fr.t_small = next_up((float)(rs * rs * (1.0 + erel)));
        for (int k = 0; k < 3; ++k) fr.L[k] = g.L[k];
    }
    static_assert(sizeof(CoordFrame) % 4 == 0, "CoordFrame is fetched word by word");
    k_fetch_words<<<1, 256, 0, c->stream>>>(reinterpret_cast<unsigned*>(a.d_cframes), reinterpret_cast<const unsigned*>(h_cframes),
                                            (int)(sizeof(CoordFrame) / 4) * c->nframes);
    CU(cudaGetLastError());
    CU(cudaEventRecord(a.ring.ev[slot], c->stream));
    CU(cudaMemsetAsync(a.d_slab_cnt, 0, sizeof(unsigned) * 2 * (size_t)c->maxK, c->stream));
    CU(cudaMemsetAsync(a.d_cnt, 0, sizeof(int) * (size_t)n0 * c->nframes, c->stream));
    if (int rc = series_reserve(c, a, (size_t)c->nframes)) return rc;
    unsigned* const tab = a.d_table + a.series_len * (size_t)s.max_count;   // this batch's slice of the per-frame log
    CU(cudaMemsetAsync(tab, 0, sizeof(unsigned) * (size_t)s.max_count * c->nframes, c->stream));

    CoordSelect sel{};
    sel.dim = s.dim; sel.use_cyl = s.use_cyl; sel.low = (double)s.low; sel.up = (double)s.up;
    sel.cx = (double)s.cx; sel.cy = (double)s.cy; sel.rlow = (double)s.Rlow; sel.rup = (double)s.Rup;
    CoordDev p{};
    p.n0 = n0; p.n1 = n1; p.rmin = (double)s.Rmin;
    // cell grid: same chooser as the RDF (cut-off = Rmin)
    RdfDev rp{};
    rp.n0 = n0; rp.n1 = n1; rp.nslab = 1; rp.Rm = (double)s.Rmin;
    // one i-cell per block: pick the register tile (1, 2 or 4 molecules per thread) whose lane utilisation x per-pair
    // overhead (fewer i per streamed j = more shared loads per pair; measured on B200) is cheapest
    int        bits = 0, reach[3] = { 0, 0, 0 }, cell_ipt = 4;
    Acc        probe; probe.cells = a.cells;
    bool       cells = false;
    {
        const int    ipts[3] = { 4, 2, 1 };
        const double over[3] = { 1.0, 1.25, 1.7 };
        double       best = 1e30;
        for (int t = 0; t < 3; ++t)
        {
            if (const char* e = getenv("B2A_COORD_IPT")) if (atoi(e) != ipts[t]) continue;
            int    b = 0, r[3];
            double cost = 0.0;
            Acc    force = probe; force.cells = probe.cells == 0 ? 0 : 1;   // get the cost even when it would not win
            if (!choose_cells(c, force, rp, false, &b, r, ipts[t] * 128, &cost)) continue;
            if (cost * over[t] < best) { best = cost * over[t]; bits = b; reach[0] = r[0]; reach[1] = r[1]; reach[2] = r[2]; cell_ipt = ipts[t]; }
        }
        int want = probe.cells;
        if (const char* e = getenv("B2A_RDF_CELLS")) want = atoi(e);
        cells = best < 1e29 && want != 0 && (want > 0 || best < 0.8);
    }
    CellDev    cd{};
    dim3       grid((n0 + 255) / 256, c->nframes);
    {
        Timed tm(c, 1);
        k5_select<<<grid, 256, 0, c->stream>>>(g0->d_cen[s.com0], c->nframes, n0, sel, a.d_slab_id, a.d_slab_cnt);
        CU(cudaGetLastError());
        if (!cells)
        {
            k3_slab_scatter<<<grid, 256, sizeof(unsigned) * 3, c->stream>>>(g0->d_fix[s.com0], a.d_slab_id, c->nframes, n0, 1, a.d_slab_cnt,
                                                                            a.d_slab_cnt + (size_t)c->maxK, a.d_sortedA);
            CU(cudaGetLastError());
        }
        else
        {
            const unsigned nk = 1u << (3 * bits);
            if (!a.d_keysA) CU(cudaMalloc(&a.d_keysA, sizeof(int) * (size_t)n0 * c->maxK));
            if (!a.d_keysB) CU(cudaMalloc(&a.d_keysB, sizeof(int) * (size_t)n1 * c->maxK));
            if (!a.d_sortedB) CU(cudaMalloc(&a.d_sortedB, sizeof(int4) * (size_t)n1 * c->maxK));
            if (a.cap_keysA < nk) { cudaFree(a.d_cellA); a.d_cellA = nullptr; CU(cudaMalloc(&a.d_cellA, sizeof(unsigned) * 3 * (size_t)(nk + 1) * c->maxK)); a.cap_keysA = nk; }
            if (a.cap_keysB < nk) { cudaFree(a.d_cellB); a.d_cellB = nullptr; CU(cudaMalloc(&a.d_cellB, sizeof(unsigned) * 3 * (size_t)(nk + 1) * c->maxK)); a.cap_keysB = nk; }
            const size_t sz = (size_t)(nk + 1) * c->maxK;
            unsigned *cntA = a.d_cellA, *stA = a.d_cellA + sz, *curA = a.d_cellA + 2 * sz;
            unsigned *cntB = a.d_cellB, *stB = a.d_cellB + sz, *curB = a.d_cellB + 2 * sz;
            CU(cudaMemsetAsync(a.d_cellA, 0, sizeof(unsigned) * 3 * sz, c->stream));
            CU(cudaMemsetAsync(a.d_cellB, 0, sizeof(unsigned) * 3 * sz, c->stream));
            dim3 gridB((n1 + 255) / 256, c->nframes);
            k3_cell_keys<<<grid, 256, 0, c->stream>>>(g0->d_fix[s.com0], a.d_slab_id, c->nframes, n0, bits, nk, a.d_keysA, cntA);
            k3_cell_keys<<<gridB, 256, 0, c->stream>>>(g1->d_fix[s.com1], nullptr, c->nframes, n1, bits, nk, a.d_keysB, cntB);
            k3_cell_scan<<<c->nframes, 1024, 0, c->stream>>>(cntA, nk, stA);
            k3_cell_scan<<<c->nframes, 1024, 0, c->stream>>>(cntB, nk, stB);
            k3_cell_scatter<<<grid, 256, 0, c->stream>>>(g0->d_fix[s.com0], a.d_keysA, c->nframes, n0, nk, stA, curA, a.d_sortedA);
            k3_cell_scatter<<<gridB, 256, 0, c->stream>>>(g1->d_fix[s.com1], a.d_keysB, c->nframes, n1, nk, stB, curB, a.d_sortedB);
            CU(cudaGetLastError());
            cd.bits = bits; cd.rx = reach[0]; cd.ry = reach[1]; cd.rz = reach[2];
            cd.nkeysA = nk; cd.nkeysB = nk; cd.startA = stA; cd.startB = stB; cd.sortedB = a.d_sortedB;
        }
    }
    {
        Timed tm(c, 6);
        if (cells)
        {
            const double per_cell = (double)n0 / (double)(1u << (3 * bits));
            const int    ny       = (int)std::min(64.0, std::max(1.0, std::ceil(1.5 * per_cell / (cell_ipt * 128))));
            const dim3   cgrid(cd.nkeysA, ny, c->nframes);
            int          rc       = 0;
            if (cell_ipt == 4) rc = launch_coord<4, 2>(c, a, *g0, *g1, p, cd, cgrid, cubic);
            else if (cell_ipt == 2) rc = launch_coord<2, 2>(c, a, *g0, *g1, p, cd, cgrid, cubic);
            else rc = launch_coord<1, 2>(c, a, *g0, *g1, p, cd, cgrid, cubic);
            if (rc) return rc;
        }
        else
        {
            constexpr int TI = 8 * 128;
            const int itiles = std::max(1, (n0 + TI - 1) / TI);
            int       jsplit = std::max(1, (c->sm_count * 24 + itiles * c->nframes - 1) / (itiles * c->nframes));
            jsplit           = std::min(jsplit, std::max(1, (n1 + kTileJ - 1) / kTileJ));
            p.jchunk         = ((n1 + jsplit - 1) / jsplit + kTileJ - 1) / kTileJ * kTileJ;
            p.jsplit         = (n1 + p.jchunk - 1) / p.jchunk;
            if (int rc = launch_coord<8, 0>(c, a, *g0, *g1, p, cd, dim3(itiles, p.jsplit, c->nframes), cubic)) return rc;
        }
        k5_hist<<<grid, 256, sizeof(unsigned) * (s.max_count + 2), c->stream>>>(a.d_cnt, a.d_slab_id, c->nframes, n0, s.max_count, tab, a.d_hist,
                                                                                 a.d_hist + a.ndev);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(a.d_series_numA + a.series_len, a.d_slab_cnt, sizeof(unsigned) * c->nframes, cudaMemcpyDeviceToDevice, c->stream));
    a.batches.push_back({ c->frame_base, c->nframes });
    a.series_len += (size_t)c->nframes;
    a.table_frames = c->nframes;
    return B2A_OK;
}

extern "C" int b2a_coord_read_frames(b2a_ctx* c, int id, uint32_t* table, uint32_t* numA)
{
    TEAM_CUR(c, b2a_coord_read_frames(m, id, table, numA));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind != ACC_COORD) return fail(B2A_ERR_ARG, "accumulator %d is not a coordination-number accumulator", id);
    if (!table || !numA) return fail(B2A_ERR_ARG, "null argument");
    if (a->table_frames <= 0 || a->series_len < (size_t)a->table_frames) return fail(B2A_ERR_STATE, "nothing accumulated yet");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(table, a->d_table + (a->series_len - a->table_frames) * (size_t)a->coord.max_count,
                       sizeof(unsigned) * (size_t)a->coord.max_count * a->table_frames, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(numA, a->d_slab_cnt, sizeof(unsigned) * a->table_frames, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2A_OK;
}

extern "C" int b2a_coord_counts(b2a_ctx* c, int id, int frame, int32_t* count)
{
    TEAM_CUR(c, b2a_coord_counts(m, id, frame, count));
    Acc* a = nullptr;
    if (int rc = get_acc(c, id, &a)) return rc;
    if (a->kind != ACC_COORD) return fail(B2A_ERR_ARG, "accumulator %d is not a coordination-number accumulator", id);
    if (!count || frame < 0 || frame >= a->table_frames) return fail(B2A_ERR_ARG, "b2a_coord_counts: bad argument");
    Group* g0 = nullptr;
    if (int rc = get_group(c, a->coord.grp0, &g0)) return rc;
    CU(cudaSetDevice(c->device));
    const size_t     n0 = (size_t)g0->nmol;
    std::vector<int> sel(n0);
    CU(cudaMemcpyAsync(count, a->d_cnt + frame * n0, sizeof(int) * n0, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(sel.data(), a->d_slab_id + frame * n0, sizeof(int) * n0, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n0; ++i)
        if (sel[i] < 0) count[i] = -1;
    return B2A_OK;
}

// host post-processing of calc_pair_dist_bulk.cpp:89-92: totPair[k] += tmpPair[k] / numA, frame by frame in double
extern "C" int b2a_coord_fold(const uint32_t* table, const uint32_t* numA, int nframes, int max_count, double* totPair)
{
    if (!table || !numA || !totPair || nframes < 0 || max_count <= 0) return fail(B2A_ERR_ARG, "b2a_coord_fold: bad argument");
    for (int f = 0; f < nframes; ++f)
    {
        if (numA[f] == 0) continue;   // tmpPair is empty: the loop at :89 does not run
        for (int k = 0; k < max_count; ++k)
        {
            const uint32_t v = table[(size_t)f * max_count + k];
            if (v) totPair[k] += (double)v / (int)numA[f];
        }
    }
    return B2A_OK;
}

#include "b2a_multi.cuh"   // multi-GPU: leader contexts, NCCL communicator, b2a_finalize, per-frame series
