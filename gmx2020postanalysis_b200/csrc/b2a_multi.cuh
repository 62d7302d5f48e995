// Multi-GPU layer of libb2a (included at the end of b2a.cu): SURVEY.md 8(e).
//
// Frames are independent units for every accumulator of the path (the reference's accumulators are sums over frames of per-frame
// integer counts: README.md:148, src/general/calc_rdf_region.cpp:91, src/mofs/calc_orient_mof.cpp:180,196), so the single frame loop
// of a program (calc_rdf_region.cpp:68-97) shards by contiguous frame blocks:
//
//   * b2a_create_multi builds a LEADER context owning one member context per device.  The caller keeps its one control thread and
//     its one loop  `submit batch -> accumulate`; b2a_submit_batch on a leader DEALS each K-frame batch (a contiguous frame block)
//     to the next device, and b2a_accumulate queues the kernels on the device that received it.  Nothing in that loop waits for a
//     kernel: uploads are double-buffered per device and ride their own copy streams, the per-frame constants travel through a
//     pinned ring (FrameRing), so one host thread keeps every device two batches deep in work; the only wait is back-pressure
//     (the upload of batch b-1 must have left the caller's pinned buffer before b2a_submit_batch(b) returns).
//   * every device accumulates into PRIVATE u64 histograms; b2a_finalize merges them with ONE ncclAllReduce(ncclUint64, ncclSum) per
//     accumulator over [hist, stats) -- grouped, in place on each device's own stream, NVLink / NVSwitch underneath -- into a
//     second buffer (d_merged), so the private counts stay intact, accumulation may continue, and every device (and every process
//     of a multi-process communicator, b2a_comm_init) reads the same totals.  Integer sums: bit-identical for any device count.
//   * per-frame outputs (calc_density_time.cpp:41-58 prints one line per frame; calc_pair_dist_bulk.cpp:89-92 normalises per frame)
//     are logged on the device that computed them together with the global index of their frames and gathered in frame order
//     (b2a_series_read; across processes with ncclAllGather inside b2a_finalize).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): single-GPU users of libb2a do not need it, and a host process that already
// loaded a libnccl (torch) shares it.
#pragma once

#include <dlfcn.h>

namespace
{
// ---- the handful of NCCL entry points used, resolved on first use -----------------------------------------------
struct NcclApi
{
    typedef struct { char internal[128]; } UniqueId;   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
    enum { Sum = 0, Int64 = 4, Uint64 = 5 };          // ncclRedOp_t / ncclDataType_t values (stable across NCCL 2.x)
    int (*GetUniqueId)(UniqueId*)                                                         = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int)                                       = nullptr;
    int (*CommInitAll)(void**, int, const int*)                                           = nullptr;
    int (*CommDestroy)(void*)                                                             = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t)           = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t)                = nullptr;
    int (*GroupStart)()                                                                   = nullptr;
    int (*GroupEnd)()                                                                     = nullptr;
    const char* (*GetErrorString)(int)                                                    = nullptr;
    void* handle = nullptr;
    bool  ok     = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names)
        if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.handle) return api;
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll    = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce      = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather      = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart     = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd       = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce && api.AllGather &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api;
}

#define NC(call)                                                                                                     \
    do {                                                                                                             \
        int r_ = (call);                                                                                             \
        if (r_ != 0) return fail(B2A_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl().GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

int need_nccl()
{
    if (!nccl().ok) return fail(B2A_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    return B2A_OK;
}

// degenerate team whose members share one device (the one-GPU test box): NCCL refuses two ranks on a device, so the merge is a
// local sum.  Real teams (distinct devices) never come here.
__global__ void k_sum_words(unsigned long long* dst, const unsigned long long* const* src, int nsrc, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        unsigned long long v = 0;
        for (int s = 0; s < nsrc; ++s) v += src[s][i];
        dst[i] = v;
    }
}

std::vector<b2a_ctx*> devices_of(b2a_ctx* c)
{
    if (c->is_leader()) return c->members;
    return { c };
}
} // namespace

static void multi_release(b2a_ctx* c)
{
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    c->comm = nullptr;
}

// ---- contexts ------------------------------------------------------------------------------------------------------
extern "C" int b2a_create_multi(b2a_ctx** out, int ndev, const int* devices, int natoms, int max_frames)
{
    if (!out || ndev <= 0) return fail(B2A_ERR_ARG, "b2a_create_multi: ndev must be positive");
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0)
    {
        cudaGetLastError();
        return fail(B2A_ERR_CUDA, "b2a_create_multi: no CUDA device visible; libb2a has no CPU fallback");
    }
    struct Undo { void operator()(b2a_ctx* p) const { b2a_destroy(p); } };
    std::unique_ptr<b2a_ctx, Undo> lead(new b2a_ctx);
    lead->natoms = natoms; lead->maxK = max_frames; lead->device = devices ? devices[0] : 0;
    std::vector<int> devs(ndev);
    bool             distinct = true;
    for (int i = 0; i < ndev; ++i)
    {
        devs[i] = devices ? devices[i] : i;
        for (int j = 0; j < i; ++j) if (devs[j] == devs[i]) distinct = false;
    }
    for (int i = 0; i < ndev; ++i)
    {
        b2a_ctx* m = nullptr;
        if (int rc = b2a_create(&m, devs[i], natoms, max_frames)) return rc;
        lead->members.push_back(m);
    }
    lead->sm_count = lead->members[0]->sm_count;
    if (ndev > 1 && distinct)
    {
        // one communicator over the devices of this process, created up front: the frame loop never pays for it
        if (int rc = need_nccl()) return rc;
        std::vector<void*> comms(ndev, nullptr);
        NC(nccl().CommInitAll(comms.data(), ndev, devs.data()));
        for (int i = 0; i < ndev; ++i)
        {
            lead->members[i]->comm = comms[i];
            lead->members[i]->comm_rank = i;
            lead->members[i]->comm_size = ndev;
        }
    }
    else if (ndev > 1)
        lead->same_device_team = true;
    lead->comm_size = ndev;
    *out = lead.release();
    return B2A_OK;
}

extern "C" int b2a_num_devices(b2a_ctx* c) { return !c ? 0 : (c->is_leader() ? (int)c->members.size() : 1); }

extern "C" b2a_ctx* b2a_member(b2a_ctx* c, int i)
{
    if (!c) return nullptr;
    if (!c->is_leader()) return i == 0 ? c : nullptr;
    return (i >= 0 && i < (int)c->members.size()) ? c->members[i] : nullptr;
}

extern "C" int b2a_current_device(b2a_ctx* c)
{
    if (!c) return -1;
    if (!c->is_leader()) return 0;
    return c->cur_member;
}

extern "C" int b2a_set_frame_base(b2a_ctx* c, long long first_frame)
{
    if (!c || first_frame < 0) return fail(B2A_ERR_ARG, "b2a_set_frame_base: bad argument");
    c->next_frame = first_frame;
    return B2A_OK;
}

// ---- multi-process communicator ----------------------------------------------------------------------------------------
extern "C" int b2a_comm_unique_id(void* id128)
{
    if (!id128) return fail(B2A_ERR_ARG, "b2a_comm_unique_id: null argument");
    if (int rc = need_nccl()) return rc;
    NcclApi::UniqueId id;
    NC(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return B2A_OK;
}

extern "C" int b2a_comm_init(b2a_ctx* c, int nprocs, int proc_rank, const void* id128)
{
    if (!c || !id128 || nprocs <= 0 || proc_rank < 0 || proc_rank >= nprocs) return fail(B2A_ERR_ARG, "b2a_comm_init: bad argument");
    if (c->same_device_team) return fail(B2A_ERR_UNSUPPORTED, "b2a_comm_init: the members of this context share a device");
    if (int rc = need_nccl()) return rc;
    std::vector<b2a_ctx*> devs = devices_of(c);
    const int             nloc = (int)devs.size();
    NcclApi::UniqueId     id;
    std::memcpy(&id, id128, sizeof(id));
    for (b2a_ctx* m : devs) multi_release(m);
    // every device of every process is one rank: rank = proc_rank * devices_per_process + local index
    NC(nccl().GroupStart());
    for (int i = 0; i < nloc; ++i)
    {
        CU(cudaSetDevice(devs[i]->device));
        NC(nccl().CommInitRank(&devs[i]->comm, nprocs * nloc, id, proc_rank * nloc + i));
    }
    NC(nccl().GroupEnd());
    for (int i = 0; i < nloc; ++i)
    {
        devs[i]->comm_rank = proc_rank * nloc + i;
        devs[i]->comm_size = nprocs * nloc;
        devs[i]->comm_procs = nprocs;
    }
    c->comm_procs = nprocs;
    c->comm_size  = nprocs * nloc;
    return B2A_OK;
}

// ---- dealing batches to devices -------------------------------------------------------------------------------------------
static int team_submit(b2a_ctx* c, int nframes, const float* x, const float* box9)
{
    const int next = (c->cur_member + 1) % (int)c->members.size();
    b2a_ctx*  m    = c->members[next];
    m->next_frame  = c->next_frame;           // global frame index of this block
    if (int rc = b2a_submit_batch(m, nframes, x, box9)) return rc;
    c->cur_member = next;
    c->frame_base = c->next_frame;
    c->next_frame += nframes;
    c->nframes = nframes;
    c->serial++;
    // The caller cycles through `host_buffers` staging buffers: before this call returns, the upload issued host_buffers - 1
    // submits ago must have left its buffer (with the default of two that is the previous one: the contract of a single device).
    // With one buffer more than devices the host never waits for a copy, and the uploads of different devices overlap.
    c->inflight.push_back({ next, m->cur });
    while ((int)c->inflight.size() > std::max(1, c->host_buffers - 1))
    {
        b2a_ctx* o = c->members[c->inflight.front().first];
        CU(cudaSetDevice(o->device));
        CU(cudaEventSynchronize(o->ev_copied[c->inflight.front().second]));   // (re-recorded by a later upload at worst: still safe)
        if (o->field_pending) { CU(cudaEventSynchronize(o->ev_field)); o->field_pending = false; }
        c->inflight.erase(c->inflight.begin());
    }
    return B2A_OK;
}

extern "C" int b2a_set_host_buffers(b2a_ctx* c, int n)
{
    if (!c || n < 2) return fail(B2A_ERR_ARG, "b2a_set_host_buffers: at least two buffers");
    c->host_buffers = n;
    return B2A_OK;
}

// ---- the merge -------------------------------------------------------------------------------------------------------------
namespace
{
// local (this process) per-frame series of one accumulator, merged over the devices and ordered by global frame index:
// words per frame = series_words(a); coordination tables are widened to u64 and followed by numA
size_t series_words(const Acc& a) { return a.kind == ACC_REGION ? a.ndev : (size_t)a.coord.max_count + 1; }
bool   has_series(const Acc& a) { return (a.kind == ACC_REGION && a.region.per_frame) || a.kind == ACC_COORD; }

int local_series(const std::vector<b2a_ctx*>& devs, int id, std::vector<long long>& idx, std::vector<unsigned long long>& data)
{
    struct Block { long long first; int n; int dev; size_t off; };
    std::vector<Block>                          blocks;
    std::vector<std::vector<unsigned long long>> raw(devs.size());
    size_t                                      W = 0, total = 0;
    for (size_t d = 0; d < devs.size(); ++d)
    {
        b2a_ctx* m = devs[d];
        Acc*     a = m->accs[id].get();
        W          = series_words(*a);
        CU(cudaSetDevice(m->device));
        raw[d].resize(a->series_len * W);
        if (a->series_len)
        {
            if (a->kind == ACC_REGION)
                CU(cudaMemcpyAsync(raw[d].data(), a->d_perframe, sizeof(unsigned long long) * a->series_len * W, cudaMemcpyDeviceToHost, m->stream));
            else
            {
                const size_t          mc = (size_t)a->coord.max_count;
                std::vector<unsigned> tab(a->series_len * mc), numA(a->series_len);
                CU(cudaMemcpyAsync(tab.data(), a->d_table, sizeof(unsigned) * tab.size(), cudaMemcpyDeviceToHost, m->stream));
                CU(cudaMemcpyAsync(numA.data(), a->d_series_numA, sizeof(unsigned) * numA.size(), cudaMemcpyDeviceToHost, m->stream));
                CU(cudaStreamSynchronize(m->stream));
                for (size_t f = 0; f < a->series_len; ++f)
                {
                    for (size_t k = 0; k < mc; ++k) raw[d][f * W + k] = tab[f * mc + k];
                    raw[d][f * W + mc] = numA[f];
                }
            }
            CU(cudaStreamSynchronize(m->stream));
        }
        size_t off = 0;
        for (auto& b : a->batches) { blocks.push_back({ b.first, b.second, (int)d, off }); off += (size_t)b.second; }
        total += a->series_len;
    }
    std::stable_sort(blocks.begin(), blocks.end(), [](const Block& p, const Block& q) { return p.first < q.first; });
    idx.clear(); data.clear();
    idx.reserve(total); data.reserve(total * W);
    for (const Block& b : blocks)
        for (int k = 0; k < b.n; ++k)
        {
            idx.push_back(b.first + k);
            const unsigned long long* src = raw[b.dev].data() + (b.off + k) * W;
            data.insert(data.end(), src, src + W);
        }
    return B2A_OK;
}

bool any_dirty(const std::vector<b2a_ctx*>& devs, int id)
{
    for (b2a_ctx* m : devs)
    {
        const Acc* a = m->accs[id].get();
        if (!a->d_merged || a->merged_epoch != a->epoch) return true;
    }
    return false;
}
} // namespace

extern "C" int b2a_finalize(b2a_ctx* c)
{
    if (!c) return fail(B2A_ERR_ARG, "null context");
    std::vector<b2a_ctx*> devs = devices_of(c);
    const size_t          nacc = devs[0]->accs.size();
    for (b2a_ctx* m : devs)
        if (m->accs.size() != nacc) return fail(B2A_ERR_STATE, "b2a_finalize: the devices hold different accumulators");
    const bool collective = devs[0]->comm != nullptr;
    if (!collective && !c->same_device_team && devs.size() > 1) return fail(B2A_ERR_STATE, "b2a_finalize: no communicator");
    for (b2a_ctx* m : devs)
    {   // accumulators with work on side streams: the merge is ordered behind them
        CU(cudaSetDevice(m->device));
        if (int rc = join_side(m)) return rc;
    }
    for (b2a_ctx* m : devs)
        for (auto& a : m->accs)
            if (a && !a->d_merged)
            {
                CU(cudaSetDevice(m->device));
                CU(cudaMalloc(&a->d_merged, sizeof(unsigned long long) * (a->ndev + B2A_STAT_NSLOTS)));
            }
    if (collective)
    {
        // ONE all-reduce per accumulator over [hist, stats): 64-bit integer sum, grouped over the accumulators and the local
        // devices, each on its device's own stream behind the kernels that produced the counts
        NC(nccl().GroupStart());
        for (size_t id = 0; id < nacc; ++id)
            for (b2a_ctx* m : devs)
            {
                Acc* a = m->accs[id].get();
                if (!a) continue;
                NC(nccl().AllReduce(a->d_hist, a->d_merged, a->ndev + B2A_STAT_NSLOTS, NcclApi::Uint64, NcclApi::Sum, m->comm, m->stream));
            }
        NC(nccl().GroupEnd());
    }
    else if (devs.size() > 1)
    {
        // members on ONE device (test configuration): sum locally on member 0's stream, then hand every member the totals
        b2a_ctx* m0 = devs[0];
        CU(cudaSetDevice(m0->device));
        for (size_t d = 1; d < devs.size(); ++d) CU(cudaStreamSynchronize(devs[d]->stream));
        for (size_t id = 0; id < nacc; ++id)
        {
            Acc* a0 = m0->accs[id].get();
            if (!a0) continue;
            const size_t                          n = a0->ndev + B2A_STAT_NSLOTS;
            std::vector<const unsigned long long*> src;
            for (b2a_ctx* m : devs) src.push_back(m->accs[id]->d_hist);
            const unsigned long long** d_src = nullptr;
            CU(cudaMalloc(&d_src, sizeof(void*) * src.size()));
            CU(cudaMemcpyAsync(d_src, src.data(), sizeof(void*) * src.size(), cudaMemcpyHostToDevice, m0->stream));
            k_sum_words<<<(unsigned)std::min<size_t>(1024, (n + 255) / 256), 256, 0, m0->stream>>>(a0->d_merged, d_src, (int)src.size(), n);
            CU(cudaGetLastError());
            for (size_t d = 1; d < devs.size(); ++d)
                CU(cudaMemcpyAsync(devs[d]->accs[id]->d_merged, a0->d_merged, sizeof(unsigned long long) * n, cudaMemcpyDeviceToDevice, m0->stream));
            CU(cudaStreamSynchronize(m0->stream));
            cudaFree(d_src);
        }
    }
    else
    {
        Acc* a = nullptr;
        for (size_t id = 0; id < nacc; ++id)
            if ((a = devs[0]->accs[id].get()))
                CU(cudaMemcpyAsync(a->d_merged, a->d_hist, sizeof(unsigned long long) * (a->ndev + B2A_STAT_NSLOTS), cudaMemcpyDeviceToDevice, devs[0]->stream));
    }
    for (b2a_ctx* m : devs)
        for (auto& a : m->accs)
            if (a) a->merged_epoch = a->epoch;

    // per-frame series across PROCESSES: every process contributes its locally merged series through its first device
    if (collective && devs[0]->comm_procs > 1)
    {
        for (size_t id = 0; id < nacc; ++id)
        {
            Acc* a0 = devs[0]->accs[id].get();
            if (!a0 || !has_series(*a0)) continue;
            std::vector<long long>          idx;
            std::vector<unsigned long long> data;
            if (int rc = local_series(devs, (int)id, idx, data)) return rc;
            const size_t W = series_words(*a0), nranks = (size_t)devs[0]->comm_size;
            // 1. how many frames does every rank bring (devices other than a process's first bring none)
            std::vector<unsigned long long*> d_cnt(devs.size(), nullptr);
            std::vector<std::vector<unsigned long long>> h_cnt(devs.size(), std::vector<unsigned long long>(nranks));
            for (size_t d = 0; d < devs.size(); ++d)
            {
                CU(cudaSetDevice(devs[d]->device));
                CU(cudaMalloc(&d_cnt[d], sizeof(unsigned long long) * (nranks + 1)));
                const unsigned long long mine = d == 0 ? idx.size() : 0;
                CU(cudaMemcpyAsync(d_cnt[d] + nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice, devs[d]->stream));
                CU(cudaStreamSynchronize(devs[d]->stream));   // `mine` is a stack variable
            }
            NC(nccl().GroupStart());
            for (size_t d = 0; d < devs.size(); ++d)
                NC(nccl().AllGather(d_cnt[d] + nranks, d_cnt[d], 1, NcclApi::Uint64, devs[d]->comm, devs[d]->stream));
            NC(nccl().GroupEnd());
            for (size_t d = 0; d < devs.size(); ++d)
            {
                CU(cudaSetDevice(devs[d]->device));
                CU(cudaMemcpyAsync(h_cnt[d].data(), d_cnt[d], sizeof(unsigned long long) * nranks, cudaMemcpyDeviceToHost, devs[d]->stream));
                CU(cudaStreamSynchronize(devs[d]->stream));
                cudaFree(d_cnt[d]);
            }
            size_t maxn = 0;
            for (unsigned long long v : h_cnt[0]) maxn = std::max<size_t>(maxn, v);
            a0->gathered_idx.clear(); a0->gathered.clear();
            if (maxn)
            {
                // 2. padded all-gather of (frame index, payload) records
                const size_t rec = W + 1;
                std::vector<unsigned long long*> d_send(devs.size(), nullptr), d_recv(devs.size(), nullptr);
                std::vector<unsigned long long>  send(maxn * rec, 0);
                for (size_t f = 0; f < idx.size(); ++f)
                {
                    send[f * rec] = (unsigned long long)idx[f];
                    std::memcpy(&send[f * rec + 1], &data[f * W], sizeof(unsigned long long) * W);
                }
                for (size_t d = 0; d < devs.size(); ++d)
                {
                    CU(cudaSetDevice(devs[d]->device));
                    CU(cudaMalloc(&d_send[d], sizeof(unsigned long long) * maxn * rec));
                    CU(cudaMalloc(&d_recv[d], sizeof(unsigned long long) * maxn * rec * nranks));
                    if (d == 0) CU(cudaMemcpyAsync(d_send[d], send.data(), sizeof(unsigned long long) * send.size(), cudaMemcpyHostToDevice, devs[d]->stream));
                    else CU(cudaMemsetAsync(d_send[d], 0, sizeof(unsigned long long) * maxn * rec, devs[d]->stream));
                }
                NC(nccl().GroupStart());
                for (size_t d = 0; d < devs.size(); ++d)
                    NC(nccl().AllGather(d_send[d], d_recv[d], maxn * rec, NcclApi::Uint64, devs[d]->comm, devs[d]->stream));
                NC(nccl().GroupEnd());
                std::vector<unsigned long long> recv(maxn * rec * nranks);
                CU(cudaSetDevice(devs[0]->device));
                CU(cudaMemcpyAsync(recv.data(), d_recv[0], sizeof(unsigned long long) * recv.size(), cudaMemcpyDeviceToHost, devs[0]->stream));
                for (size_t d = 0; d < devs.size(); ++d)
                {
                    CU(cudaSetDevice(devs[d]->device));
                    CU(cudaStreamSynchronize(devs[d]->stream));
                    cudaFree(d_send[d]); cudaFree(d_recv[d]);
                }
                // 3. order by frame index
                std::vector<std::pair<long long, const unsigned long long*>> recs;
                for (size_t r = 0; r < nranks; ++r)
                    for (size_t f = 0; f < h_cnt[0][r]; ++f)
                    {
                        const unsigned long long* p = &recv[(r * maxn + f) * rec];
                        recs.push_back({ (long long)p[0], p + 1 });
                    }
                std::stable_sort(recs.begin(), recs.end(), [](const auto& p, const auto& q) { return p.first < q.first; });
                for (auto& r : recs)
                {
                    a0->gathered_idx.push_back(r.first);
                    a0->gathered.insert(a0->gathered.end(), r.second, r.second + W);
                }
            }
            a0->gathered_epoch = a0->epoch;
        }
    }
    return B2A_OK;
}

static int team_read(b2a_ctx* c, int id, uint64_t* counts, uint64_t* stats)
{
    b2a_ctx* m0 = c->members[0];
    if (id < 0 || id >= (int)m0->accs.size() || !m0->accs[id]) return fail(B2A_ERR_ARG, "unknown accumulator %d", id);
    // One process owns every device: merge on demand.  (With a multi-process communicator the merge is a collective every
    // process must enter, so there the caller calls b2a_finalize itself and a stale read returns this process's share.)
    if (any_dirty(c->members, id) && m0->comm_procs == 1)
        if (int rc = b2a_finalize(c)) return rc;
    if (any_dirty(c->members, id))
    {
        // not finalized across processes: this process's own devices, summed on the host
        Acc*                  a0 = m0->accs[id].get();
        std::vector<uint64_t> tmp(a0->ncounters), st(B2A_STAT_NSLOTS);
        for (size_t i = 0; i < a0->ncounters; ++i) counts[i] = 0;
        if (stats) for (int i = 0; i < B2A_STAT_NSLOTS; ++i) stats[i] = 0;
        for (b2a_ctx* m : c->members)
        {
            if (int rc = b2a_acc_read(m, id, tmp.data(), st.data())) return rc;
            for (size_t i = 0; i < a0->ncounters; ++i) counts[i] += tmp[i];
            if (stats) for (int i = 0; i < B2A_STAT_NSLOTS; ++i) stats[i] += st[i];
        }
        return B2A_OK;
    }
    return b2a_acc_read(m0, id, counts, stats);
}

// ---- per-frame series --------------------------------------------------------------------------------------------------------
extern "C" int b2a_series_size(b2a_ctx* c, int id, size_t* nframes, size_t* words_per_frame)
{
    if (!c || !nframes || !words_per_frame) return fail(B2A_ERR_ARG, "b2a_series_size: null argument");
    std::vector<b2a_ctx*> devs = devices_of(c);
    if (id < 0 || id >= (int)devs[0]->accs.size() || !devs[0]->accs[id]) return fail(B2A_ERR_ARG, "unknown accumulator %d", id);
    Acc* a0 = devs[0]->accs[id].get();
    if (!has_series(*a0)) return fail(B2A_ERR_ARG, "accumulator %d keeps no per-frame series", id);
    *words_per_frame = series_words(*a0);
    if (a0->gathered_epoch == a0->epoch && devs[0]->comm_procs > 1) { *nframes = a0->gathered_idx.size(); return B2A_OK; }
    size_t n = 0;
    for (b2a_ctx* m : devs) n += m->accs[id]->series_len;
    *nframes = n;
    return B2A_OK;
}

extern "C" int b2a_series_read(b2a_ctx* c, int id, long long* frame_index, uint64_t* out)
{
    if (!c || !out) return fail(B2A_ERR_ARG, "b2a_series_read: null argument");
    std::vector<b2a_ctx*> devs = devices_of(c);
    if (id < 0 || id >= (int)devs[0]->accs.size() || !devs[0]->accs[id]) return fail(B2A_ERR_ARG, "unknown accumulator %d", id);
    Acc* a0 = devs[0]->accs[id].get();
    if (!has_series(*a0)) return fail(B2A_ERR_ARG, "accumulator %d keeps no per-frame series", id);
    if (a0->gathered_epoch == a0->epoch && devs[0]->comm_procs > 1)
    {
        if (frame_index) std::memcpy(frame_index, a0->gathered_idx.data(), sizeof(long long) * a0->gathered_idx.size());
        std::memcpy(out, a0->gathered.data(), sizeof(unsigned long long) * a0->gathered.size());
        return B2A_OK;
    }
    std::vector<long long>          idx;
    std::vector<unsigned long long> data;
    if (int rc = local_series(devs, id, idx, data)) return rc;
    if (frame_index) std::memcpy(frame_index, idx.data(), sizeof(long long) * idx.size());
    std::memcpy(out, data.data(), sizeof(unsigned long long) * data.size());
    return B2A_OK;
}
