// FrameBatcher -- shared by the drop-in handles: decodes K frames at a time into a pinned host batch, ships them to
// the GPU with one asynchronous copy (b2a_submit_batch) and keeps a t_trxframe pointing at the current frame.
// It is the piece that lets the per-frame API (readFirstFrame / readNextFrame / hd.fr->x) survive on top of batches
// (SURVEY.md H4).
#pragma once

#include <b2a.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace itp
{
namespace b2a_detail
{
    inline void check(int rc)
    {
        if (rc != B2A_OK) gmx_fatal(FARGS, "%s", b2a_last_error());   // the reference aborts through gmx_fatal too
    }

    class FrameBatcher
    {
    public:
        FrameBatcher()
        {
            if (const char* e = std::getenv("B2A_BATCH")) k_ = std::max(1, std::atoi(e));
        }
        ~FrameBatcher()
        {
            if (std::getenv("B2A_TRACE"))   // where the control thread spent its time (one line on stderr)
                std::fprintf(stderr, "b2a batcher: %ld batches, host seconds: read/decode %.4f, submit %.4f\n", serial_, t_read_, t_submit_);
            if (ctx_) b2a_destroy(ctx_);
            for (size_t p = 0; p < px_.size(); ++p)
            {
                if (px_[p]) b2a_pinned_free(px_[p]);
                if (pv_[p]) b2a_pinned_free(pv_[p]);
                if (pf_[p]) b2a_pinned_free(pf_[p]);
            }
        }
        FrameBatcher(const FrameBatcher&) = delete;
        FrameBatcher& operator=(const FrameBatcher&) = delete;

        b2a_ctx* ctx(int natoms)
        {
            if (!ctx_)
            {
                // B2A_DEVICES=0,1,2,3 (or "all"): one context over several GPUs -- batches are dealt to the devices in turn and the
                // histograms merged by one NCCL all-reduce when they are read (b2a_create_multi); B2A_DEVICE=n: that one device
                std::vector<int> devs;
                if (const char* e = std::getenv("B2A_DEVICES"))
                {
                    if (std::strcmp(e, "all") == 0)
                        for (int d = 0; d < b2a_device_count(); ++d) devs.push_back(d);
                    else
                        for (const char* p = e; *p;)
                        {
                            char* q = nullptr;
                            const long v = std::strtol(p, &q, 10);
                            if (q == p) break;
                            devs.push_back((int)v);
                            p = (*q == ',') ? q + 1 : q;
                        }
                }
                if (devs.size() > 1) check(b2a_create_multi(&ctx_, (int)devs.size(), devs.data(), natoms, k_));
                else
                {
                    int dev = devs.empty() ? 0 : devs[0];
                    if (const char* e = std::getenv("B2A_DEVICE")) dev = std::atoi(e);
                    check(b2a_create(&ctx_, dev, natoms, k_));
                }
            }
            return ctx_;
        }
        b2a_ctx* ctx() const { return ctx_; }
        int      frames() const { return n_; }
        float    time_of(int k) const { return btime_[k]; }   // fr->time of frame k of the resident batch
        // fr->time of every frame handed to the device so far, by global frame index (per-frame series are printed after the loop)
        const std::vector<float>& all_times() const { return times_all_; }
        int      current() const { return cur_; }
        int      capacity() const { return k_; }
        long     serial() const { return serial_; }

        // (re)start: `first` was just filled by read_first_frame
        bool start(gmx_output_env_t* oenv, t_trxstatus* status, const t_trxframe& first, t_trxframe* fr)
        {
            oenv_ = oenv; status_ = status; io_ = first; eof_ = false;
            times_all_.clear();
            if (ctx_) check(b2a_set_frame_base(ctx_, 0));   // a restarted trajectory counts its frames from 0 again
            natoms_ = io_.natoms;
            const size_t n3 = (size_t)natoms_ * 3;
            // staging buffers: two for one device (ping-pong); one more than devices otherwise, so that no submit waits for a copy
            const int ndev = b2a_num_devices(ctx(natoms_));
            if (px_.empty())
            {
                const size_t nb = ndev > 1 ? (size_t)ndev + 1 : 2;
                px_.assign(nb, nullptr); pv_.assign(nb, nullptr); pf_.assign(nb, nullptr);
                if (ndev > 1) check(b2a_set_host_buffers(ctx_, (int)nb));
            }
            for (size_t p = 0; p < px_.size(); ++p)
            {
                if (!px_[p]) px_[p] = static_cast<float*>(b2a_pinned_alloc(n3 * sizeof(float) * k_));
                if (io_.bV && !pv_[p]) pv_[p] = static_cast<float*>(b2a_pinned_alloc(n3 * sizeof(float) * k_));
                if (io_.bF && !pf_[p]) pf_[p] = static_cast<float*>(b2a_pinned_alloc(n3 * sizeof(float) * k_));
                if (!px_[p] || (io_.bV && !pv_[p]) || (io_.bF && !pf_[p])) gmx_fatal(FARGS, "%s", b2a_last_error());
            }
            bbox_.resize(9 * (size_t)k_);
            btime_.resize(k_);
            if (!fill(true)) return false;
            point(0, fr);
            return true;
        }

        // advance one frame; refills the batch when it is exhausted
        bool next(t_trxframe* fr)
        {
            if (cur_ + 1 < n_) { point(cur_ + 1, fr); return true; }
            if (eof_ || !fill(false)) return false;
            point(0, fr);
            return true;
        }

        // number of frames skipped when the caller consumes the rest of the batch as a block
        int skip_to_batch_end()
        {
            const int skipped = n_ - 1 - cur_;
            cur_ = n_ - 1;
            return skipped;
        }

    private:
        bool fill(bool first)
        {
            const size_t n3 = (size_t)natoms_ * 3;
            int          n  = 0;
            flip_ = (flip_ + 1) % (int)px_.size();
            bx_ = px_[flip_]; bv_ = pv_[flip_]; bf_ = pf_[flip_];
#ifdef GMXSTUB_BATCH_READ
            // The libgromacs stand-in can hand out K frames per call and decodes XTC frames on K host threads, straight
            // into the pinned batch.  With the real libgromacs this block compiles away and the loop below runs.
            if (io_.bX && !io_.bV && !io_.bF && !std::getenv("B2A_NO_BATCH_READ"))
            {
                if (first)
                {
                    std::memcpy(bx_, io_.x, n3 * sizeof(float));
                    std::memcpy(bbox_.data(), &io_.box[0][0], 9 * sizeof(float));
                    btime_[0] = io_.time;
                    n         = 1;
                }
                const auto t0 = std::chrono::steady_clock::now();
                if (!eof_ && n < k_)
                {
                    const int got = gmxstub_read_frames(status_, k_ - n, bx_ + (size_t)n * n3, bbox_.data() + 9 * (size_t)n, btime_.data() + n);
                    if (got < k_ - n) eof_ = true;
                    n += got;
                }
                if (n == 0) return false;
                n_ = n;
                ++serial_;
                times_all_.insert(times_all_.end(), btime_.begin(), btime_.begin() + n);
                const auto t1 = std::chrono::steady_clock::now();
                check(b2a_submit_batch(ctx(natoms_), n, bx_, bbox_.data()));
                t_read_ += std::chrono::duration<double>(t1 - t0).count();
                t_submit_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
                return true;
            }
#endif
            for (; n < k_; ++n)
            {
                if (!(first && n == 0))
                    if (eof_ || !read_next_frame(oenv_, status_, &io_)) { eof_ = true; break; }
                if (io_.bX) std::memcpy(bx_ + (size_t)n * n3, io_.x, n3 * sizeof(float));
                if (io_.bV) std::memcpy(bv_ + (size_t)n * n3, io_.v, n3 * sizeof(float));
                if (io_.bF) std::memcpy(bf_ + (size_t)n * n3, io_.f, n3 * sizeof(float));
                std::memcpy(bbox_.data() + 9 * (size_t)n, &io_.box[0][0], 9 * sizeof(float));
                btime_[n] = io_.time;
            }
            if (n == 0) return false;
            n_ = n;
            ++serial_;
            times_all_.insert(times_all_.end(), btime_.begin(), btime_.begin() + n);
            if (io_.bX) check(b2a_submit_batch(ctx(natoms_), n, bx_, bbox_.data()));
            // velocities / forces (TRR) ride along so that loadVelocity[Center] / loadForce[Center] run on the device too
            if (io_.bX && io_.bV) check(b2a_submit_field(ctx(natoms_), B2A_FIELD_V, n, bv_));
            if (io_.bX && io_.bF) check(b2a_submit_field(ctx(natoms_), B2A_FIELD_F, n, bf_));
            return true;
        }

        void point(int k, t_trxframe* fr)
        {
            const size_t n3 = (size_t)natoms_ * 3;
            *fr        = io_;
            fr->x      = io_.bX ? reinterpret_cast<rvec*>(bx_ + (size_t)k * n3) : nullptr;
            fr->v      = io_.bV ? reinterpret_cast<rvec*>(bv_ + (size_t)k * n3) : nullptr;
            fr->f      = io_.bF ? reinterpret_cast<rvec*>(bf_ + (size_t)k * n3) : nullptr;
            fr->time   = btime_[k];
            fr->natoms = natoms_;
            std::memcpy(&fr->box[0][0], bbox_.data() + 9 * (size_t)k, 9 * sizeof(float));
            cur_ = k;
        }

        b2a_ctx*           ctx_ = nullptr;
        gmx_output_env_t*  oenv_ = nullptr;
        t_trxstatus*       status_ = nullptr;
        t_trxframe         io_{};
        int                natoms_ = 0, k_ = 16, n_ = 0, cur_ = 0;
        bool               eof_ = false;
        long               serial_ = 0;
        // Pinned staging buffers, used in turn (two for one device; devices + 1 for a multi-device context, b2a_set_host_buffers): the asynchronous H2D copy of batch k may still be reading its
        // buffer while the host already decodes batch k+1 into the other one; b2a_submit_batch(k+1) waits for the upload of batch k (and for pending field uploads)
        // before it returns, so by the time buffer k%2 is refilled (batch k+2) its copy has completed.
        std::vector<float*> px_, pv_, pf_;
        float*             bx_ = nullptr, *bv_ = nullptr, *bf_ = nullptr;   // the buffers of the batch being filled / served
        int                flip_ = 0;
        std::vector<float> bbox_, btime_, times_all_;
        double             t_read_ = 0, t_submit_ = 0;
    };
} // namespace b2a_detail
} // namespace itp
