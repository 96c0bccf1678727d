// Host pipeline, CPU side: worker threads gather the few bytes per env the step needs out of the simulator's two sparse AoS
// tensors (rigid_body: 40 B of 52 * num_bodies per env; net_contact: 2 x 12 B of 12 * num_bodies) and out of root_states (28 of
// 104 B: robot position, ball xy position / velocity) into ONE compact PINNED record per env (pack_layout: 96 B for BezKick),
// so that one dense cudaMemcpyAsync per chunk moves them at the link's streaming rate.  The copy engine that serves the H2D direction is row-rate-bound on strided pulls (1.4 ns per row,
// profiles/r02_host_link.md): 786 432 rows per 262 144-env step cost it 1.4 ms, the same bytes packed densely 0.35 ms, and
// the host cores that would otherwise idle between simulate() calls do the gather while the engine moves the dense tensors.
//
// No CUDA call in this file: plain host threads (no allocation per call; a fixed ring of job slots; workers block on a
// condition variable between rollouts, and poll for a bounded time between the steps of one).
#include "bezk_internal.h"
#include <atomic>
#include <chrono>
#include <pthread.h>
#include <sched.h>
#include <climits>
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#define BEZK_CPU_RELAX() _mm_pause()
#else
#define BEZK_CPU_RELAX() do {} while (0)
#endif

namespace bezk {
namespace {

constexpr int kSlots = 64;            // outstanding jobs (a step issues one per chunk)
constexpr int64_t kPiece = 1024;      // envs a worker claims at a time
constexpr int kPrefetchAhead = 12;    // envs; the AoS rows are 1144 / 264 B apart, one or two cache lines each

struct Job {
    const float* rb; const float* cf; const float* root; float* rec;        // rec: record of env env0 (the job's own destination)
    const float* dof; float* dof_dst;                                       // optional: dense dof_state rows copied along
    const float* values;                                                    // optional: (N,) critic values -> the record's last float
    int64_t rb_stride, rb_off, cf_stride, cf_l, cf_r, root_stride;       // floats
    PackLayout L;
    int64_t env0, n;
    std::atomic<int64_t> next{0};        // next piece to claim (env offset inside the job)
    std::atomic<int64_t> done{0};        // envs packed
    std::atomic<int64_t> ticket{0};      // 0 = slot free / finished
};

// One record: gathered into a local array, then written with non-temporal 16-byte stores (the record is only ever read by the
// copy engine: no read-for-ownership of the destination line, no cache pollution).  STRIDE == 0: run-time stride.
template <int STRIDE>
inline void pack_one(const Job& j, const PackLayout& L, const float* rb, const float* cl, const float* cr, int64_t e) {
    const int stride = STRIDE ? STRIDE : L.stride;
    alignas(16) float t[STRIDE ? STRIDE : 48];
    const int fw = STRIDE == 24 ? 3 : L.feet_w;
    memcpy(t, rb + e * j.rb_stride, 40);
    memcpy(t + 10, cl + e * j.cf_stride, (size_t)fw * 4);
    memcpy(t + 10 + fw, cr + e * j.cf_stride, (size_t)fw * 4);
    const float* r = j.root + e * j.root_stride;              // dense rows: the hardware prefetcher follows them
    float* q = t + 10 + 2 * fw;
    q[0] = r[0]; q[1] = r[1]; q[2] = r[2];
    int used = 10 + 2 * fw + 3;
    if (STRIDE == 24 || L.root_n == 7) { q[3] = r[13]; q[4] = r[14]; q[5] = r[20]; q[6] = r[21]; used += 4; }
    for (int k = used; k < stride; ++k) t[k] = 0.0f;
    if (j.values) t[stride - 1] = j.values[e];               // the pad slot (always >= 1 float): value bootstrap of the reward epilogue
    float* d = j.rec + (e - j.env0) * stride;
#if defined(__x86_64__)
    for (int k = 0; k < stride; k += 4) _mm_stream_ps(d + k, _mm_load_ps(t + k));
#else
    memcpy(d, t, (size_t)stride * 4);
#endif
}

void pack_piece(const Job& j, int64_t lo, int64_t hi) {
    const float* rb = j.rb + j.rb_off;
    const float* cl = j.cf + j.cf_l;
    const float* cr = j.cf + j.cf_r;
    const PackLayout& L = j.L;
    const int fw = L.feet_w;
    const bool kick24 = L.stride == 24 && L.root_n == 7 && fw == 3;
    for (int64_t e = lo; e < hi; ++e) {
        const int64_t p = e + kPrefetchAhead;
        __builtin_prefetch(rb + p * j.rb_stride);
        __builtin_prefetch(rb + p * j.rb_stride + 9);
        __builtin_prefetch(cl + p * j.cf_stride);
        __builtin_prefetch(cl + p * j.cf_stride + fw - 1);
        __builtin_prefetch(cr + p * j.cf_stride);
        __builtin_prefetch(cr + p * j.cf_stride + fw - 1);
        __builtin_prefetch(j.root + (e + 2 * kPrefetchAhead) * j.root_stride);
        if (kick24) pack_one<24>(j, L, rb, cl, cr, e);
        else pack_one<0>(j, L, rb, cl, cr, e);
    }
    if (j.dof) {                                             // this piece's dof_state rows: one contiguous streaming copy
        const float* src = j.dof + lo * 36;
        float* dst = j.dof_dst + (lo - j.env0) * 36;
        const int64_t nf = (hi - lo) * 36;                   // a multiple of 4 floats
#if defined(__x86_64__)
        for (int64_t k = 0; k < nf; k += 4) {
            if ((k & 15) == 0) __builtin_prefetch(src + k + 256);
            _mm_stream_ps(dst + k, _mm_loadu_ps(src + k));
        }
#else
        memcpy(dst, src, (size_t)nf * 4);
#endif
    }
#if defined(__x86_64__)
    _mm_sfence();                                            // the streamed records are globally visible before the job is marked done
#endif
}

class Pool {
public:
    static Pool& get() { static Pool p; return p; }

    int ensure_threads(int want, int spin_us = -1, int pin = -1) {
        std::lock_guard<std::mutex> g(m_);
        if (spin_us >= 0) spin_us_.store(spin_us, std::memory_order_relaxed);
        if (pin >= 0 && workers_.empty()) pin_ = pin != 0;
        if (want <= 0 && !workers_.empty()) return (int)workers_.size();
        if (want <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            want = hw > 2 ? (int)hw - 1 : 1;          // leave the caller's core alone
            if (want > 16) want = 16;
        }
        if (want > 64) want = 64;
        while ((int)workers_.size() < want) { const int id = (int)workers_.size(); workers_.emplace_back([this, id] { run(id); }); }
        return (int)workers_.size();
    }

    int64_t begin(const Job& proto) {
        std::lock_guard<std::mutex> issue(issue_m_);    // issuing threads take turns (uncontended in the one-thread-per-env case)
        const int64_t t = seq_.fetch_add(1, std::memory_order_relaxed) + 1;
        Job& j = slots_[t % kSlots];
        while (j.ticket.load(std::memory_order_acquire) != 0) {      // the ring wrapped onto a job still in flight: finish it
            if (!help(j)) BEZK_CPU_RELAX();
        }
        // a worker that picked this slot up for its PREVIOUS job may still be about to claim from it: park the claim counter
        // out of range while the fields change, open it (release) only when the job is complete
        j.next.store(INT64_MAX / 2, std::memory_order_relaxed);
        j.rb = proto.rb; j.cf = proto.cf; j.root = proto.root; j.rec = proto.rec; j.dof = proto.dof; j.dof_dst = proto.dof_dst;
        j.values = proto.values;
        j.rb_stride = proto.rb_stride; j.rb_off = proto.rb_off; j.cf_stride = proto.cf_stride; j.cf_l = proto.cf_l; j.cf_r = proto.cf_r;
        j.root_stride = proto.root_stride; j.L = proto.L;
        j.env0 = proto.env0; j.n = proto.n;
        j.done.store(0, std::memory_order_relaxed);
        j.ticket.store(t, std::memory_order_release);
        j.next.store(0, std::memory_order_release);
        {
            std::lock_guard<std::mutex> g(m_);
            pending_.push_back(&j);
        }
        issued_.fetch_add(1, std::memory_order_acq_rel);
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        return t;
    }

    // Blocks until job `t` is packed; the caller packs pieces itself while it waits.
    int wait(int64_t t) {
        if (t == 0) return 0;                            // an empty job
        if (t < 0 || t > seq_.load(std::memory_order_relaxed)) return -1;
        Job& j = slots_[t % kSlots];
        for (;;) {
            const int64_t cur = j.ticket.load(std::memory_order_acquire);
            if (cur != t) return cur == 0 || cur > t ? 0 : -1;       // finished (slot freed or reused by a later job)
            if (!help(j)) BEZK_CPU_RELAX();
        }
    }

private:
    Pool() = default;
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            stopping_.store(true, std::memory_order_relaxed);
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }

    // Claims and packs one piece of `j`; false when nothing is left to claim.
    bool help(Job& j) {
        const int64_t lo = j.next.fetch_add(kPiece, std::memory_order_acq_rel);
        if (lo >= j.n) return false;
        const int64_t hi = lo + kPiece < j.n ? lo + kPiece : j.n;
        pack_piece(j, j.env0 + lo, j.env0 + hi);
        if (j.done.fetch_add(hi - lo, std::memory_order_acq_rel) + (hi - lo) == j.n) {
            {
                std::lock_guard<std::mutex> g(m_);
                for (size_t i = 0; i < pending_.size(); ++i)
                    if (pending_[i] == &j) { pending_.erase(pending_.begin() + i); break; }
            }
            j.ticket.store(0, std::memory_order_release);
        }
        return true;
    }

    void run(int id) {
        if (pin_) {                                  // one CPU per worker, never CPU 0 of the allowed set (left to the caller)
            cpu_set_t allowed, one;
            CPU_ZERO(&allowed);
            if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
                std::vector<int> cpus;
                for (int c = 0; c < CPU_SETSIZE; ++c)
                    if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
                if (cpus.size() > 1) {
                    CPU_ZERO(&one);
                    CPU_SET(cpus[1 + id % (cpus.size() - 1)], &one);
                    pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
                }
            }
        }
        for (;;) {
            const int64_t seen = issued_.load(std::memory_order_acquire);
            Job* j = nullptr;
            {
                std::lock_guard<std::mutex> g(m_);
                if (stop_) return;
                j = first_claimable();
            }
            if (j) {
                while (help(*j)) {}
                continue;
            }
            // Nothing to claim.  Poll for the next job for spin_us_ before blocking: the jobs of one step arrive microseconds
            // apart.  The default is SHORT (100 us): between steps the workers sleep and leave the cores to the simulator; on
            // the B200 hosts a futex wake costs the issuing thread ~30 us and polling for 0 / 100 / 2000 us measures the same
            // (profiles/r02_host_pack.md).  On hosts where a wake is expensive (an overcommitted guest: the woken thread runs
            // on the waker's core for milliseconds) raise it to the step period with bezk_host_pack_config.
            bool again = false;
            const int spin_us = spin_us_.load(std::memory_order_relaxed);
            if (spin_us > 0) {
                const auto t0 = std::chrono::steady_clock::now();
                for (int k = 0;; ++k) {
                    if (issued_.load(std::memory_order_acquire) != seen || stopping_.load(std::memory_order_relaxed)) { again = true; break; }
                    BEZK_CPU_RELAX();
                    if ((k & 255) == 255 &&
                        std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us)
                        break;
                }
            }
            if (again) continue;
            std::unique_lock<std::mutex> g(m_);
            sleepers_.fetch_add(1, std::memory_order_acq_rel);
            cv_.wait(g, [&] { return stop_ || first_claimable() != nullptr; });
            sleepers_.fetch_sub(1, std::memory_order_acq_rel);
            if (stop_) return;
        }
    }

    Job* first_claimable() {        // m_ held; jobs are served in issue order (chunk 0 of a step finishes first)
        for (Job* j : pending_)
            if (j->next.load(std::memory_order_relaxed) < j->n) return j;
        return nullptr;
    }

    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    std::vector<Job*> pending_;
    Job slots_[kSlots];
    std::atomic<int64_t> issued_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<int> spin_us_{100};
    std::atomic<bool> stopping_{false};
    std::mutex issue_m_;
    std::atomic<int64_t> seq_{0};
    bool stop_ = false;
    bool pin_ = false;
};

}  // namespace

int host_pack_config(int threads, int spin_us, int pin) { return Pool::get().ensure_threads(threads, spin_us, pin); }

// Record of one env, in floats: [imu-link q, v, w (10) | left foot (cleat) rows | right | root subset | pad to 16 B].
// root subset: robot position xyz (+ ball xy position and xy velocity for BezKick) -- all the fused step reads of root_states.
PackLayout pack_layout(int task, const BezkTaskCfg& cfg) {
    PackLayout L;
    L.feet_w = (cfg.flags & BEZK_F_CLEATS) ? 12 : 3;
    L.l_off = 10;
    L.r_off = L.l_off + L.feet_w;
    L.root_off = L.r_off + L.feet_w;
    L.root_n = task == BEZK_TASK_KICK ? 7 : 3;
    L.stride = (L.root_off + L.root_n + 1 + 3) / 4 * 4;       // >= 1 pad float: the critic value's slot
    return L;
}

int64_t host_pack_begin(int task, const float* rigid_body, const float* net_contact, const float* root_states, const float* dof_state,
                        const float* values, const BezkTaskCfg& cfg, float* dst, int64_t env0, int64_t n) {
    if (n <= 0) return 0;
    Pool& p = Pool::get();
    p.ensure_threads(0);
    Job j;
    j.rb = rigid_body; j.cf = net_contact; j.root = root_states;
    j.dof = dof_state; j.dof_dst = dof_state ? dst : nullptr;
    j.values = values;
    j.rec = dof_state ? dst + n * 36 : dst;                  // [n x 36 dof rows |] n records
    j.rb_stride = (int64_t)cfg.num_bodies * 13; j.rb_off = (int64_t)cfg.imu_body * 13 + 3;
    j.cf_stride = (int64_t)cfg.num_bodies * 3; j.cf_l = (int64_t)cfg.left_foot_body * 3; j.cf_r = (int64_t)cfg.right_foot_body * 3;
    j.root_stride = task == BEZK_TASK_KICK ? 26 : 13;
    j.L = pack_layout(task, cfg);
    j.env0 = env0; j.n = n;
    return p.begin(j);
}

int host_pack_wait(int64_t ticket) { return Pool::get().wait(ticket); }

}  // namespace bezk
