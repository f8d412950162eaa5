// Host-buffer entry points of libb200comp.so: what a ctypes / cffi binding of the reference's
// composite() / fill_solid() / fill_gradient() calls with plain host arrays.  They only use the
// device-level C ABI declared in include/b200comp.h plus CUDA runtime copies.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200comp.h"

extern "C" int b200comp_pool_alloc_(void **p, size_t bytes, void *stream);  // b200comp.cu: the library's own memory pool
extern "C" void b200comp_pool_trim_(void);

namespace {

thread_local std::string g_host_err;

// Device memory from the stream-ordered pool, allocated and freed on one of the call's own streams (never the
// legacy default stream, which would serialise concurrent callers).  The pool keeps freed blocks (release
// threshold raised below), so the staging buffers of repeated host-buffer calls cost a pool lookup.
struct DevBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        st = stream;
        return (cudaError_t)b200comp_pool_alloc_(&p, std::max<size_t>(bytes, 16), stream);
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// rows of an image between host and device; one linear copy when both sides are contiguous (the copy engine moves
// a linear 33 MB canvas faster than 2160 rows of a 2-D copy)
inline cudaError_t copy_image_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes, int rows,
                                    cudaMemcpyKind kind, cudaStream_t st) {
    if (dst_pitch == row_bytes && src_pitch == row_bytes) return cudaMemcpyAsync(dst, src, row_bytes * (size_t)rows, kind, st);
    return cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, (size_t)rows, kind, st);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int b200comp_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return B200COMP_EINVAL;
    return cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocDefault) == cudaSuccess ? 0 : B200COMP_ENOMEM;
}

int b200comp_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? 0 : B200COMP_ECUDA; }

// internal: set the message returned by b200comp_last_error() (defined in b200comp.cu)
int b200comp_set_error_(int code, const char *msg);
// internal: stream-ordered copy of a plan's status word into pinned host memory (defined in b200comp.cu)
int b200comp_plan_status_async_(b200comp_plan *plan, int *pinned_host_dst, void *stream);

// Pipelined host-buffer batch.  Canvases are processed in super-chunks (`8 * chunk_canvases`, double
// buffered on the device); inside a super-chunk three streams form a copy-in / compute / copy-out
// pipeline over sub-chunks of `chunk_canvases`, so both PCIe directions and the SMs work concurrently,
// while a helper thread resolves the NEXT super-chunk's plan (coefficient tables on the host threads).
int b200comp_composite_batch_host(const b200comp_canvas *canvases, int n_canvases,
                                  const b200comp_placement *placements, int n_placements, int n_host_threads,
                                  int chunk_canvases, int n_streams) {
    (void)n_streams;  // kept for ABI stability: the pipeline always uses three streams
    if (n_canvases < 1 || !canvases || n_placements < 0 || (n_placements > 0 && !placements))
        return b200comp_set_error_(B200COMP_EINVAL, "composite_batch_host: empty batch or null arrays");
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return b200comp_set_error_(B200COMP_ECUDA, "no CUDA device");
    // B200COMP_TRACE=1: host-side timeline of the pipeline on stderr
    static const bool trace = std::getenv("B200COMP_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what, int idx) {
        if (!trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        std::fprintf(stderr, "[b200comp host] %8.2f ms  %s %d\n", ms, what, idx);
    };
    if (n_host_threads <= 0) n_host_threads = std::max(1u, std::thread::hardware_concurrency());

    // ---- validate, size the staging buffers, find the distinct cutouts ----
    typedef std::tuple<const uint8_t *, int, int, int64_t> SrcKey;
    std::map<SrcKey, size_t> src_off;
    std::vector<SrcKey> src_order;
    size_t pool_bytes = 0;
    for (int i = 0; i < n_placements; ++i) {
        const b200comp_placement &p = placements[i];
        if (!p.src || p.sw < 1 || p.sh < 1 || p.w < 1 || p.h < 1 || p.src_pitch < (int64_t)p.sw * 4)
            return b200comp_set_error_(B200COMP_EINVAL, "composite_batch_host: bad placement");
        SrcKey k(p.src, p.sw, p.sh, p.src_pitch);
        if (src_off.find(k) == src_off.end()) {
            src_off[k] = pool_bytes;
            src_order.push_back(k);
            pool_bytes = align_up(pool_bytes + align_up((size_t)p.sw * 4, 16) * p.sh, 256);
        }
    }
    size_t max_canvas_bytes = 0;
    bool any_bg = false;
    for (int c = 0; c < n_canvases; ++c) {
        const b200comp_canvas &cv = canvases[c];
        if (!cv.out || cv.W < 1 || cv.H < 1 || cv.out_pitch < (int64_t)cv.W * 4 || (cv.bg && cv.bg_pitch < (int64_t)cv.W * 4))
            return b200comp_set_error_(B200COMP_EINVAL, "composite_batch_host: bad canvas");
        if (cv.n_placements < 0 || cv.first_placement < 0 || (int64_t)cv.first_placement + cv.n_placements > n_placements)
            return b200comp_set_error_(B200COMP_EINVAL, "composite_batch_host: placement range out of bounds");
        max_canvas_bytes = std::max(max_canvas_bytes, align_up((size_t)cv.W * 4, 16) * cv.H);
        any_bg |= cv.bg != nullptr;
    }
    max_canvas_bytes = align_up(max_canvas_bytes, 256);
    // sub-chunk = unit of the copy-in / compute / copy-out pipeline.  Default: about 32 MB of canvas per sub-chunk
    // (one 4K canvas: measured best on the B200's PCIe link, 9.1 vs 8.4 GB of canvas per second with four), so
    // small canvases still travel in copies large enough to amortise their launch.
    if (chunk_canvases <= 0)
        chunk_canvases = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)32 << 20) / max_canvas_bytes));
    // super-chunk (one staging set, one plan): about a quarter of the batch, between 2 and 16 sub-chunks
    int super_canvases = ((n_canvases + 3) / 4 + chunk_canvases - 1) / chunk_canvases * chunk_canvases;
    super_canvases = std::max(2 * chunk_canvases, std::min(16 * chunk_canvases, super_canvases));
    super_canvases = std::min(n_canvases, super_canvases);
    const int n_super = (n_canvases + super_canvases - 1) / super_canvases;

    // ---- device resources ----
    // three staging sets: while the host waits for the oldest super-chunk, two more are queued on the copy
    // streams, so neither PCIe direction idles
    constexpr int kBufs = 3;
    const int n_buf = std::min(n_super, kBufs);
    cudaStream_t s_in = nullptr, s_exec = nullptr, s_out = nullptr, s_plan = nullptr;
    struct StreamGuard {
        cudaStream_t *s[4];
        ~StreamGuard() { for (auto p : s) if (*p) cudaStreamDestroy(*p); }
    } sguard{{&s_in, &s_exec, &s_out, &s_plan}};
    if (cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s_exec, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s_plan, cudaStreamNonBlocking) != cudaSuccess)
        return b200comp_set_error_(B200COMP_ECUDA, "stream creation failed");
    // this call's streams only: nothing here waits for, or makes wait, other threads' work on the device
    auto sync_all = [&] {
        cudaStreamSynchronize(s_in);
        cudaStreamSynchronize(s_exec);
        cudaStreamSynchronize(s_out);
        cudaStreamSynchronize(s_plan);
    };
    // staging memory: declared after the streams, so it is freed (stream-ordered, on s_exec) before they are destroyed
    DevBuf pool, pool_tight, d_out[kBufs], d_bg[kBufs];
    if (pool.alloc(pool_bytes, s_exec) != cudaSuccess) return b200comp_set_error_(B200COMP_ENOMEM, "cutout pool allocation failed");
    // Cutouts whose rows are not a multiple of 16 bytes get a padded pitch on the device (TMA, 128-bit loads).  A 2-D
    // host-to-device copy of kilobyte rows runs far below the link rate, so tightly packed host cutouts cross PCIe as
    // ONE linear copy each into `pool_tight` and are re-pitched by a device-to-device copy.
    size_t tight_bytes = 0;
    for (const SrcKey &k : src_order) {
        const size_t row = (size_t)std::get<1>(k) * 4;
        if ((size_t)std::get<3>(k) == row && align_up(row, 16) != row) tight_bytes += align_up(row * std::get<2>(k), 256);
    }
    if (tight_bytes && pool_tight.alloc(tight_bytes, s_exec) != cudaSuccess)
        return b200comp_set_error_(B200COMP_ENOMEM, "cutout pool allocation failed");
    for (int b = 0; b < n_buf; ++b)
        if (d_out[b].alloc(max_canvas_bytes * super_canvases, s_exec) != cudaSuccess ||
            (any_bg && d_bg[b].alloc(max_canvas_bytes * super_canvases, s_exec) != cudaSuccess))
            return b200comp_set_error_(B200COMP_ENOMEM, "canvas staging allocation failed");
    cudaStreamSynchronize(s_exec);  // the allocations are ordered on s_exec; the other streams may use them from here on
    stamp("staging allocated", 0);
    const int max_sub = (super_canvases + chunk_canvases - 1) / chunk_canvases;
    std::vector<cudaEvent_t> ev_in((size_t)max_sub * 3), ev_exec((size_t)max_sub * 3);  // one set per staging buffer
    struct EventGuard {
        std::vector<cudaEvent_t> *a, *b;
        ~EventGuard() { for (auto e : *a) if (e) cudaEventDestroy(e); for (auto e : *b) if (e) cudaEventDestroy(e); }
    } eguard{&ev_in, &ev_exec};
    for (auto &e : ev_in) e = nullptr;
    for (auto &e : ev_exec) e = nullptr;
    cudaEvent_t ev_pool = nullptr, ev_done[kBufs] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3 * max_sub; ++i)
        if (cudaEventCreateWithFlags(&ev_in[(size_t)i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_exec[(size_t)i], cudaEventDisableTiming) != cudaSuccess)
            return b200comp_set_error_(B200COMP_ECUDA, "event creation failed");
    cudaEventCreateWithFlags(&ev_pool, cudaEventDisableTiming);
    for (int b = 0; b < kBufs; ++b) cudaEventCreateWithFlags(&ev_done[b], cudaEventDisableTiming);

    // cutouts: each distinct host cutout once, 16-byte aligned pitch (copy-in stream)
    size_t tight_off = 0;
    for (const SrcKey &k : src_order) {
        const int sw = std::get<1>(k), sh = std::get<2>(k);
        const size_t row = (size_t)sw * 4, dpitch = align_up(row, 16);
        uint8_t *dst = (uint8_t *)pool.p + src_off[k];
        if ((size_t)std::get<3>(k) == row && dpitch != row) {
            uint8_t *tmp = (uint8_t *)pool_tight.p + tight_off;
            tight_off += align_up(row * sh, 256);
            cudaMemcpyAsync(tmp, std::get<0>(k), row * sh, cudaMemcpyHostToDevice, s_in);
            cudaMemcpy2DAsync(dst, dpitch, tmp, row, row, sh, cudaMemcpyDeviceToDevice, s_in);
        } else {
            copy_image_async(dst, dpitch, std::get<0>(k), (size_t)std::get<3>(k), row, sh, cudaMemcpyHostToDevice, s_in);
        }
    }
    cudaEventRecord(ev_pool, s_in);

    // ---- plan of one super-chunk (runs on the helper thread: host-side table building) ----
    struct SuperPlan {
        b200comp_plan *plan = nullptr;
        std::vector<b200comp_canvas> cc;
        std::vector<b200comp_placement> pp;
        int rc = 0;
        std::string err;
    };
    auto build_plan = [&](int si, SuperPlan *sp) {
        cudaSetDevice(device);
        const int c_lo = si * super_canvases, c_hi = std::min(n_canvases, c_lo + super_canvases);
        const int buf = si % n_buf;
        for (int c = c_lo; c < c_hi; ++c) {
            b200comp_canvas cv = canvases[c];
            const size_t dp = align_up((size_t)cv.W * 4, 16);
            if (cv.bg) {
                cv.bg = (const uint8_t *)d_bg[buf].p + (size_t)(c - c_lo) * max_canvas_bytes;
                cv.bg_pitch = (int64_t)dp;
            }
            cv.out = (uint8_t *)d_out[buf].p + (size_t)(c - c_lo) * max_canvas_bytes;
            cv.out_pitch = (int64_t)dp;
            const int first = (int)sp->pp.size();
            for (int i = 0; i < cv.n_placements; ++i) {
                b200comp_placement p = placements[cv.first_placement + i];
                SrcKey k(p.src, p.sw, p.sh, p.src_pitch);
                p.src = (const uint8_t *)pool.p + src_off[k];
                p.src_pitch = (int64_t)align_up((size_t)p.sw * 4, 16);
                sp->pp.push_back(p);
            }
            cv.first_placement = first;
            sp->cc.push_back(cv);
        }
        sp->rc = b200comp_plan_create(sp->cc.data(), (int)sp->cc.size(), sp->pp.data(), (int)sp->pp.size(),
                                      n_host_threads, s_plan, &sp->plan);
        if (sp->rc) sp->err = b200comp_last_error();
    };

    int rc = 0;
    std::string err;
    // one pinned status word per staging set (allocated once per thread, kept for the thread's lifetime)
    thread_local int *h_status = nullptr;
    if (!h_status && cudaHostAlloc((void **)&h_status, kBufs * sizeof(int), cudaHostAllocDefault) != cudaSuccess)
        return b200comp_set_error_(B200COMP_ENOMEM, "pinned status allocation failed");
    SuperPlan live[kBufs];  // plans of the super-chunks in flight, by staging buffer
    SuperPlan built;    // plan being resolved by the helper thread
    auto retire = [&](int buf) {  // wait for the super-chunk that used `buf`, check it, free its plan
        if (!live[buf].plan) return;
        if (rc) sync_all();  // error path: nothing may still be using the plan's memory
        cudaError_t e = cudaEventSynchronize(ev_done[buf]);
        if (rc == 0 && e != cudaSuccess) {
            rc = B200COMP_ECUDA;
            err = cudaGetErrorString(e);
        }
        if (rc == 0 && h_status[buf] != 0) {
            // the status word travelled on the copy-out stream right behind this super-chunk's canvases (a copy
            // of its own on another stream would queue behind the NEXT super-chunks' canvases in the copy engine
            // and stall the host for their whole transfer)
            rc = B200COMP_EINTERNAL;
            err = "binning reported a sizing violation (status " + std::to_string(h_status[buf]) + ")";
        }
        cudaStreamSynchronize(s_plan);
        b200comp_plan_destroy(live[buf].plan);
        live[buf] = SuperPlan();
    };
    std::thread helper(build_plan, 0, &built);
    for (int si = 0; si < n_super; ++si) {
        const int c_lo = si * super_canvases, c_hi = std::min(n_canvases, c_lo + super_canvases);
        const int buf = si % n_buf;
        const int n_sub = (c_hi - c_lo + chunk_canvases - 1) / chunk_canvases;
        cudaEvent_t *e_in = ev_in.data() + (size_t)buf * max_sub, *e_exec = ev_exec.data() + (size_t)buf * max_sub;
        retire(buf);  // the previous user of this staging buffer (two super-chunks ago)
        stamp("staging buffer free", si);
        if (rc) break;
        // copy-in of the backgrounds overlaps the helper thread's table building
        for (int j = 0; j < n_sub; ++j) {
            const int lo = c_lo + j * chunk_canvases, hi = std::min(c_hi, lo + chunk_canvases);
            for (int c = lo; c < hi; ++c) {
                const b200comp_canvas &cv = canvases[c];
                if (!cv.bg) continue;
                copy_image_async((uint8_t *)d_bg[buf].p + (size_t)(c - c_lo) * max_canvas_bytes, align_up((size_t)cv.W * 4, 16),
                                 cv.bg, (size_t)cv.bg_pitch, (size_t)cv.W * 4, cv.H, cudaMemcpyHostToDevice, s_in);
            }
            cudaEventRecord(e_in[j], s_in);
        }
        stamp("copy-in enqueued, waiting for plan", si);
        helper.join();  // plan of this super-chunk
        stamp("plan ready", si);
        live[buf] = std::move(built);
        built = SuperPlan();
        if (live[buf].rc) {
            rc = live[buf].rc;
            err = live[buf].err;
            helper = std::thread([] {});
            break;
        }
        if (si + 1 < n_super)
            helper = std::thread(build_plan, si + 1, &built);
        else
            helper = std::thread([] {});
        cudaStreamSynchronize(s_plan);  // descriptors and tables of this plan are on the device
        cudaStreamWaitEvent(s_exec, ev_pool, 0);
        rc = b200comp_plan_prepare(live[buf].plan, s_exec);
        for (int j = 0; j < n_sub && rc == 0; ++j) {
            const int lo = c_lo + j * chunk_canvases, hi = std::min(c_hi, lo + chunk_canvases);
            cudaStreamWaitEvent(s_exec, e_in[j], 0);
            rc = b200comp_plan_run_canvases(live[buf].plan, lo - c_lo, hi - lo, s_exec);
            cudaEventRecord(e_exec[j], s_exec);
            cudaStreamWaitEvent(s_out, e_exec[j], 0);
            for (int c = lo; c < hi; ++c) {
                const b200comp_canvas &cv = canvases[c];
                copy_image_async(cv.out, (size_t)cv.out_pitch, (uint8_t *)d_out[buf].p + (size_t)(c - c_lo) * max_canvas_bytes,
                                 align_up((size_t)cv.W * 4, 16), (size_t)cv.W * 4, cv.H, cudaMemcpyDeviceToHost, s_out);
            }
        }
        h_status[buf] = 0;
        if (rc == 0) rc = b200comp_plan_status_async_(live[buf].plan, &h_status[buf], s_out);
        cudaEventRecord(ev_done[buf], s_out);
        stamp("super-chunk enqueued", si);
        if (rc) {
            err = b200comp_last_error();
            break;
        }
    }
    if (helper.joinable()) helper.join();
    if (built.plan) {  // built but never launched (error path)
        cudaStreamSynchronize(s_plan);
        b200comp_plan_destroy(built.plan);
    }
    {
        const int keep_rc = rc;
        const std::string keep_err = err;
        for (int b = 0; b < kBufs; ++b) retire(b);
        if (keep_rc) {
            rc = keep_rc;
            err = keep_err;
        }
    }
    sync_all();  // everything this call queued is done: the staging buffers can go back to the pool
    stamp("all done", 0);
    if (ev_pool) cudaEventDestroy(ev_pool);
    for (int b = 0; b < kBufs; ++b) if (ev_done[b]) cudaEventDestroy(ev_done[b]);
    stamp("events destroyed", 0);
    if (rc) return b200comp_set_error_(rc, err.c_str());
    return 0;
}

// ---- single canvas, host buffers: the call behind the drop-in composite() (compositor.py:6-22) ----
// Everything a call needs besides the plan is cached per thread and only ever grows: one stream, one event, pinned
// bounce buffers (pageable caller memory is copied through them in chunks, so the copy engine overlaps the host
// memcpy), device staging for background, canvas and cutouts.  No helper thread, no device-wide synchronisation:
// concurrent callers (one Streamlit session thread each) do not serialise on anything but the GPU itself.
namespace {

struct LeanCtx {
    int device = -1;
    cudaStream_t st = nullptr;
    uint8_t *pin_in = nullptr, *pin_out = nullptr;
    size_t pin_in_cap = 0, pin_out_cap = 0;
    uint8_t *d_bg = nullptr, *d_out = nullptr, *d_pool = nullptr;
    size_t d_bg_cap = 0, d_out_cap = 0, d_pool_cap = 0;
    int *h_status = nullptr;
    std::vector<cudaEvent_t> events;  // one per copy-out chunk, created once

    void release() {
        if (device >= 0) {
            int cur = 0;
            cudaGetDevice(&cur);
            if (cur != device) cudaSetDevice(device);
            if (st) cudaStreamSynchronize(st);
            if (pin_in) cudaFreeHost(pin_in);
            if (pin_out) cudaFreeHost(pin_out);
            if (d_bg) cudaFree(d_bg);
            if (d_out) cudaFree(d_out);
            if (d_pool) cudaFree(d_pool);
            if (h_status) cudaFreeHost(h_status);
            for (cudaEvent_t e : events) cudaEventDestroy(e);
            if (st) cudaStreamDestroy(st);
            if (cur != device) cudaSetDevice(cur);
        }
        *this = LeanCtx();
    }
    ~LeanCtx() {
        // thread exit: the CUDA context may already be gone at process exit; errors are ignored
        if (device >= 0 && cudaSetDevice(device) == cudaSuccess) release();
    }
    static bool grow_pinned(uint8_t **p, size_t *cap, size_t need) {
        if (need <= *cap) return true;
        if (*p) cudaFreeHost(*p);
        *p = nullptr;
        *cap = 0;
        const size_t want = align_up(need + need / 4, 1 << 20);
        if (cudaHostAlloc((void **)p, want, cudaHostAllocDefault) != cudaSuccess) return false;
        *cap = want;
        return true;
    }
    static bool grow_device(uint8_t **p, size_t *cap, size_t need, cudaStream_t st) {
        if (need <= *cap) return true;
        if (*p) {
            cudaStreamSynchronize(st);
            cudaFree(*p);
        }
        *p = nullptr;
        *cap = 0;
        const size_t want = align_up(need + need / 4, 1 << 20);
        if (cudaMalloc((void **)p, want) != cudaSuccess) return false;
        *cap = want;
        return true;
    }
};
thread_local LeanCtx g_lean;

// ---- a few helper threads for large host memcpys ---------------------------------------------------------
// One core moves about 10 GB/s, which would make the host copy, not PCIe, the slowest stage of a 4K single-canvas
// call (33 MB in, 33 MB out, tens of MB of cutouts).  The workers are started once and sleep on a condition variable
// (spawning threads per copy cost ~30 us each); one parallel loop runs at a time -- concurrent callers queue up, which
// costs nothing: the loop is memory-bandwidth bound.  Never destroyed: worker threads must not be joined from a static
// destructor at process exit.
class CopyPool {
public:
    static CopyPool &get() {
        static CopyPool *pool = new CopyPool();
        return *pool;
    }
    int width() const { return (int)workers_.size() + 1; }
    // fn(i) for i in [0, n): on the workers and the calling thread
    void parallel_for(int n, const std::function<void(int)> &fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) {
            for (int i = 0; i < n; ++i) fn(i);
            return;
        }
        std::lock_guard<std::mutex> one_loop(loop_mu_);
        {
            std::lock_guard<std::mutex> lock(mu_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            left_ = n;
            ++generation_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lock(mu_);
        done_cv_.wait(lock, [&] { return left_ == 0; });
        fn_ = nullptr;
    }

private:
    CopyPool() {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int n = (int)std::min(7u, hw > 1 ? hw - 1 : 0u);
        for (int i = 0; i < n; ++i) {
            workers_.emplace_back([this] {
                uint64_t seen = 0;
                for (;;) {
                    {
                        std::unique_lock<std::mutex> lock(mu_);
                        cv_.wait(lock, [&] { return generation_ != seen; });
                        seen = generation_;
                    }
                    work();
                }
            });
            workers_.back().detach();
        }
    }
    void work() {
        for (;;) {
            const std::function<void(int)> *fn;
            int i;
            {
                std::lock_guard<std::mutex> lock(mu_);
                fn = fn_;
                i = fn ? next_.fetch_add(1) : n_;
                if (!fn || i >= n_) return;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> lock(mu_);
            if (--left_ == 0) done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex loop_mu_, mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, left_ = 0;
    uint64_t generation_ = 0;
};

struct RowCopy {
    uint8_t *dst;
    size_t dst_pitch;
    const uint8_t *src;
    size_t src_pitch, row_bytes;
    int rows;
};
inline void copy_rows_serial(const RowCopy &c, int r0, int r1) {
    if (c.dst_pitch == c.src_pitch && c.row_bytes + 16 > c.dst_pitch) {
        std::memcpy(c.dst + (size_t)r0 * c.dst_pitch, c.src + (size_t)r0 * c.src_pitch,
                    (size_t)(r1 - r0) * c.dst_pitch - (c.dst_pitch - c.row_bytes));
    } else {
        for (int r = r0; r < r1; ++r) std::memcpy(c.dst + (size_t)r * c.dst_pitch, c.src + (size_t)r * c.src_pitch, c.row_bytes);
    }
}
// host memcpy of a set of images (rows of each); small sets on the calling thread, large ones in ~1 MB pieces on the pool
void copy_images(const RowCopy *copies, int n) {
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += copies[i].row_bytes * (size_t)copies[i].rows;
    // Below 4 MB the calling thread copies alone: waking the workers costs 50-150 us on the boxes measured (a 1 MB canvas
    // call went from 0.42 to 0.73 ms with the pool), a 1 MB memcpy about as much.  B200COMP_COPY_POOL_MIN=bytes overrides.
    static const size_t serial_below = [] {
        const char *e = std::getenv("B200COMP_COPY_POOL_MIN");
        return e && e[0] ? (size_t)std::atoll(e) : ((size_t)4 << 20);
    }();
    if (total < serial_below) {
        for (int i = 0; i < n; ++i) copy_rows_serial(copies[i], 0, copies[i].rows);
        return;
    }
    // pieces of 1 MB for large sets, smaller ones (down to 64 KB) so that a 1 MB canvas still spreads over the workers
    const size_t piece_bytes = std::min<size_t>((size_t)1 << 20, std::max<size_t>((size_t)64 << 10, total / 16));
    struct Piece { int img, r0, r1; };
    std::vector<Piece> pieces;
    for (int i = 0; i < n; ++i) {
        const int step = (int)std::max<size_t>(1, piece_bytes / std::max<size_t>(1, copies[i].row_bytes));
        for (int r0 = 0; r0 < copies[i].rows; r0 += step) pieces.push_back(Piece{i, r0, std::min(copies[i].rows, r0 + step)});
    }
    CopyPool::get().parallel_for((int)pieces.size(), [&](int k) { copy_rows_serial(copies[pieces[(size_t)k].img], pieces[(size_t)k].r0, pieces[(size_t)k].r1); });
}
void copy_rows(uint8_t *dst, size_t dst_pitch, const uint8_t *src, size_t src_pitch, size_t row_bytes, int rows) {
    const RowCopy c{dst, dst_pitch, src, src_pitch, row_bytes, rows};
    copy_images(&c, 1);
}

// rows of a pageable (or pinned) host image -> device, through the pinned bounce buffer in ~4 MB chunks
void upload_rows(uint8_t *dev, size_t dev_pitch, const uint8_t *host, size_t host_pitch, size_t row_bytes, int rows,
                 uint8_t *pin, cudaStream_t st) {
    const int chunk_rows = (int)std::max<size_t>(1, ((size_t)8 << 20) / std::max<size_t>(1, dev_pitch));
    for (int r0 = 0; r0 < rows; r0 += chunk_rows) {
        const int n = std::min(chunk_rows, rows - r0);
        uint8_t *p = pin + (size_t)r0 * dev_pitch;
        copy_rows(p, dev_pitch, host + (size_t)r0 * host_pitch, host_pitch, row_bytes, n);
        cudaMemcpyAsync(dev + (size_t)r0 * dev_pitch, p, (size_t)n * dev_pitch, cudaMemcpyHostToDevice, st);
    }
}

// the calling thread's context on its current device (created on first use)
int acquire_lean(LeanCtx **out, const char *who) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return b200comp_set_error_(B200COMP_ECUDA, "no CUDA device");
    LeanCtx &cx = g_lean;
    if (cx.device != device) {
        cx.release();
        cx.device = device;
        if (cudaStreamCreateWithFlags(&cx.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaHostAlloc((void **)&cx.h_status, sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
            cx.release();
            return b200comp_set_error_(B200COMP_ECUDA, (std::string(who) + ": stream / pinned status allocation failed").c_str());
        }
    }
    *out = &cx;
    return 0;
}

// device canvas (pitch dp) -> caller's host canvas, in chunks: while the copy engine fills the next chunk of the
// pinned buffer, the host copies the previous one into the caller's (pageable) memory.  Synchronises cx.st.
int download_rows(LeanCtx &cx, uint8_t *out, size_t out_pitch, const uint8_t *dev, size_t dp, size_t row_bytes, int H,
                  const char *who) {
    cudaStream_t st = cx.st;
    const int chunk_rows = (int)std::max<size_t>(1, ((size_t)8 << 20) / dp);
    const int n_chunks = (H + chunk_rows - 1) / chunk_rows;
    while ((int)cx.events.size() < n_chunks) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            cudaStreamSynchronize(st);
            return b200comp_set_error_(B200COMP_ECUDA, (std::string(who) + ": event creation failed").c_str());
        }
        cx.events.push_back(ev);
    }
    std::vector<cudaEvent_t> &evs = cx.events;
    for (int c = 0; c < n_chunks; ++c) {
        const int r0 = c * chunk_rows, n = std::min(chunk_rows, H - r0);
        cudaMemcpyAsync(cx.pin_out + (size_t)r0 * dp, dev + (size_t)r0 * dp, (size_t)n * dp, cudaMemcpyDeviceToHost, st);
        cudaEventRecord(evs[(size_t)c], st);
    }
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < n_chunks; ++c) {
        const int r0 = c * chunk_rows, n = std::min(chunk_rows, H - r0);
        const cudaError_t ec = cudaEventSynchronize(evs[(size_t)c]);
        if (ec != cudaSuccess) e = ec;
        if (e == cudaSuccess) copy_rows(out + (size_t)r0 * out_pitch, out_pitch, cx.pin_out + (size_t)r0 * dp, dp, row_bytes, n);
    }
    if (e != cudaSuccess) return b200comp_set_error_(B200COMP_ECUDA, cudaGetErrorString(e));
    return 0;
}

}  // namespace

int b200comp_composite_host_ex(const uint8_t *bg, uint32_t solid_rgba, int W, int H, size_t bg_pitch, uint8_t *out,
                               size_t out_pitch, const b200comp_placement *placements, int n_placements) {
    if (!out || W < 1 || H < 1 || out_pitch < (size_t)W * 4 || (bg && bg_pitch < (size_t)W * 4) || n_placements < 0 ||
        (n_placements > 0 && !placements))
        return b200comp_set_error_(B200COMP_EINVAL, "composite_host: bad canvas or placement argument");
    LeanCtx *cxp = nullptr;
    if (int rc = acquire_lean(&cxp, "composite_host")) return rc;
    LeanCtx &cx = *cxp;
    cudaStream_t st = cx.st;
    const size_t dp = align_up((size_t)W * 4, 16), canvas_bytes = dp * H;

    // distinct host cutouts of the call -> one staging pool (16-byte aligned pitches); device cutouts are used in place
    struct Src { const uint8_t *p; int sw, sh; int64_t pitch; size_t off; };
    std::vector<Src> srcs;
    std::vector<b200comp_placement> pp((size_t)n_placements);
    size_t pool_bytes = 0;
    for (int i = 0; i < n_placements; ++i) {
        b200comp_placement p = placements[i];
        if (!p.src || p.sw < 1 || p.sh < 1 || p.w < 1 || p.h < 1 || p.src_pitch < (int64_t)p.sw * 4)
            return b200comp_set_error_(B200COMP_EINVAL, "composite_host: bad placement");
        if (!(p.flags & B200COMP_SRC_DEVICE)) {
            size_t k = 0;
            for (; k < srcs.size(); ++k)
                if (srcs[k].p == p.src && srcs[k].sw == p.sw && srcs[k].sh == p.sh && srcs[k].pitch == p.src_pitch) break;
            if (k == srcs.size()) {
                srcs.push_back(Src{p.src, p.sw, p.sh, p.src_pitch, pool_bytes});
                pool_bytes = align_up(pool_bytes + align_up((size_t)p.sw * 4, 16) * p.sh, 256);
            }
            p.src = reinterpret_cast<const uint8_t *>(srcs[k].off);  // offset for now; the pool may still move
            p.src_pitch = (int64_t)align_up((size_t)p.sw * 4, 16);
        }
        p.flags &= ~B200COMP_SRC_DEVICE;
        pp[(size_t)i] = p;
    }
    const size_t in_bytes = pool_bytes + (bg ? canvas_bytes : 0);
    if (!LeanCtx::grow_pinned(&cx.pin_in, &cx.pin_in_cap, in_bytes) || !LeanCtx::grow_pinned(&cx.pin_out, &cx.pin_out_cap, canvas_bytes) ||
        !LeanCtx::grow_device(&cx.d_out, &cx.d_out_cap, canvas_bytes, st) ||
        (bg && !LeanCtx::grow_device(&cx.d_bg, &cx.d_bg_cap, canvas_bytes, st)) ||
        (pool_bytes && !LeanCtx::grow_device(&cx.d_pool, &cx.d_pool_cap, pool_bytes, st)))
        return b200comp_set_error_(B200COMP_ENOMEM, "composite_host: staging allocation failed");
    for (int i = 0; i < n_placements; ++i)
        if (!(placements[i].flags & B200COMP_SRC_DEVICE)) pp[(size_t)i].src = cx.d_pool + reinterpret_cast<size_t>(pp[(size_t)i].src);

    // copy-in: cutouts first (the plan's prepare kernel reads them), then the background
    if (!srcs.empty()) {
        std::vector<RowCopy> cc;
        for (const Src &s : srcs)
            cc.push_back(RowCopy{cx.pin_in + s.off, align_up((size_t)s.sw * 4, 16), s.p, (size_t)s.pitch, (size_t)s.sw * 4, s.sh});
        copy_images(cc.data(), (int)cc.size());  // all cutouts in one parallel pass over the helper threads
        cudaMemcpyAsync(cx.d_pool, cx.pin_in, pool_bytes, cudaMemcpyHostToDevice, st);
    }
    if (bg) upload_rows(cx.d_bg, dp, bg, bg_pitch, (size_t)W * 4, H, cx.pin_in + pool_bytes, st);

    b200comp_canvas cv;
    std::memset(&cv, 0, sizeof cv);
    cv.out = cx.d_out;
    cv.out_pitch = (int64_t)dp;
    cv.bg = bg ? cx.d_bg : nullptr;
    cv.bg_pitch = bg ? (int64_t)dp : 0;
    cv.solid_rgba = solid_rgba;
    cv.W = W;
    cv.H = H;
    cv.first_placement = 0;
    cv.n_placements = n_placements;
    b200comp_plan *plan = nullptr;
    int rc = b200comp_plan_create(&cv, 1, pp.data(), n_placements, 1, st, &plan);
    if (rc) return rc;
    rc = b200comp_plan_run(plan, st);
    *cx.h_status = 0;
    if (!rc) rc = b200comp_plan_status_async_(plan, cx.h_status, st);
    if (rc) {
        cudaStreamSynchronize(st);
        b200comp_plan_destroy(plan);
        return rc;
    }
    rc = download_rows(cx, out, out_pitch, cx.d_out, dp, (size_t)W * 4, H, "composite_host");
    const int status = *cx.h_status;  // travelled on the stream ahead of the canvas
    b200comp_plan_destroy(plan);      // (its stream is idle: every chunk event has fired)
    if (rc) return rc;
    if (status != 0)
        return b200comp_set_error_(B200COMP_EINTERNAL, ("tile kernel / binning status " + std::to_string(status)).c_str());
    return 0;
}

int b200comp_composite_host(const uint8_t *bg, int W, int H, size_t bg_pitch, uint8_t *out, size_t out_pitch,
                            const b200comp_placement *placements, int n_placements) {
    if (!bg) return b200comp_set_error_(B200COMP_EINVAL, "composite_host: null canvas");
    return b200comp_composite_host_ex(bg, 0u, W, H, bg_pitch, out, out_pitch, placements, n_placements);
}

int b200comp_device_upload(const uint8_t *img, int w, int h, size_t pitch, uint8_t **dev, size_t *dev_pitch) {
    if (!img || !dev || !dev_pitch || w < 1 || h < 1 || pitch < (size_t)w * 4)
        return b200comp_set_error_(B200COMP_EINVAL, "device_upload: bad argument");
    const size_t dp = align_up((size_t)w * 4, 16);
    uint8_t *d = nullptr;
    if (cudaMalloc((void **)&d, dp * h) != cudaSuccess) return b200comp_set_error_(B200COMP_ENOMEM, "device_upload: allocation failed");
    // through the calling thread's pinned bounce buffer and stream (no legacy-stream copy of pageable memory)
    LeanCtx *cx = nullptr;
    cudaError_t e = cudaSuccess;
    if (acquire_lean(&cx, "device_upload") != 0 || !LeanCtx::grow_pinned(&cx->pin_in, &cx->pin_in_cap, dp * h)) {
        e = cudaErrorMemoryAllocation;
    } else {
        upload_rows(d, dp, img, pitch, (size_t)w * 4, h, cx->pin_in, cx->st);
        e = cudaStreamSynchronize(cx->st);
    }
    if (e != cudaSuccess) {
        cudaFree(d);
        return b200comp_set_error_(B200COMP_ECUDA, cudaGetErrorString(e));
    }
    *dev = d;
    *dev_pitch = dp;
    return 0;
}

int b200comp_device_free(uint8_t *dev) { return cudaFree(dev) == cudaSuccess ? 0 : b200comp_set_error_(B200COMP_ECUDA, "device_free failed"); }

int b200comp_trim(void) {
    g_lean.release();
    b200comp_pool_trim_();
    return 0;
}

// decoded image (HOST) -> device copy in the calling thread's staging buffer (pitch = 16-byte aligned row), on its stream
static int upload_image(LeanCtx &cx, const uint8_t *img, int W, int H, size_t pitch, const uint8_t **dev, size_t *dev_pitch,
                        const char *who) {
    if (!img || W < 1 || H < 1 || pitch < (size_t)W * 4)
        return b200comp_set_error_(B200COMP_EINVAL, (std::string(who) + ": bad image argument").c_str());
    const size_t dp = align_up((size_t)W * 4, 16), bytes = dp * H;
    if (!LeanCtx::grow_pinned(&cx.pin_in, &cx.pin_in_cap, bytes) || !LeanCtx::grow_device(&cx.d_bg, &cx.d_bg_cap, bytes, cx.st))
        return b200comp_set_error_(B200COMP_ENOMEM, (std::string(who) + ": staging allocation failed").c_str());
    upload_rows(cx.d_bg, dp, img, pitch, (size_t)W * 4, H, cx.pin_in, cx.st);
    *dev = cx.d_bg;
    *dev_pitch = dp;
    return 0;
}

// _edge_strip_median_colors (background_resizing.py:36-55): left, right, top, bottom strips
static int edge_medians_dev(const uint8_t *d_img, size_t pitch, int W, int H, int strip_px, int32_t edges[12], cudaStream_t st) {
    const int rects[4][4] = {{0, 0, std::min(strip_px, W), H},
                             {std::max(0, W - strip_px), 0, W, H},
                             {0, 0, W, std::min(strip_px, H)},
                             {0, std::max(0, H - strip_px), W, H}};
    for (int i = 0; i < 4; ++i) {
        int rc = b200comp_masked_median_rgb(d_img, W, H, pitch, rects[i][0], rects[i][1], rects[i][2], rects[i][3],
                                            edges + 3 * i, st);
        if (rc) return rc;
    }
    return 0;
}

int b200comp_masked_median_rgb_host(const uint8_t *img, int W, int H, size_t pitch, int x0, int y0, int x1, int y1,
                                    int32_t out_rgb[3]) {
    LeanCtx *cx = nullptr;
    if (int rc = acquire_lean(&cx, "masked_median_rgb_host")) return rc;
    const uint8_t *d = nullptr;
    size_t dp = 0;
    if (int rc = upload_image(*cx, img, W, H, pitch, &d, &dp, "masked_median_rgb_host")) return rc;
    return b200comp_masked_median_rgb(d, W, H, dp, x0, y0, x1, y1, out_rgb, cx->st);
}

int b200comp_edge_strip_medians_host(const uint8_t *img, int W, int H, size_t pitch, int strip_px,
                                     int32_t out_edges[12]) {
    if (strip_px < 1 || !out_edges) return b200comp_set_error_(B200COMP_EINVAL, "edge_strip_medians_host: bad argument");
    LeanCtx *cx = nullptr;
    if (int rc = acquire_lean(&cx, "edge_strip_medians_host")) return rc;
    const uint8_t *d = nullptr;
    size_t dp = 0;
    if (int rc = upload_image(*cx, img, W, H, pitch, &d, &dp, "edge_strip_medians_host")) return rc;
    return edge_medians_dev(d, dp, W, H, strip_px, out_edges, cx->st);
}

// canvas of the thread's staging set, filled on its stream by `fill`, copied out to the caller
extern "C++" template <typename Fill>
int fill_and_download(LeanCtx &cx, uint8_t *out, int W, int H, size_t out_pitch, const char *who, Fill fill) {
    const size_t dp = align_up((size_t)W * 4, 16), bytes = dp * H;
    if (!LeanCtx::grow_pinned(&cx.pin_out, &cx.pin_out_cap, bytes) || !LeanCtx::grow_device(&cx.d_out, &cx.d_out_cap, bytes, cx.st))
        return b200comp_set_error_(B200COMP_ENOMEM, (std::string(who) + ": staging allocation failed").c_str());
    if (int rc = fill(cx.d_out, dp)) return rc;
    return download_rows(cx, out, out_pitch, cx.d_out, dp, (size_t)W * 4, H, who);
}

int b200comp_fill_solid_host(const uint8_t *bg, int Wb, int Hb, size_t bg_pitch, uint8_t *out, int W, int H,
                             size_t out_pitch, int32_t out_rgb[3]) {
    if (!out || W < 1 || H < 1 || out_pitch < (size_t)W * 4)
        return b200comp_set_error_(B200COMP_EINVAL, "fill_solid_host: bad canvas argument");
    LeanCtx *cx = nullptr;
    if (int rc = acquire_lean(&cx, "fill_solid_host")) return rc;
    const uint8_t *d_bg = nullptr;
    size_t bp = 0;
    if (int rc = upload_image(*cx, bg, Wb, Hb, bg_pitch, &d_bg, &bp, "fill_solid_host")) return rc;
    int32_t rgb[3];
    if (int rc = b200comp_masked_median_rgb(d_bg, Wb, Hb, bp, 0, 0, Wb, Hb, rgb, cx->st)) return rc;
    const uint32_t rgba = (uint32_t)rgb[0] | ((uint32_t)rgb[1] << 8) | ((uint32_t)rgb[2] << 16) | 0xff000000u;
    cudaStream_t st = cx->st;
    int rc = fill_and_download(*cx, out, W, H, out_pitch, "fill_solid_host",
                               [&](uint8_t *d, size_t dp) { return b200comp_fill_rgba(d, W, H, dp, rgba, st); });
    if (rc) return rc;
    if (out_rgb) std::memcpy(out_rgb, rgb, sizeof rgb);
    return 0;
}

int b200comp_fill_gradient_host(const uint8_t *bg, int Wb, int Hb, size_t bg_pitch, uint8_t *out, int W, int H,
                                size_t out_pitch, int strip_px, int32_t out_edges[12], int *out_horizontal) {
    if (!out || W < 1 || H < 1 || strip_px < 1 || out_pitch < (size_t)W * 4)
        return b200comp_set_error_(B200COMP_EINVAL, "fill_gradient_host: bad argument");
    LeanCtx *cx = nullptr;
    if (int rc = acquire_lean(&cx, "fill_gradient_host")) return rc;
    const uint8_t *d_bg = nullptr;
    size_t bp = 0;
    if (int rc = upload_image(*cx, bg, Wb, Hb, bg_pitch, &d_bg, &bp, "fill_gradient_host")) return rc;
    int32_t edges[12];
    if (int rc = edge_medians_dev(d_bg, bp, Wb, Hb, strip_px, edges, cx->st)) return rc;
    // _axis_variance (:58-60) and the direction choice (:69-80): squared colour distance, ties -> horizontal
    auto dist = [&](int a, int b) {
        double s = 0;
        for (int c = 0; c < 3; ++c) {
            const double d = (double)edges[3 * a + c] - (double)edges[3 * b + c];
            s += d * d;
        }
        return s;
    };
    const int horizontal = dist(0, 1) <= dist(2, 3) ? 1 : 0;
    const int32_t *c1 = horizontal ? edges : edges + 6;
    const int32_t *c2 = horizontal ? edges + 3 : edges + 9;
    cudaStream_t st = cx->st;
    int rc = fill_and_download(*cx, out, W, H, out_pitch, "fill_gradient_host", [&](uint8_t *d, size_t dp) {
        return b200comp_fill_gradient(d, W, H, dp, horizontal, c1, c2, st);
    });
    if (rc) return rc;
    if (out_edges) std::memcpy(out_edges, edges, sizeof edges);
    if (out_horizontal) *out_horizontal = horizontal;
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
