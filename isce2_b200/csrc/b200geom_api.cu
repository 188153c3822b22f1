// b200geom_api.cu -- host side of the C ABI declared in include/b200geom.h.
//
// Owns device memory, streams and events; prepares the per-scene constants the reference keeps in Fortran
// module globals; launches the kernels of topo_kernels.cu / geo2rdr_kernels.cu.  No CPU fallback: every
// compute entry point fails with B200_ENODEVICE when no CUDA device is usable.
#include "../../include/b200geom.h"

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <cerrno>
#include <mutex>
#include <sched.h>
#include <unistd.h>
#include <shared_mutex>
#include <new>
#include <memory>
#include <thread>
#include <vector>

#include "geo2rdr_kernels.cuh"
#include "geozero_kernels.cuh"
#include "orbit_poly.h"
#include "post_kernels.cuh"
#include "resamp_kernels.cuh"
#include "topo_kernels.cuh"

using namespace b2;

namespace {

int fail(char *err, size_t errlen, int code, const char *fmt, ...)
{
    if (err && errlen) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(err, errlen, B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int select_device(int device, char *err, size_t errlen)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(err, errlen, B200_ENODEVICE,
                    "no CUDA device available (%s); libb200geom has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(err, errlen, B200_ENODEVICE, "device %d out of range [0,%d)", device, n);
    CK(cudaSetDevice(device));
    return B200_OK;
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}


// -------------------------------------------------------------------------------------------------
// Device workspace cache.  A topo + geo2rdr pass over a swath needs ~20 GB of layers; cudaMalloc / cudaFree of that
// costs more than the kernels once a stack of bursts or dates is processed call after call.  Freed buffers are kept
// per device and handed out again to requests of similar size; b200_release_cached_memory() (or a failed cudaMalloc)
// returns them to the driver.  B200_NO_CACHE=1 disables the cache.
// -------------------------------------------------------------------------------------------------
struct CacheBlock {
    void *ptr;
    size_t size;
    int device;
    bool free;
};
std::mutex g_cache_mu;
std::vector<CacheBlock> g_cache;

void cache_release_locked(int device /* -1: all */)
{
    for (size_t i = 0; i < g_cache.size();) {
        if (g_cache[i].free && (device < 0 || g_cache[i].device == device)) {
            cudaSetDevice(g_cache[i].device);
            cudaFree(g_cache[i].ptr);
            g_cache[i] = g_cache.back();
            g_cache.pop_back();
        } else {
            i++;
        }
    }
}

cudaError_t dmalloc(void **p, size_t bytes)
{
    static const bool no_cache = getenv("B200_NO_CACHE") != nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (bytes == 0) bytes = 1;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (!no_cache) {
        int best = -1;
        for (size_t i = 0; i < g_cache.size(); i++) {
            const CacheBlock &b = g_cache[i];
            if (b.free && b.device == dev && b.size >= bytes && b.size <= bytes + bytes / 4 + 4096 &&
                (best < 0 || b.size < g_cache[best].size))
                best = (int)i;
        }
        if (best >= 0) {
            g_cache[best].free = false;
            *p = g_cache[best].ptr;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { // give cached memory back and retry once
        cudaGetLastError();
        cache_release_locked(dev);
        cudaSetDevice(dev);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess && !no_cache) g_cache.push_back(CacheBlock{*p, bytes, dev, false});
    return e;
}

void dfree(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (CacheBlock &b : g_cache)
        if (b.ptr == p) {
            b.free = true;
            return;
        }
    cudaFree(p); // not ours (cache disabled)
}

template <typename T>
cudaError_t dmalloc(T **p, size_t bytes)
{
    return dmalloc(reinterpret_cast<void **>(p), bytes);
}


// -------------------------------------------------------------------------------------------------
// HostSink: device -> host copies of the output layers, whatever kind of host memory the caller handed over.
//
// Page-locked destinations (b200_alloc_pinned, cudaHostRegister) take the plain cudaMemcpyAsync path: the DMA engine
// writes them directly, asynchronously, at PCIe speed.  PAGEABLE destinations -- what the reference's callers have: a
// numpy.memmap over the .rdr / .off file being written (Topozero.py:274-302) -- would make cudaMemcpyAsync synchronous and
// single threaded: the driver bounces every 64 KB through its own small staging buffer and one CPU thread pays all the
// page faults of a file that does not exist yet.  For those the sink bounces through its own ring of page-locked slots
// and a pool of copier threads that memcpy finished slots into the destination in parallel (page faults of a fresh
// mapping included), while the DMA engine fills the next slots.  B200_COPY_THREADS (default: the CPUs of the
// process, at most 16) sizes the pool; B200_COPY_THREADS=0 disables the bounce path (plain cudaMemcpyAsync for everything).
// -------------------------------------------------------------------------------------------------
constexpr size_t kSinkSlotBytes = 32u << 20; // one bounce slot
constexpr int kSinkSlotsMax = 24;
// Slots of the ring of one sink (32 MB each, page-locked, cached between calls); B200_SINK_SLOTS overrides (2 .. 24).
// Sixteen: a slot bound for a registered file is written by ONE pwrite, so the slots in flight are the writers in flight
// (Component path, 16.5 GB swath on tmpfs: 1.29 s with 6 slots, 0.91 s with 16, 0.89 s with 24).
int sink_slots()
{
    static const int n = [] {
        int v = 16;
        if (const char *e = getenv("B200_SINK_SLOTS")) v = atoi(e);
        return v < 2 ? 2 : (v > kSinkSlotsMax ? kSinkSlotsMax : v);
    }();
    return n;
}
// Parts a slot bound for a registered file is written in (B200_FILE_PARTS).  One: writers of neighbouring pieces of one
// file mostly wait for each other (4 parts: 1.16 s against 0.91 s), whole slots mostly belong to different rasters.
int file_parts()
{
    static const int n = [] {
        int v = 1;
        if (const char *e = getenv("B200_FILE_PARTS")) v = atoi(e);
        return v < 1 ? 1 : (v > 64 ? 64 : v);
    }();
    return n;
}

struct SinkRing {
    char *buf[kSinkSlotsMax] = {};
    cudaEvent_t ev[kSinkSlotsMax] = {}; // events belong to the device that was current when they were created
    int device = -1;
    bool ok = false;
};
std::mutex g_ring_mu;
std::vector<SinkRing *> g_free_rings; // rings of finished calls, handed to the next one (pinning 512 MB costs ~0.1 s)

SinkRing *ring_acquire(int device)
{
    {
        std::lock_guard<std::mutex> lk(g_ring_mu);
        for (size_t i = 0; i < g_free_rings.size(); i++)
            if (g_free_rings[i]->device == device) {
                SinkRing *r = g_free_rings[i];
                g_free_rings[i] = g_free_rings.back();
                g_free_rings.pop_back();
                return r;
            }
    }
    SinkRing *r = new (std::nothrow) SinkRing;
    if (!r) return nullptr;
    r->device = device;
    r->ok = true;
    for (int i = 0; i < sink_slots() && r->ok; i++) {
        r->ok = cudaHostAlloc((void **)&r->buf[i], kSinkSlotBytes, cudaHostAllocPortable) == cudaSuccess &&
                cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!r->ok) {
        cudaGetLastError();
        for (int i = 0; i < sink_slots(); i++) {
            if (r->buf[i]) cudaFreeHost(r->buf[i]);
            if (r->ev[i]) cudaEventDestroy(r->ev[i]);
        }
        delete r;
        return nullptr;
    }
    return r;
}
void ring_release(SinkRing *r)
{
    std::lock_guard<std::mutex> lk(g_ring_mu);
    g_free_rings.push_back(r);
}

int sink_threads()
{
    static const int n = [] {
        if (const char *e = getenv("B200_COPY_THREADS")) return atoi(e) < 0 ? 0 : (atoi(e) > 64 ? 64 : atoi(e));
        // every CPU this process may run on, at most 16: the copies are page-fault bound and scale with threads up to
        // there (tmpfs, 16.5 GB of fresh rasters: 2.9 s with 4, 2.1 s with 8, 1.8 s with 16 -- profiles/r02_component_threads_sweep.log)
        int t = 0;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) t = CPU_COUNT(&set);
        if (t <= 0) t = (int)std::thread::hardware_concurrency();
        return t < 2 ? 2 : (t > 16 ? 16 : t);
    }();
    return n;
}

// File-backed destinations (b200_host_file_register): a range of host addresses that is a shared mapping of a file.  The
// copier threads write results bound for such a range with pwrite on the file instead of storing through the mapping:
// the page cache is filled without one page fault (and one zeroed page) per 4 KB of a raster that does not exist yet; the
// mapping sees the same pages.  What matters is the shape of the writes (profiles/r02_component_file_sweep.log, 16.5 GB
// swath through the Components on tmpfs): a slot cut into one piece per thread, as the stores are, is SLOWER than the
// stores (2.6 s against 1.8 s: writers of neighbouring pieces of one file wait for each other), one pwrite per 32 MB slot
// with 16 slots in flight is twice as fast (0.9 s).
struct FileRange {
    const char *base;
    size_t bytes;
    int fd; // our own duplicate of the caller's descriptor
    long long off;
};
std::shared_mutex g_file_mu;
std::vector<FileRange> g_files;
std::atomic<unsigned long long> g_file_bytes{0};      // written with pwrite so far (b200_host_file_bytes)
std::atomic<unsigned long long> g_file_bytes_read{0}; // read with pread so far (b200_host_file_bytes_read)
bool file_lookup(const void *dst, size_t n, int &fd, long long &off)
{
    std::shared_lock<std::shared_mutex> lk(g_file_mu);
    const char *p = (const char *)dst;
    for (const FileRange &r : g_files)
        if (p >= r.base && p + n <= r.base + r.bytes) {
            fd = r.fd;
            off = r.off + (long long)(p - r.base);
            return true;
        }
    return false;
}
// pwrite of the whole part; false (nothing more written) on the first error other than EINTR
bool pwrite_all(int fd, const char *src, size_t n, long long off, size_t &done)
{
    done = 0;
    while (done < n) {
        const ssize_t w = pwrite(fd, src + done, n - done, (off_t)(off + (long long)done));
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        if (w == 0) return false;
        done += (size_t)w;
    }
    return true;
}

// pread of the whole part; false on the first error other than EINTR or at end of file
bool pread_all(int fd, char *dst, size_t n, long long off, size_t &done)
{
    done = 0;
    while (done < n) {
        const ssize_t r = pread(fd, dst + done, n - done, (off_t)(off + (long long)done));
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        if (r == 0) return false;
        done += (size_t)r;
    }
    return true;
}

// process-wide pool of copier threads: a task is one memcpy (or pwrite / pread) of a part of a slot
class CopyPool {
  public:
    struct Task {
        cudaEvent_t ready;         // the slot's D2H has landed once this event completed
        int device;
        void *dst;
        const void *src;
        size_t bytes;
        std::atomic<int> *pending; // per slot: parts still to copy; 0 == slot reusable
        std::mutex *mu;            // owner's mutex / cv, signalled when pending reaches 0
        std::condition_variable *cv;
        int fd = -1;               // >= 0: dst (or, with from_file, src) lies in a registered file mapping: write (read) the
        long long foff = 0;        //       file at offset foff instead
        bool from_file = false;
    };
    static CopyPool &get()
    {
        static CopyPool *p = new CopyPool(sink_threads()); // leaked on purpose: threads may outlive static destructors
        return *p;
    }
    void push(const Task &t)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(t);
        }
        cv_.notify_one();
    }
    int size() const { return n_; }

  private:
    explicit CopyPool(int n) : n_(n)
    {
        for (int i = 0; i < n; i++) std::thread([this] { run(); }).detach();
    }
    void run()
    {
        int cur = -1;
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !q_.empty(); });
                t = q_.front();
                q_.pop_front();
            }
            if (cur != t.device) {
                cudaSetDevice(t.device);
                cur = t.device;
            }
            if (t.ready) cudaEventSynchronize(t.ready); // (host -> slot copies of HostSource have nothing to wait for)
            size_t done = 0;
            if (t.fd >= 0 && t.from_file) {
                pread_all(t.fd, (char *)t.dst, t.bytes, t.foff, done);
                g_file_bytes_read.fetch_add(done, std::memory_order_relaxed);
            } else if (t.fd >= 0) {
                pwrite_all(t.fd, (const char *)t.src, t.bytes, t.foff, done);
                g_file_bytes.fetch_add(done, std::memory_order_relaxed);
            }
            if (done < t.bytes) memcpy((char *)t.dst + done, (const char *)t.src + done, t.bytes - done); // no file / write failed

            if (t.pending->fetch_sub(1) == 1) {
                std::lock_guard<std::mutex> lk(*t.mu);
                t.cv->notify_all();
            }
        }
    }
    int n_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Task> q_;
};

class HostSink {
  public:
    HostSink(cudaStream_t s, int device) : s_(s), device_(device) {}
    ~HostSink()
    {
        finish();
        if (ring_) ring_release(ring_);
    }
    // enqueue `bytes` from device memory `src` to host memory `dst` behind everything already on the stream
    cudaError_t copy(void *dst, const void *src, size_t bytes)
    {
        if (!bytes) return cudaSuccess;
        if (!pageable(dst)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s_);
        if (!ring_) {
            ring_ = ring_acquire(device_);
            if (!ring_) { // no page-locked memory to bounce through: the driver's own path still works
                bounce_failed_ = true;
                return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s_);
            }
            for (int i = 0; i < sink_slots(); i++) pending_[i].store(0);
        }
        const int nt = CopyPool::get().size();
        int fd = -1;
        long long foff = 0;
        if (!file_lookup(dst, bytes, fd, foff)) fd = -1;
        for (size_t o = 0; o < bytes; o += kSinkSlotBytes) {
            const size_t n = bytes - o < kSinkSlotBytes ? bytes - o : kSinkSlotBytes;
            const int k = next_;
            next_ = (next_ + 1) % sink_slots();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return pending_[k].load() == 0; });
            }
            cudaError_t e = cudaMemcpyAsync(ring_->buf[k], (const char *)src + o, n, cudaMemcpyDeviceToHost, s_);
            if (e == cudaSuccess) e = cudaEventRecord(ring_->ev[k], s_);
            if (e != cudaSuccess) return e;
            // parts of >= 1 MB, page aligned within the slot, one per copier thread (file-backed: file_parts() per slot)
            const int np = fd >= 0 ? file_parts() : nt;
            size_t part = (n + np - 1) / np;
            part = (part + 4095) & ~(size_t)4095;
            if (part < (1u << 20)) part = 1u << 20;
            const int nparts = (int)((n + part - 1) / part);
            pending_[k].store(nparts);
            for (int q = 0; q < nparts; q++) {
                const size_t po = (size_t)q * part, pn = n - po < part ? n - po : part;
                CopyPool::get().push(CopyPool::Task{ring_->ev[k], device_, (char *)dst + o + po, ring_->buf[k] + po, pn, &pending_[k],
                                                    &mu_, &cv_, fd, foff + (long long)(o + po)});
            }
        }
        return cudaSuccess;
    }
    // all bounced bytes are in their destination (plain copies still need the caller's stream synchronize)
    void finish()
    {
        if (!ring_) return;
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
            for (int i = 0; i < sink_slots(); i++)
                if (pending_[i].load() != 0) return false;
            return true;
        });
    }
  private:
    bool pageable(const void *p)
    {
        if (sink_threads() == 0 || bounce_failed_) return false;
        if (p == last_ptr_) return last_pageable_; // the layers are asked about chunk after chunk
        // same allocation as a pointer seen before?  cudaPointerGetAttributes costs ~1 us; there are <= 10 distinct buffers
        cudaPointerAttributes a;
        bool pg = false;
        if (cudaPointerGetAttributes(&a, p) == cudaSuccess) pg = (a.type == cudaMemoryTypeUnregistered);
        else cudaGetLastError();
        last_ptr_ = p;
        last_pageable_ = pg;
        return pg;
    }
    cudaStream_t s_;
    int device_;
    SinkRing *ring_ = nullptr;
    std::atomic<int> pending_[kSinkSlotsMax];
    std::mutex mu_;
    std::condition_variable cv_;
    int next_ = 0;
    bool bounce_failed_ = false;
    const void *last_ptr_ = nullptr;
    bool last_pageable_ = false;
};


// HostSource: the mirror image for INPUTS that live in pageable memory (the lat / lon / hgt rasters of geo2rdr arrive as
// numpy.memmaps of the .rdr files topo wrote, Geo2rdr.py:208-226).  The copier threads fill a page-locked slot from the
// source in parallel (page-cache reads and page faults included), the DMA engine uploads it while they fill the next one.
class HostSource {
  public:
    HostSource(cudaStream_t s, int device) : s_(s), device_(device) {}
    ~HostSource()
    {
        if (ring_) {
            for (int i = 0; i < sink_slots(); i++)
                if (used_[i]) cudaEventSynchronize(ring_->ev[i]); // the slots must not be reused while a DMA still reads them
            ring_release(ring_);
        }
    }
    // enqueue `bytes` from host memory `src` to device memory `dst` on the stream
    cudaError_t copy(void *dst, const void *src, size_t bytes)
    {
        if (!bytes) return cudaSuccess;
        if (!pageable(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s_);
        if (!ring_) {
            ring_ = ring_acquire(device_);
            if (!ring_) {
                failed_ = true;
                return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s_);
            }
        }
        const int nt = CopyPool::get().size();
        int fd = -1; // the source is a registered file mapping (the .rdr rasters geo2rdr reads): pread instead of page faults
        long long foff = 0;
        if (!file_lookup(src, bytes, fd, foff)) fd = -1;
        for (size_t o = 0; o < bytes; o += kSinkSlotBytes) {
            const size_t n = bytes - o < kSinkSlotBytes ? bytes - o : kSinkSlotBytes;
            const int k = next_;
            next_ = (next_ + 1) % sink_slots();
            if (used_[k]) {
                cudaError_t e = cudaEventSynchronize(ring_->ev[k]); // the slot's previous upload has left it
                if (e != cudaSuccess) return e;
            }
            size_t part = (n + nt - 1) / nt;
            part = (part + 4095) & ~(size_t)4095;
            if (part < (1u << 20)) part = 1u << 20;
            const int nparts = (int)((n + part - 1) / part);
            pending_.store(nparts);
            for (int q = 0; q < nparts; q++) {
                const size_t po = (size_t)q * part, pn = n - po < part ? n - po : part;
                CopyPool::get().push(CopyPool::Task{nullptr, device_, ring_->buf[k] + po, (const char *)src + o + po, pn, &pending_, &mu_, &cv_,
                                                    fd, foff + (long long)(o + po), true});
            }
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return pending_.load() == 0; });
            }
            cudaError_t e = cudaMemcpyAsync((char *)dst + o, ring_->buf[k], n, cudaMemcpyHostToDevice, s_);
            if (e == cudaSuccess) e = cudaEventRecord(ring_->ev[k], s_);
            if (e != cudaSuccess) return e;
            used_[k] = true;
        }
        return cudaSuccess;
    }

  private:
    bool pageable(const void *p)
    {
        if (sink_threads() == 0 || failed_) return false;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) == cudaSuccess) return a.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        return false;
    }
    cudaStream_t s_;
    int device_;
    SinkRing *ring_ = nullptr;
    bool used_[kSinkSlotsMax] = {};
    std::atomic<int> pending_{0};
    std::mutex mu_;
    std::condition_variable cv_;
    int next_ = 0;
    bool failed_ = false;
};

// lines per pipeline chunk: ~8 Mpixel, so that copies of one chunk overlap the kernels of the next
int chunk_lines(int width, int nlines)
{
    long long c = 8000000LL / (width > 0 ? width : 1);
    if (c < 32) c = 32;
    if (c > nlines) c = nlines;
    return (int)c;
}

struct DeviceOrbit {
    double *buf = nullptr;
    OrbitView view{0, nullptr, nullptr, nullptr};
};

int upload_orbit(const b200_orbit *o, DeviceOrbit &d, cudaStream_t s, char *err, size_t errlen)
{
    const int n = o->nvec;
    std::vector<double> h((size_t)n * 7);
    memcpy(h.data(), o->t, sizeof(double) * n);
    memcpy(h.data() + n, o->pos, sizeof(double) * 3 * n);
    memcpy(h.data() + 4 * (size_t)n, o->vel, sizeof(double) * 3 * n);
    CK(dmalloc(&d.buf, sizeof(double) * 7 * (size_t)n));
    CK(cudaMemcpyAsync(d.buf, h.data(), sizeof(double) * 7 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s)); // h goes out of scope
    d.view = OrbitView{n, d.buf, d.buf + n, d.buf + 4 * (size_t)n};
    return B200_OK;
}

int check_orbit(const b200_orbit *o, int method, char *err, size_t errlen)
{
    if (!o || !o->t || !o->pos || !o->vel) return fail(err, errlen, B200_EINVAL, "orbit is NULL");
    if (method != B200_ORBIT_HERMITE && method != B200_ORBIT_SCH && method != B200_ORBIT_LEGENDRE)
        return fail(err, errlen, B200_EINVAL, "Undefined orbit interpolation method.");
    // topozero.f90:104-131 / geo2rdr.f90:64-91
    if (method == B200_ORBIT_LEGENDRE && o->nvec < 9)
        return fail(err, errlen, B200_EORBIT, "Need atleast 9 state vectors for using legendre polynomial interpolation");
    if (method == B200_ORBIT_HERMITE && o->nvec < 4)
        return fail(err, errlen, B200_EORBIT, "Need atleast 4 state vectors for using hermite polynomial interpolation");
    if (method == B200_ORBIT_SCH && o->nvec < 4)
        return fail(err, errlen, B200_EORBIT, "Need atleast 4 state vectors for using SCH interpolation");
    return B200_OK;
}

int fill_poly2d(const b200_poly2d *src, Poly2dDev &dst, const char *what, char *err, size_t errlen)
{
    if (!src || !src->coeffs) return fail(err, errlen, B200_EINVAL, "%s polynomial is NULL", what);
    const int n = (src->range_order + 1) * (src->azimuth_order + 1);
    if (src->range_order < 0 || src->azimuth_order < 0 || n > kMaxPoly2dCoeffs)
        return fail(err, errlen, B200_EINVAL, "%s polynomial has %d coefficients (max %d)", what, n, kMaxPoly2dCoeffs);
    dst.range_order = src->range_order;
    dst.azimuth_order = src->azimuth_order;
    dst.mean_range = src->mean_range;
    dst.mean_azimuth = src->mean_azimuth;
    dst.norm_range = src->norm_range;
    dst.norm_azimuth = src->norm_azimuth;
    dst.inv_norm_range = 1.0 / src->norm_range;
    dst.inv_norm_azimuth = 1.0 / src->norm_azimuth;
    memset(dst.c, 0, sizeof dst.c);
    memcpy(dst.c, src->coeffs, sizeof(double) * n);
    return B200_OK;
}

} // namespace

// =================================================================================================
// topozero
// =================================================================================================
struct b200_topo_plan {
    b200_topo_params p{};
    int line0 = 0, nlines = 0;
    TopoConst C{};
    DeviceOrbit orb;
    float *d_dem = nullptr;
    float *d_sinc = nullptr;
    double *d_dem64 = nullptr;
    void *d_raw = nullptr;
    int *d_maxkey = nullptr;
    double *d_rho = nullptr;  // block rows of the slant-range image
    double *d_rho0 = nullptr; // its first row (bbox stage)
    LineState *d_states = nullptr;
    TopoLayers layers{};
    TopoStats *d_stats = nullptr;
    MaskScratch scr{};
    int mask_grid = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr, evs = nullptr;
    float ms_pixels = 0.f, ms_mask = 0.f, ms_solve = 0.f;
    int dem_x0 = 0, dem_y0 = 0;
    float dem_max = 0.f;
    float ms_setup = 0.f, ms_kernels = 0.f;
    int launches = 0;
    bool executed = false;

    ~b200_topo_plan()
    {
        cudaSetDevice(p.device);
        dfree(orb.buf);
        dfree(d_dem);
        dfree(d_sinc);
        dfree(d_dem64);
        dfree(d_raw);
        dfree(d_maxkey);
        dfree(d_rho);
        dfree(d_rho0);
        dfree(d_states);
        dfree(layers.lat);
        dfree(layers.lon);
        dfree(layers.hgt);
        dfree(layers.los);
        dfree(layers.inc);
        dfree(layers.mask);
        dfree(layers.ctrack);
        dfree(layers.elev);
        dfree(d_stats);
        dfree(scr.orng_sorted);
        dfree(scr.pm);
        dfree(scr.sm);
        dfree(scr.rank);
        dfree(scr.cs);
        dfree(scr.lats);
        dfree(scr.lons);
        dfree(scr.orng);
        dfree(scr.ctr_sorted);
        dfree(scr.oflag);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (evm) cudaEventDestroy(evm);
        if (evs) cudaEventDestroy(evs);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int topo_plan_build(b200_topo_plan *pl, const void *dem, int dem_dtype, const b200_orbit *orbit,
                           const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image, int want_los,
                           int want_inc, int want_mask, char *err, size_t errlen)
{
    const b200_topo_params &p = pl->p;
    int rc;
    if (p.width < 2 || p.length < 1) return fail(err, errlen, B200_EINVAL, "bad radar grid %d x %d", p.length, p.width);
    if (!dem) return fail(err, errlen, B200_EINVAL, "dem is NULL");
    if (dem_dtype != B200_DEM_F32 && dem_dtype != B200_DEM_I16) return fail(err, errlen, B200_EINVAL, "bad dem_dtype %d", dem_dtype);
    if (p.dem_method < B200_DEM_SINC || p.dem_method > B200_DEM_BIQUINTIC)
        return fail(err, errlen, B200_EINVAL, "Undefined interpolation method."); // topozero.f90:96-99
    if ((rc = check_orbit(orbit, p.orbit_method, err, errlen)) != B200_OK) return rc;
    if (!slrng && !rho_image)
        return fail(err, errlen, B200_EINVAL, "Both the slant range accessor and starting range are zero"); // topozero.f90:156-159
    if (p.prf <= 0 || p.nazlooks < 1 || p.nrnglooks < 1) return fail(err, errlen, B200_EINVAL, "bad prf / looks");

    pl->line0 = p.line0 < 0 ? 0 : p.line0;
    pl->nlines = (p.nlines < 0 || pl->line0 + p.nlines > p.length) ? p.length - pl->line0 : p.nlines;
    if (pl->nlines <= 0) return fail(err, errlen, B200_EINVAL, "empty line block (line0=%d nlines=%d length=%d)", p.line0, p.nlines, p.length);

    if ((rc = select_device(p.device, err, errlen)) != B200_OK) return rc;
    CK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&pl->ev0));
    CK(cudaEventCreate(&pl->ev1));
    CK(cudaEventCreate(&pl->evm));
    CK(cudaEventCreate(&pl->evs));
    cudaStream_t s = pl->stream;
    CK(cudaEventRecord(pl->ev0, s));

    // ---- constants ----
    TopoConst &C = pl->C;
    C.elp = make_ellipsoid(p.major, p.e2);
    C.wvl = p.wvl;
    C.thresh = p.thresh;
    C.ilrl = p.look_side;
    C.numiter = p.numiter;
    C.extraiter = p.extraiter;
    C.deltalat = p.delta_lat;
    C.deltalon = p.delta_lon;
    C.method = p.dem_method;
    C.width = p.width;
    C.length = p.length;
    C.nazlooks = p.nazlooks;
    C.t0 = p.t0;
    C.prf = p.prf;
    C.peghdg = p.peg_heading;
    C.pi = 4.0 * atan(1.0); // fortranUtils.f90:38-41
    C.r2d = 180.0 / C.pi;
    C.inv_r2d = 1.0 / C.r2d;
    C.inv_dlat = 1.0 / p.delta_lat;
    C.inv_dlon = 1.0 / p.delta_lon;
    C.orbit_method = p.orbit_method;
    if ((rc = fill_poly2d(dop, C.dop, "doppler", err, errlen)) != B200_OK) return rc;
    if (slrng) {
        if ((rc = fill_poly2d(slrng, C.slr, "slant range", err, errlen)) != B200_OK) return rc;
    } else {
        memset(&C.slr, 0, sizeof C.slr);
        C.slr.norm_range = C.slr.norm_azimuth = C.slr.inv_norm_range = C.slr.inv_norm_azimuth = 1.0;
    }
    spline6_make_table(C.spl);

    if ((rc = upload_orbit(orbit, pl->orb, s, err, errlen)) != B200_OK) return rc;

    if (rho_image) {
        const size_t w = (size_t)p.width;
        CK(dmalloc(&pl->d_rho0, sizeof(double) * w));
        CK(cudaMemcpyAsync(pl->d_rho0, rho_image, sizeof(double) * w, cudaMemcpyHostToDevice, s));
        CK(dmalloc(&pl->d_rho, sizeof(double) * w * pl->nlines));
        CK(cudaMemcpyAsync(pl->d_rho, rho_image + (size_t)pl->line0 * w, sizeof(double) * w * pl->nlines,
                           cudaMemcpyHostToDevice, s));
    }

    // ---- bbox of interest (topozero.f90:192-263) ----
    double *d_bbox = nullptr;
    double h_bbox[24];
    CK(dmalloc(&d_bbox, sizeof h_bbox));
    {
        TopoConst Cb = C;
        Cb.dem = DemView{nullptr, 0, 0, nullptr, nullptr, 0};
        Cb.rho_image = pl->d_rho0; // row 1 of the image, or NULL -> polynomial
        launch_topo_bbox(Cb, pl->orb.view, d_bbox, s);
        pl->launches++;
    }
    cudaError_t ce = cudaMemcpyAsync(h_bbox, d_bbox, sizeof h_bbox, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    dfree(d_bbox);
    if (ce != cudaSuccess) return fail(err, errlen, B200_ECUDA, "bbox kernel failed: %s", cudaGetErrorString(ce));
    double min_lat = 10000., max_lat = -10000., min_lon = 10000., max_lon = -10000.;
    int nok = 0;
    for (int i = 0; i < 8; i++) {
        if (h_bbox[3 * i + 2] == 0.0) continue;
        nok++;
        min_lat = fmin(min_lat, h_bbox[3 * i]);
        max_lat = fmax(max_lat, h_bbox[3 * i]);
        min_lon = fmin(min_lon, h_bbox[3 * i + 1]);
        max_lon = fmax(max_lon, h_bbox[3 * i + 1]);
    }
    if (nok == 0 || !(min_lat <= max_lat) || !(min_lon <= max_lon))
        return fail(err, errlen, B200_EORBIT, "Error getting statevector for bounds computation");
    // reference angles for the in-kernel trigonometry (geom_device.cuh: RefAngle): centre of the corner points;
    // used only when every corner lies well inside the series' range, otherwise the kernels use libm
    {
        const double d2r = C.pi / 180.0;
        const double clat = 0.5 * (min_lat + max_lat) * d2r, clon = 0.5 * (min_lon + max_lon) * d2r;
        C.ref.lat = make_ref_angle(clat);
        C.ref.lon = make_ref_angle(clon);
        const double elat = fmax(fabs(max_lat * d2r - C.ref.lat.a0), fabs(min_lat * d2r - C.ref.lat.a0));
        const double elon = fmax(fabs(max_lon * d2r - C.ref.lon.a0), fabs(min_lon * d2r - C.ref.lon.a0));
        // 0.03 rad of slack: pixels inside the scene can bulge past the corner box, and the DEM-edge clamps
        // never move a point by more than the 0.15 deg margin
        C.ref.use_ref = (elat + 0.03 <= C.ref.lat.dmax && elon + 0.03 <= C.ref.lon.dmax) ? 1 : 0;
        if (getenv("B200_FORCE_LIBM_TRIG")) C.ref.use_ref = 0; // developer switch for A/B measurements
        // narrow blocks (a Sentinel-1 swath, a NISAR frame) take the truncated series; 0.008 rad of slack covers
        // the 0.15 deg DEM margin (0.0026 rad) and the bulge of the footprint past its corner box (~1e-4 rad)
        C.ref.lat.narrow = (elat + 0.008 <= kNarrowAngle) ? 1 : 0;
        C.ref.lon.narrow = (elon + 0.008 <= kNarrowAngle) ? 1 : 0;
        if (getenv("B200_FORCE_LONG_SERIES")) C.ref.lat.narrow = C.ref.lon.narrow = 0;
    }
    const double MARGIN = 0.15; // topozeroState.f:74-75
    min_lon -= MARGIN; max_lon += MARGIN; min_lat -= MARGIN; max_lat += MARGIN;

    // ---- usable part of the DEM (topozero.f90:281-320), typo at :314 kept ----
    const double firstlon = p.first_lon, firstlat = p.first_lat, deltalon = p.delta_lon, deltalat = p.delta_lat;
    const int idemwidth = p.dem_width, idemlength = p.dem_length;
    double umin_lon = fmax(min_lon, firstlon);
    double umax_lon = fmin(max_lon, firstlon + (idemwidth - 1) * deltalon);
    double umax_lat = fmin(max_lat, firstlat);
    double umin_lat = fmax(min_lat, firstlat + (idemlength - 1) * deltalat);
    int ustartx = (int)((umin_lon - firstlon) / deltalon) + 1;
    if (ustartx < 1) ustartx = 1;
    int uendx = (int)((umax_lon - firstlon) / deltalon + 0.5) + 1;
    if (uendx > idemwidth) uendx = idemwidth;
    int ustarty = (int)((umax_lat - firstlat) / deltalat) + 1;
    if (ustarty < 1) ustarty = 1;
    int uendy = (int)((umin_lat - firstlat) / deltalat + 0.5) + 1;
    if (uendy > idemlength) ustarty = idemlength;
    const int udemwidth = uendx - ustartx + 1, udemlength = uendy - ustarty + 1;
    if (udemwidth < 2 || udemlength < 2 || uendy > idemlength || ustartx > idemwidth || ustarty > idemlength)
        return fail(err, errlen, B200_EDEM, "DEM does not cover the scene: needs lon [%f,%f] lat [%f,%f]", min_lon, max_lon,
                    min_lat, max_lat);
    C.ufirstlon = firstlon + deltalon * (ustartx - 1);
    C.ufirstlat = firstlat + deltalat * (ustarty - 1);
    pl->dem_x0 = ustartx;
    pl->dem_y0 = ustarty;

    // ---- crop + float32 conversion + demmax (topozero.f90:333-345) ----
    const size_t ncell = (size_t)udemwidth * (size_t)udemlength;
    const size_t esz = dem_dtype == B200_DEM_I16 ? 2 : 4;
    CK(dmalloc(&pl->d_dem, sizeof(float) * ncell));
    CK(dmalloc(&pl->d_maxkey, sizeof(int)));
    const char *src = (const char *)dem + ((size_t)(ustarty - 1) * (size_t)idemwidth + (size_t)(ustartx - 1)) * esz;
    void *dst = pl->d_dem;
    if (dem_dtype == B200_DEM_I16) {
        CK(dmalloc(&pl->d_raw, esz * ncell));
        dst = pl->d_raw;
    }
    CK(cudaMemcpy2DAsync(dst, (size_t)udemwidth * esz, src, (size_t)idemwidth * esz, (size_t)udemwidth * esz,
                         (size_t)udemlength, cudaMemcpyHostToDevice, s));
    {
        int init = (int)0x80000000; // smaller than every order key
        CK(cudaMemcpyAsync(pl->d_maxkey, &init, sizeof init, cudaMemcpyHostToDevice, s));
        launch_dem_prepare(dst, dem_dtype, pl->d_dem, ncell, pl->d_maxkey, s);
        pl->launches++;
        int key = 0;
        CK(cudaMemcpyAsync(&key, pl->d_maxkey, sizeof key, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        pl->dem_max = dem_max_decode(key);
    }
    if (pl->d_raw) {
        dfree(pl->d_raw);
        pl->d_raw = nullptr;
    }
    if (p.dem_method == B200_DEM_SINC) { // prepareMethods (topozeroMethods.f:46-63)
        std::vector<float> tab((size_t)kSincSub * kSincLen);
        sinc_make_table(tab.data());
        CK(dmalloc(&pl->d_sinc, sizeof(float) * tab.size()));
        CK(cudaMemcpyAsync(pl->d_sinc, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    int stride64 = 0;
    if (p.dem_method == B200_DEM_BIQUINTIC) { // padded double copy for the 6x6 spline window (DemView::d64)
        stride64 = udemwidth + 1;
        CK(dmalloc(&pl->d_dem64, sizeof(double) * (size_t)(udemlength + 1) * (size_t)stride64));
        launch_dem_pad64(pl->d_dem, udemwidth, udemlength, pl->d_dem64, stride64, s);
        pl->launches++;
    }
    C.dem = DemView{pl->d_dem, udemwidth, udemlength, pl->d_sinc, pl->d_dem64, stride64};
    C.rho_image = pl->d_rho ? pl->d_rho - (size_t)pl->line0 * (size_t)p.width : nullptr;

    // ---- per-line state ----
    CK(dmalloc(&pl->d_states, sizeof(LineState) * (size_t)pl->nlines));
    launch_line_setup(C, pl->orb.view, pl->line0, pl->nlines, pl->d_states, s);
    pl->launches++;

    // ---- resident output layers ----
    const size_t npix = (size_t)pl->nlines * (size_t)p.width;
    CK(dmalloc(&pl->layers.lat, sizeof(double) * npix));
    CK(dmalloc(&pl->layers.lon, sizeof(double) * npix));
    CK(dmalloc(&pl->layers.hgt, sizeof(double) * npix));
    if (want_los) CK(dmalloc(&pl->layers.los, sizeof(float) * 2 * npix));
    if (want_inc) CK(dmalloc(&pl->layers.inc, sizeof(float) * 2 * npix));
    CK(dmalloc(&pl->layers.ctrack, sizeof(double) * npix)); // SCH height between solve and final pass, then ctrack
    if (want_mask) {
        CK(dmalloc(&pl->layers.mask, npix));
        CK(dmalloc(&pl->layers.elev, sizeof(float) * npix));
        const int ow = 2 * p.width + 1;
        const int g = mask_grid_size(pl->nlines);
        pl->mask_grid = g;
        CK(dmalloc(&pl->scr.cs, sizeof(double) * (size_t)g * p.width));
        CK(dmalloc(&pl->scr.lats, sizeof(double) * (size_t)g * p.width));
        CK(dmalloc(&pl->scr.lons, sizeof(double) * (size_t)g * p.width));
        CK(dmalloc(&pl->scr.orng, sizeof(double) * (size_t)g * ow));
        CK(dmalloc(&pl->scr.ctr_sorted, sizeof(double) * (size_t)g * ow));
        CK(dmalloc(&pl->scr.orng_sorted, sizeof(double) * (size_t)g * ow));
        CK(dmalloc(&pl->scr.pm, sizeof(double) * (size_t)g * ow));
        CK(dmalloc(&pl->scr.sm, sizeof(double) * (size_t)g * ow));
        CK(dmalloc(&pl->scr.rank, sizeof(int) * (size_t)g * ow));
        CK(dmalloc(&pl->scr.oflag, (size_t)g * ow));
    }
    CK(dmalloc(&pl->d_stats, sizeof(TopoStats)));
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&pl->ms_setup, pl->ev0, pl->ev1));
    return B200_OK;
}

extern "C" int b200_topo_plan_create(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                                     const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                                     int want_los, int want_inc, int want_mask, b200_topo_plan **plan, char *err,
                                     size_t errlen)
{
    if (!p || !plan) return fail(err, errlen, B200_EINVAL, "params/plan is NULL");
    *plan = nullptr;
    b200_topo_plan *pl = new (std::nothrow) b200_topo_plan;
    if (!pl) return fail(err, errlen, B200_ENOMEM, "out of host memory");
    pl->p = *p;
    int rc = topo_plan_build(pl, dem, dem_dtype, orbit, dop, slrng, rho_image, want_los, want_inc, want_mask, err, errlen);
    if (rc != B200_OK) {
        delete pl;
        return rc;
    }
    *plan = pl;
    return B200_OK;
}

extern "C" int b200_topo_plan_execute(b200_topo_plan *pl, float *ms_kernels, char *err, size_t errlen)
{
    if (!pl) return fail(err, errlen, B200_EINVAL, "plan is NULL");
    CK(cudaSetDevice(pl->p.device));
    cudaStream_t s = pl->stream;
    TopoStats init;
    init.min_lat = init.min_lon = 0x7fffffffffffffffLL;
    init.max_lat = init.max_lon = (long long)0x8000000000000000ULL;
    init.converged = init.iterations = 0;
    CK(cudaMemcpyAsync(pl->d_stats, &init, sizeof init, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(pl->ev0, s));
    if (launch_topo_pixels(pl->C, pl->d_states, pl->line0, pl->nlines, pl->layers, pl->d_stats, s, pl->evs) != 0)
        return fail(err, errlen, B200_EINVAL, "cannot launch the pixel kernel (method %d, %d lines)", pl->C.method, pl->nlines);
    int launches = topo_pixel_launches(pl->C.method);
    CK(cudaEventRecord(pl->evm, s));
    if (pl->layers.mask) {
        if (launch_topo_mask(pl->C, pl->d_states, pl->line0, pl->nlines, pl->layers, pl->dem_max, pl->scr, pl->mask_grid, s) != 0)
            return fail(err, errlen, B200_EINVAL, "cannot launch the mask kernel");
        launches++;
    }
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&pl->ms_kernels, pl->ev0, pl->ev1));
    CK(cudaEventElapsedTime(&pl->ms_pixels, pl->ev0, pl->evm));
    CK(cudaEventElapsedTime(&pl->ms_solve, pl->ev0, pl->evs));
    CK(cudaEventElapsedTime(&pl->ms_mask, pl->evm, pl->ev1));
    if (!pl->executed) pl->launches += launches;
    pl->executed = true;
    if (ms_kernels) *ms_kernels = pl->ms_kernels;
    return B200_OK;
}

extern "C" int b200_topo_plan_fetch(b200_topo_plan *pl, const b200_topo_outputs *out, b200_topo_result *res, char *err,
                                    size_t errlen)
{
    if (!pl) return fail(err, errlen, B200_EINVAL, "plan is NULL");
    if (!pl->executed) return fail(err, errlen, B200_EINVAL, "plan was not executed");
    CK(cudaSetDevice(pl->p.device));
    cudaStream_t s = pl->stream;
    const size_t npix = (size_t)pl->nlines * (size_t)pl->p.width;
    if (out) {
        HostSink sink(s, pl->p.device);
        if (out->lat) CK(sink.copy(out->lat, pl->layers.lat, sizeof(double) * npix));
        if (out->lon) CK(sink.copy(out->lon, pl->layers.lon, sizeof(double) * npix));
        if (out->hgt) CK(sink.copy(out->hgt, pl->layers.hgt, sizeof(double) * npix));
        if (out->los && pl->layers.los) CK(sink.copy(out->los, pl->layers.los, sizeof(float) * 2 * npix));
        if (out->inc && pl->layers.inc) CK(sink.copy(out->inc, pl->layers.inc, sizeof(float) * 2 * npix));
        if (out->mask && pl->layers.mask) CK(sink.copy(out->mask, pl->layers.mask, npix));
        sink.finish();
    }
    TopoStats st;
    CK(cudaMemcpyAsync(&st, pl->d_stats, sizeof st, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (res) {
        res->min_lat = stats_decode(st.min_lat);
        res->max_lat = stats_decode(st.max_lat);
        res->min_lon = stats_decode(st.min_lon);
        res->max_lon = stats_decode(st.max_lon);
        res->converged = (long long)st.converged;
        res->iterations = (long long)st.iterations;
        res->dem_x0 = pl->dem_x0;
        res->dem_y0 = pl->dem_y0;
        res->dem_nx = pl->C.dem.nx;
        res->dem_ny = pl->C.dem.ny;
        res->dem_max = pl->dem_max;
        res->ms_setup = pl->ms_setup;
        res->ms_kernels = pl->ms_kernels;
        res->ms_pixels = pl->ms_pixels;
        res->ms_solve = pl->ms_solve;
        res->ms_mask = pl->ms_mask;
        res->ms_total = 0.f;
        res->gpu_launches = pl->launches;
    }
    return B200_OK;
}

extern "C" int b200_topo_plan_device_layers(b200_topo_plan *pl, const double **lat, const double **lon, const double **hgt)
{
    if (!pl) return B200_EINVAL;
    if (lat) *lat = pl->layers.lat;
    if (lon) *lon = pl->layers.lon;
    if (hgt) *hgt = pl->layers.hgt;
    return B200_OK;
}

extern "C" void b200_topo_plan_destroy(b200_topo_plan *pl) { delete pl; }

// =================================================================================================
// geo2rdr
// =================================================================================================
struct b200_geo_plan {
    b200_geo_params p{};
    int line0 = 0, nlines = 0;
    double *d_lat = nullptr, *d_lon = nullptr, *d_hgt = nullptr;
    bool owns_inputs = true;
    void *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    int out_f32 = 1;
    GeoStats *d_stats = nullptr;
    GeoMid *d_mid = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float ms_setup = 0.f, ms_kernels = 0.f;
    int launches = 0;
    bool executed = false;
    int want[4] = {0, 0, 0, 0};
    // frozen geometry (b200_geo_plan_freeze_geometry): ECEF copy of lat / lon / hgt for the given ellipsoid
    double *d_xyz[3] = {nullptr, nullptr, nullptr};
    double xyz_major = 0.0, xyz_e2 = 0.0;

    ~b200_geo_plan()
    {
        cudaSetDevice(p.device);
        for (double *q : d_xyz) dfree(q);
        if (owns_inputs) {
            dfree(d_lat);
            dfree(d_lon);
            dfree(d_hgt);
        }
        for (void *q : d_out) dfree(q);
        dfree(d_stats);
        dfree(d_mid);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int geo_plan_common(b200_geo_plan *pl, char *err, size_t errlen)
{
    const b200_geo_params &p = pl->p;
    if (p.dem_width < 1 || p.dem_length < 1) return fail(err, errlen, B200_EINVAL, "bad lat/lon/hgt grid %d x %d", p.dem_length, p.dem_width);
    pl->line0 = p.line0 < 0 ? 0 : p.line0;
    pl->nlines = (p.nlines < 0 || pl->line0 + p.nlines > p.dem_length) ? p.dem_length - pl->line0 : p.nlines;
    if (pl->nlines <= 0) return fail(err, errlen, B200_EINVAL, "empty line block");
    int rc = select_device(p.device, err, errlen);
    if (rc != B200_OK) return rc;
    CK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&pl->ev0));
    CK(cudaEventCreate(&pl->ev1));
    CK(dmalloc(&pl->d_stats, sizeof(GeoStats)));
    CK(dmalloc(&pl->d_mid, sizeof(GeoMid)));
    return B200_OK;
}

extern "C" int b200_geo_plan_create(const b200_geo_params *p, const double *lat, const double *lon, const double *hgt,
                                    b200_geo_plan **plan, char *err, size_t errlen)
{
    if (!p || !plan) return fail(err, errlen, B200_EINVAL, "params/plan is NULL");
    *plan = nullptr;
    if (!lat || !lon || !hgt) return fail(err, errlen, B200_EINVAL, "lat/lon/hgt is NULL");
    b200_geo_plan *pl = new (std::nothrow) b200_geo_plan;
    if (!pl) return fail(err, errlen, B200_ENOMEM, "out of host memory");
    pl->p = *p;
    int rc = geo_plan_common(pl, err, errlen);
    if (rc == B200_OK) {
        const size_t npix = (size_t)pl->nlines * (size_t)p->dem_width, off = (size_t)pl->line0 * (size_t)p->dem_width;
        HostSource source(pl->stream, p->device); // the geometry may come as memmaps of the reference date's .rdr files
        auto up = [&](double **d, const double *h) -> int {
            CK(dmalloc(d, sizeof(double) * npix));
            CK(source.copy(*d, h + off, sizeof(double) * npix));
            return B200_OK;
        };
        cudaEventRecord(pl->ev0, pl->stream);
        rc = up(&pl->d_lat, lat);
        if (rc == B200_OK) rc = up(&pl->d_lon, lon);
        if (rc == B200_OK) rc = up(&pl->d_hgt, hgt);
        if (rc == B200_OK) {
            cudaEventRecord(pl->ev1, pl->stream);
            cudaError_t e = cudaStreamSynchronize(pl->stream);
            if (e != cudaSuccess) rc = fail(err, errlen, B200_ECUDA, "upload failed: %s", cudaGetErrorString(e));
            else cudaEventElapsedTime(&pl->ms_setup, pl->ev0, pl->ev1);
        }
    }
    if (rc != B200_OK) {
        delete pl;
        return rc;
    }
    *plan = pl;
    return B200_OK;
}

extern "C" int b200_geo_plan_create_from_topo(const b200_geo_params *p, b200_topo_plan *topo, b200_geo_plan **plan,
                                              char *err, size_t errlen)
{
    if (!p || !plan || !topo) return fail(err, errlen, B200_EINVAL, "params/plan/topo is NULL");
    *plan = nullptr;
    if (!topo->executed) return fail(err, errlen, B200_EINVAL, "topo plan was not executed");
    if (p->device != topo->p.device || p->dem_width != topo->p.width)
        return fail(err, errlen, B200_EINVAL, "geo2rdr plan must live on the topo plan's device and share its width");
    b200_geo_plan *pl = new (std::nothrow) b200_geo_plan;
    if (!pl) return fail(err, errlen, B200_ENOMEM, "out of host memory");
    pl->p = *p;
    // the borrowed layers hold the topo plan's block of lines
    pl->p.line0 = topo->line0;
    pl->p.nlines = topo->nlines;
    int rc = geo_plan_common(pl, err, errlen);
    if (rc != B200_OK) {
        delete pl;
        return rc;
    }
    pl->owns_inputs = false;
    pl->d_lat = topo->layers.lat;
    pl->d_lon = topo->layers.lon;
    pl->d_hgt = topo->layers.hgt;
    *plan = pl;
    return B200_OK;
}

namespace {

// Everything one orbit / Doppler / radar-grid combination needs on the device (geo2rdr.f90:118-208)
struct GeoRun {
    GeoConst C{};
    DeviceOrbit dorb;
    double *d_poly = nullptr;
    OrbitPolyView op{};
    bool use_poly = false;
    ~GeoRun()
    {
        dfree(dorb.buf);
        dfree(d_poly);
    }
};

int geo_prepare(b200_geo_plan *pl, const b200_geo_params &p, const b200_orbit *orbit, const b200_poly1d *dop, GeoRun &R,
                char *err, size_t errlen)
{
    int rc;
    if ((rc = check_orbit(orbit, p.orbit_method, err, errlen)) != B200_OK) return rc;
    if (!dop || !dop->coeffs || dop->order < 0 || dop->order + 1 > kMaxPoly1dCoeffs)
        return fail(err, errlen, B200_EINVAL, "bad doppler polynomial");
    if (p.dem_width != pl->p.dem_width) return fail(err, errlen, B200_EINVAL, "dem_width differs from the plan");
    if (p.prf <= 0 || p.nazlooks < 1 || p.nrnglooks < 1) return fail(err, errlen, B200_EINVAL, "bad prf / looks");
    cudaStream_t s = pl->stream;

    // ---- scalars of geo2rdr.f90:118-133 ----
    GeoConst &C = R.C;
    C.elp = make_ellipsoid(p.major, p.e2);
    C.wvl = p.wvl;
    C.tstart = p.t0;
    C.dtaz = p.nazlooks / p.prf;
    C.tend = p.t0 + (p.length - 1) * C.dtaz;
    C.tmid = 0.5 * (C.tstart + C.tend);
    C.rngstart = p.rho0;
    C.dmrg = p.nrnglooks * p.drho;
    C.rngend = p.rho0 + (p.width - 1) * C.dmrg;
    C.orbit_method = p.orbit_method;
    C.bistatic = p.bistatic;
    C.demwidth = p.dem_width;
    const double pi = 4.0 * atan(1.0);
    C.deg2rad = pi / 180.0;
    C.sol = 299792458.0; // fortranUtils.f90:43-46
    C.xyz_in = 0;
    // ---- doppler-vs-range polynomial and its derivative (:161-189) ----
    C.fd.order = dop->order;
    C.fd.mean = p.rho0 + dop->mean * p.drho;
    C.fd.norm = dop->norm * p.drho;
    for (int k = 0; k <= dop->order; k++) C.fd.c[k] = dop->coeffs[k] * p.prf;
    if (C.fd.order == 0) {
        C.fdd.order = 0;
        C.fdd.mean = 0.0;
        C.fdd.norm = 1.0;
        C.fdd.c[0] = 0.0;
    } else {
        C.fdd.order = C.fd.order - 1;
        C.fdd.mean = C.fd.mean;
        C.fdd.norm = C.fd.norm;
        for (int k = 1; k <= dop->order; k++) C.fdd.c[k - 1] = k * C.fd.c[k] / C.fd.norm;
    }
    if ((rc = upload_orbit(orbit, R.dorb, s, err, errlen)) != B200_OK) return rc;

    // ---- mid-scene state (:194-208) ----
    launch_geo_setup(p.orbit_method, R.dorb.view, C.tmid, pl->d_mid, s);
    GeoMid mid;
    CK(cudaMemcpyAsync(&mid, pl->d_mid, sizeof mid, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (mid.stat_mid != 0) return fail(err, errlen, B200_EORBIT, "Cannot interpolate orbits at the center of scene.");
    if (mid.stat_acc != 0) return fail(err, errlen, B200_EORBIT, "Cannot compute acceleration at the center of scene.");
    C.xyz_mid = Vec3{mid.xyz[0], mid.xyz[1], mid.xyz[2]};
    C.vel_mid = Vec3{mid.vel[0], mid.vel[1], mid.vel[2]};
    C.acc_mid = Vec3{mid.acc[0], mid.acc[1], mid.acc[2]};

    R.use_poly = (p.orbit_method == B200_ORBIT_HERMITE || p.orbit_method == B200_ORBIT_LEGENDRE);
    if (R.use_poly) {
        // The reference marks a pixel invalid as soon as ANY of its 9-11 fixed-point iterates leaves the state-vector span
        // (geo2rdr.f90:287-291).  Its iterates approach the solution from tmid with a ratio of ~0.1 per step, so those of a
        // pixel that ends inside the acquisition window stay within ~0.1 x the half-duration of it: when the state
        // vectors cover the window with the margin below, the span test can only hit pixels the bounds test (:308-316)
        // rejects anyway and the Newton kernel (which sees other iterates) decides validity identically.  When they do
        // not -- an orbit cut within a fraction of a second of the scene -- the kernel that performs the reference's own
        // iteration runs instead.
        const double half = 0.5 * (C.tend - C.tstart);
        const double margin = 0.25 * half + 0.05;
        if (orbit->t[0] > C.tstart - margin || orbit->t[orbit->nvec - 1] < C.tend + margin) R.use_poly = false;
    }
    if (R.use_poly) { // Newton solve on per-window orbit polynomials (orbit_poly.h)
        HostOrbitPoly hp;
        if (!build_orbit_poly(p.orbit_method, orbit->nvec, orbit->t, orbit->pos, orbit->vel, hp))
            return fail(err, errlen, B200_EORBIT, "cannot build the orbit polynomials");
        const size_t nt = (size_t)hp.n, nw = (size_t)hp.nwin, nc = hp.cp.size(), nv = hp.cv.size();
        std::vector<double> blob(nt + 2 * nw + nc + nv);
        memcpy(blob.data(), orbit->t, sizeof(double) * nt);
        memcpy(blob.data() + nt, hp.tc.data(), sizeof(double) * nw);
        memcpy(blob.data() + nt + nw, hp.inv_h.data(), sizeof(double) * nw);
        memcpy(blob.data() + nt + 2 * nw, hp.cp.data(), sizeof(double) * nc);
        if (nv) memcpy(blob.data() + nt + 2 * nw + nc, hp.cv.data(), sizeof(double) * nv);
        CK(dmalloc(&R.d_poly, sizeof(double) * blob.size()));
        CK(cudaMemcpyAsync(R.d_poly, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s)); // blob goes out of scope
        R.op = OrbitPolyView{hp.method, hp.n, hp.nwin, hp.ncoef, R.d_poly, R.d_poly + nt, R.d_poly + nt + nw,
                             R.d_poly + nt + 2 * nw, nv ? R.d_poly + nt + 2 * nw + nc : nullptr};
    }
    return B200_OK;
}

int geo_launch(const GeoRun &R, int line0_abs, int nlines, const GeoLayers &L, int out_f32, GeoStats *stats, cudaStream_t s)
{
    if (R.use_poly) return launch_geo2rdr_poly(R.C, R.op, line0_abs, nlines, L, out_f32, stats, s);
    return launch_geo2rdr(R.C, R.dorb.view, line0_abs, nlines, L, out_f32, stats, s);
}

int geo_alloc_outputs(b200_geo_plan *pl, const b200_geo_params &p, const int want[4], char *err, size_t errlen)
{
    const size_t npix = (size_t)pl->nlines * (size_t)p.dem_width;
    const size_t esz = p.out_f32 ? 4 : 8;
    for (int i = 0; i < 4; i++) {
        if (pl->d_out[i] && (pl->out_f32 != p.out_f32 || !want[i])) {
            dfree(pl->d_out[i]);
            pl->d_out[i] = nullptr;
        }
        if (want[i] && !pl->d_out[i]) CK(dmalloc(&pl->d_out[i], esz * npix));
        pl->want[i] = want[i];
    }
    pl->out_f32 = p.out_f32;
    return B200_OK;
}

} // namespace

extern "C" int b200_geo_plan_execute(b200_geo_plan *pl, const b200_geo_params *pp, const b200_orbit *orbit,
                                     const b200_poly1d *dop, int want_azt, int want_rgm, int want_azoff, int want_rgoff,
                                     float *ms_kernels, char *err, size_t errlen)
{
    if (!pl || !pp) return fail(err, errlen, B200_EINVAL, "plan/params is NULL");
    const b200_geo_params &p = *pp;
    CK(cudaSetDevice(pl->p.device));
    cudaStream_t s = pl->stream;
    CK(cudaEventRecord(pl->ev0, s)); // ms_kernels covers the mid-scene setup kernel and the solve
    GeoRun R;
    int rc = geo_prepare(pl, p, orbit, dop, R, err, errlen);
    if (rc != B200_OK) return rc;
    const int want[4] = {want_azt, want_rgm, want_azoff, want_rgoff};
    if ((rc = geo_alloc_outputs(pl, p, want, err, errlen)) != B200_OK) return rc;
    GeoLayers L{pl->d_lat, pl->d_lon, pl->d_hgt, pl->d_out[0], pl->d_out[1], pl->d_out[2], pl->d_out[3]};
    if (pl->d_xyz[0] && pl->xyz_major == p.major && pl->xyz_e2 == p.e2) { // frozen geometry of this ellipsoid
        R.C.xyz_in = 1;
        L.lat = pl->d_xyz[0];
        L.lon = pl->d_xyz[1];
        L.hgt = pl->d_xyz[2];
    }
    CK(cudaMemsetAsync(pl->d_stats, 0, sizeof(GeoStats), s));
    if (geo_launch(R, pl->line0, pl->nlines, L, p.out_f32, pl->d_stats, s) != 0)
        return fail(err, errlen, B200_EINVAL, "cannot launch the geo2rdr kernel");
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&pl->ms_kernels, pl->ev0, pl->ev1));
    pl->launches = 2;
    pl->executed = true;
    if (ms_kernels) *ms_kernels = pl->ms_kernels;
    return B200_OK;
}

extern "C" int b200_geo_plan_freeze_geometry(b200_geo_plan *pl, double major, double e2, char *err, size_t errlen)
{
    if (!pl) return fail(err, errlen, B200_EINVAL, "plan is NULL");
    if (!(major > 0.0) || !(e2 >= 0.0 && e2 < 1.0)) return fail(err, errlen, B200_EINVAL, "bad ellipsoid (a = %g, e2 = %g)", major, e2);
    CK(cudaSetDevice(pl->p.device));
    const size_t npix = (size_t)pl->nlines * (size_t)pl->p.dem_width;
    for (int i = 0; i < 3; i++)
        if (!pl->d_xyz[i]) CK(dmalloc(&pl->d_xyz[i], sizeof(double) * npix));
    GeoConst C{};
    C.elp = make_ellipsoid(major, e2);
    C.deg2rad = 4.0 * atan(1.0) / 180.0;
    launch_llh_to_xyz(C, pl->d_lat, pl->d_lon, pl->d_hgt, pl->d_xyz[0], pl->d_xyz[1], pl->d_xyz[2], npix, pl->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(pl->stream));
    pl->xyz_major = major;
    pl->xyz_e2 = e2;
    return B200_OK;
}

extern "C" int b200_geo_plan_fetch(b200_geo_plan *pl, const b200_geo_outputs *out, b200_geo_result *res, char *err,
                                   size_t errlen)
{
    if (!pl) return fail(err, errlen, B200_EINVAL, "plan is NULL");
    if (!pl->executed) return fail(err, errlen, B200_EINVAL, "plan was not executed");
    CK(cudaSetDevice(pl->p.device));
    cudaStream_t s = pl->stream;
    const size_t bytes = (size_t)pl->nlines * (size_t)pl->p.dem_width * (pl->out_f32 ? 4 : 8);
    if (out) {
        HostSink sink(s, pl->p.device);
        void *h[4] = {out->azt, out->rgm, out->azoff, out->rgoff};
        for (int i = 0; i < 4; i++)
            if (h[i] && pl->d_out[i]) CK(sink.copy(h[i], pl->d_out[i], bytes));
        sink.finish();
    }
    GeoStats st;
    CK(cudaMemcpyAsync(&st, pl->d_stats, sizeof st, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (res) {
        res->num_outside = (long long)st.outside;
        res->num_valid = (long long)st.valid;
        res->num_converged = (long long)st.converged;
        res->iterations = (long long)st.iterations;
        res->ms_setup = pl->ms_setup;
        res->ms_kernels = pl->ms_kernels;
        res->ms_total = 0.f;
        res->gpu_launches = pl->launches;
    }
    return B200_OK;
}

extern "C" void b200_geo_plan_destroy(b200_geo_plan *pl) { delete pl; }

extern "C" int b200_geo2rdr_run(const b200_geo_params *p, const double *lat, const double *lon, const double *hgt,
                                const b200_orbit *orbit, const b200_poly1d *dop, const b200_geo_outputs *out,
                                b200_geo_result *res, char *err, size_t errlen)
{
    if (!out || (!out->azt && !out->rgm && !out->azoff && !out->rgoff))
        return fail(err, errlen, B200_EINVAL, "No outputs requested from geo2rdr. Check again."); // Geo2rdr.py:271-274
    if (!p) return fail(err, errlen, B200_EINVAL, "params is NULL");
    if (!lat || !lon || !hgt) return fail(err, errlen, B200_EINVAL, "lat/lon/hgt is NULL");
    const double t0 = now_ms();
    b200_geo_plan plan;
    b200_geo_plan *pl = &plan;
    pl->p = *p;
    int rc = geo_plan_common(pl, err, errlen);
    if (rc != B200_OK) return rc;
    const size_t w = (size_t)p->dem_width, npix = (size_t)pl->nlines * w, off = (size_t)pl->line0 * w;
    CK(dmalloc(&pl->d_lat, sizeof(double) * npix));
    CK(dmalloc(&pl->d_lon, sizeof(double) * npix));
    CK(dmalloc(&pl->d_hgt, sizeof(double) * npix));
    cudaStream_t s = pl->stream;
    CK(cudaEventRecord(pl->ev0, s));
    GeoRun R;
    if ((rc = geo_prepare(pl, *p, orbit, dop, R, err, errlen)) != B200_OK) return rc;
    const int want[4] = {out->azt != nullptr, out->rgm != nullptr, out->azoff != nullptr, out->rgoff != nullptr};
    if ((rc = geo_alloc_outputs(pl, *p, want, err, errlen)) != B200_OK) return rc;
    CK(cudaMemsetAsync(pl->d_stats, 0, sizeof(GeoStats), s));

    // ---- three-stage pipeline over blocks of lines: H2D(c+1) | kernel(c) | D2H(c-1) ----
    struct Streams {
        cudaStream_t h = nullptr, d = nullptr;
        std::vector<cudaEvent_t> ev;
        ~Streams()
        {
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
            if (h) cudaStreamDestroy(h);
            if (d) cudaStreamDestroy(d);
        }
    } st;
    CK(cudaStreamCreateWithFlags(&st.h, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&st.d, cudaStreamNonBlocking));
    const int cl = chunk_lines(p->dem_width, pl->nlines);
    const size_t esz = p->out_f32 ? 4 : 8;
    void *hout[4] = {out->azt, out->rgm, out->azoff, out->rgoff};
    int launches = 1;
    HostSink sink(st.d, p->device);
    HostSource source(st.h, p->device); // lat / lon / hgt may be memmaps of the .rdr files
    for (int c0 = 0; c0 < pl->nlines; c0 += cl) {
        const int n = (c0 + cl <= pl->nlines) ? cl : pl->nlines - c0;
        const size_t o = (size_t)c0 * w, cnt = (size_t)n * w;
        cudaEvent_t eh, ek;
        CK(cudaEventCreateWithFlags(&eh, cudaEventDisableTiming));
        st.ev.push_back(eh);
        CK(cudaEventCreateWithFlags(&ek, cudaEventDisableTiming));
        st.ev.push_back(ek);
        CK(source.copy(pl->d_lat + o, lat + off + o, sizeof(double) * cnt));
        CK(source.copy(pl->d_lon + o, lon + off + o, sizeof(double) * cnt));
        CK(source.copy(pl->d_hgt + o, hgt + off + o, sizeof(double) * cnt));
        CK(cudaEventRecord(eh, st.h));
        CK(cudaStreamWaitEvent(s, eh, 0));
        GeoLayers L{pl->d_lat + o, pl->d_lon + o, pl->d_hgt + o, nullptr, nullptr, nullptr, nullptr};
        void **lo[4] = {&L.azt, &L.rgm, &L.azoff, &L.rgoff};
        for (int i = 0; i < 4; i++)
            if (pl->d_out[i]) *lo[i] = (char *)pl->d_out[i] + o * esz;
        if (geo_launch(R, pl->line0 + c0, n, L, p->out_f32, pl->d_stats, s) != 0)
            return fail(err, errlen, B200_EINVAL, "cannot launch the geo2rdr kernel");
        launches++;
        CK(cudaEventRecord(ek, s));
        CK(cudaStreamWaitEvent(st.d, ek, 0));
        for (int i = 0; i < 4; i++)
            if (hout[i] && pl->d_out[i])
                CK(sink.copy((char *)hout[i] + o * esz, (char *)pl->d_out[i] + o * esz, cnt * esz));
    }
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(st.d));
    sink.finish();
    CK(cudaEventElapsedTime(&pl->ms_kernels, pl->ev0, pl->ev1));
    pl->launches = launches;
    pl->executed = true;
    rc = b200_geo_plan_fetch(pl, nullptr, res, err, errlen);
    if (rc == B200_OK && res) res->ms_total = (float)(now_ms() - t0);
    return rc;
}

// =================================================================================================
// the topo verb (host buffers), optionally with geo2rdr jobs fused behind every block of lines
// =================================================================================================
static int topo_run_impl(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                         const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                         const b200_topo_outputs *out, b200_topo_result *res, int njobs, const b200_geo_job *jobs, char *err,
                         size_t errlen)
{
    if (!out || !out->lat || !out->lon || !out->hgt)
        return fail(err, errlen, B200_EINVAL, "lat/lon/hgt output buffers are mandatory (Topozero.py:274-302)");
    const double t0 = now_ms();
    b200_topo_plan *pl = nullptr;
    int rc = b200_topo_plan_create(p, dem, dem_dtype, orbit, dop, slrng, rho_image, out->los != nullptr, out->inc != nullptr,
                                   out->mask != nullptr, &pl, err, errlen);
    if (rc != B200_OK) return rc;
    struct PlanGuard {
        b200_topo_plan *p;
        ~PlanGuard() { delete p; }
    } pg{pl};
    // ---- two-stage pipeline over blocks of lines: kernels(c+1) | D2H(c) ----
    struct Streams {
        cudaStream_t d = nullptr;
        std::vector<cudaEvent_t> ev;
        ~Streams()
        {
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
            if (d) cudaStreamDestroy(d);
        }
    } st;
    CK(cudaStreamCreateWithFlags(&st.d, cudaStreamNonBlocking));
    cudaStream_t s = pl->stream;
    // ---- fused geo2rdr jobs: each borrows the resident lat / lon / hgt of this block (no trip through the host) ----
    struct Job {
        std::unique_ptr<b200_geo_plan> gp;
        std::unique_ptr<GeoRun> R;
        b200_geo_params p{};
        void *hout[4] = {nullptr, nullptr, nullptr, nullptr};
        size_t esz = 4;
    };
    std::vector<Job> fused((size_t)(njobs > 0 ? njobs : 0));
    for (int j = 0; j < njobs; j++) {
        const b200_geo_job &jb = jobs[j];
        Job &J = fused[j];
        if (!jb.p || !jb.out || (!jb.out->azt && !jb.out->rgm && !jb.out->azoff && !jb.out->rgoff))
            return fail(err, errlen, B200_EINVAL, "No outputs requested from geo2rdr. Check again. (job %d)", j);
        if (jb.p->dem_width != pl->p.width || jb.p->dem_length != pl->p.length)
            return fail(err, errlen, B200_EINVAL, "fused geo2rdr job %d: lat/lon/hgt grid %d x %d is not the topo grid %d x %d", j,
                        jb.p->dem_length, jb.p->dem_width, pl->p.length, pl->p.width);
        J.p = *jb.p;
        J.p.line0 = pl->line0;
        J.p.nlines = pl->nlines;
        J.p.device = pl->p.device;
        J.gp.reset(new (std::nothrow) b200_geo_plan);
        J.R.reset(new (std::nothrow) GeoRun);
        if (!J.gp || !J.R) return fail(err, errlen, B200_ENOMEM, "out of host memory");
        J.gp->p = J.p;
        J.gp->owns_inputs = false;
        if ((rc = geo_plan_common(J.gp.get(), err, errlen)) != B200_OK) return rc;
        J.gp->d_lat = pl->layers.lat;
        J.gp->d_lon = pl->layers.lon;
        J.gp->d_hgt = pl->layers.hgt;
        if ((rc = geo_prepare(J.gp.get(), J.p, jb.orbit, jb.dop, *J.R, err, errlen)) != B200_OK) return rc;
        const int want[4] = {jb.out->azt != nullptr, jb.out->rgm != nullptr, jb.out->azoff != nullptr, jb.out->rgoff != nullptr};
        if ((rc = geo_alloc_outputs(J.gp.get(), J.p, want, err, errlen)) != B200_OK) return rc;
        CK(cudaMemsetAsync(J.gp->d_stats, 0, sizeof(GeoStats), J.gp->stream));
        CK(cudaStreamSynchronize(J.gp->stream));
        J.hout[0] = jb.out->azt; J.hout[1] = jb.out->rgm; J.hout[2] = jb.out->azoff; J.hout[3] = jb.out->rgoff;
        J.esz = J.p.out_f32 ? 4 : 8;
        J.gp->launches = 1;
    }
    TopoStats init;
    init.min_lat = init.min_lon = 0x7fffffffffffffffLL;
    init.max_lat = init.max_lon = (long long)0x8000000000000000ULL;
    init.converged = init.iterations = 0;
    CK(cudaMemcpyAsync(pl->d_stats, &init, sizeof init, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(pl->ev0, s));
    const size_t w = (size_t)pl->p.width;
    const int cl = chunk_lines(pl->p.width, pl->nlines);
    int launches = 0;
    // several jobs on one ellipsoid (the stack shape): the ECEF coordinates of a block are formed once for all of them
    struct XyzScratch {
        double *p[3] = {nullptr, nullptr, nullptr};
        ~XyzScratch()
        {
            for (double *q : p) dfree(q);
        }
    } xs;
    bool share_xyz = fused.size() >= 2;
    for (size_t j = 1; j < fused.size() && share_xyz; j++)
        share_xyz = fused[j].R->use_poly && fused[0].R->use_poly && fused[j].p.major == fused[0].p.major && fused[j].p.e2 == fused[0].p.e2;
    if (share_xyz) {
        const size_t cpx = (size_t)(cl < pl->nlines ? cl : pl->nlines) * w;
        for (int i = 0; i < 3; i++) CK(dmalloc(&xs.p[i], sizeof(double) * cpx));
        for (Job &J : fused) J.R->C.xyz_in = 1;
    }
    auto launch_chunk = [&](int c0, int n, cudaEvent_t done) -> int {
        const size_t o = (size_t)c0 * w;
        TopoLayers L = pl->layers;
        L.lat += o; L.lon += o; L.hgt += o; L.ctrack += o;
        if (L.los) L.los += 2 * o;
        if (L.inc) L.inc += 2 * o;
        if (L.mask) L.mask += o;
        if (L.elev) L.elev += o;
        if (launch_topo_pixels(pl->C, pl->d_states + c0, pl->line0 + c0, n, L, pl->d_stats, s) != 0) return -1;
        launches += topo_pixel_launches(pl->C.method);
        if (L.mask) {
            const int g = pl->mask_grid < n ? pl->mask_grid : n;
            if (launch_topo_mask(pl->C, pl->d_states + c0, pl->line0 + c0, n, L, pl->dem_max, pl->scr, g, s) != 0) return -1;
            launches++;
        }
        if (share_xyz) {
            launch_llh_to_xyz(fused[0].R->C, pl->layers.lat + o, pl->layers.lon + o, pl->layers.hgt + o, xs.p[0], xs.p[1], xs.p[2],
                              (size_t)n * w, s);
            launches++;
        }
        for (Job &J : fused) { // same stream: the block's lat / lon / hgt are complete
            GeoLayers G{pl->layers.lat + o, pl->layers.lon + o, pl->layers.hgt + o, nullptr, nullptr, nullptr, nullptr};
            if (share_xyz) {
                G.lat = xs.p[0];
                G.lon = xs.p[1];
                G.hgt = xs.p[2];
            }
            void **lo[4] = {&G.azt, &G.rgm, &G.azoff, &G.rgoff};
            for (int i = 0; i < 4; i++)
                if (J.gp->d_out[i]) *lo[i] = (char *)J.gp->d_out[i] + o * J.esz;
            if (geo_launch(*J.R, pl->line0 + c0, n, G, J.p.out_f32, J.gp->d_stats, s) != 0) return -1;
            J.gp->launches++;
        }
        return cudaEventRecord(done, s) == cudaSuccess ? 0 : -1;
    };
    HostSink sink(st.d, pl->p.device); // pageable destinations (numpy.memmap over the output files) are bounced in parallel
    auto copy_chunk = [&](int c0, int n, cudaEvent_t done) -> cudaError_t {
        const size_t o = (size_t)c0 * w, cnt = (size_t)n * w;
        cudaError_t e = cudaStreamWaitEvent(st.d, done, 0);
        if (e == cudaSuccess) e = sink.copy(out->lat + o, pl->layers.lat + o, sizeof(double) * cnt);
        if (e == cudaSuccess) e = sink.copy(out->lon + o, pl->layers.lon + o, sizeof(double) * cnt);
        if (e == cudaSuccess) e = sink.copy(out->hgt + o, pl->layers.hgt + o, sizeof(double) * cnt);
        if (e == cudaSuccess && out->los && pl->layers.los) e = sink.copy(out->los + 2 * o, pl->layers.los + 2 * o, sizeof(float) * 2 * cnt);
        if (e == cudaSuccess && out->inc && pl->layers.inc) e = sink.copy(out->inc + 2 * o, pl->layers.inc + 2 * o, sizeof(float) * 2 * cnt);
        if (e == cudaSuccess && out->mask && pl->layers.mask) e = sink.copy(out->mask + o, pl->layers.mask + o, cnt);
        for (Job &J : fused)
            for (int i = 0; i < 4; i++)
                if (e == cudaSuccess && J.hout[i] && J.gp->d_out[i])
                    e = sink.copy((char *)J.hout[i] + o * J.esz, (char *)J.gp->d_out[i] + o * J.esz, cnt * J.esz);
        return e;
    };
    const int nchunks = (pl->nlines + cl - 1) / cl;
    st.ev.resize(nchunks, nullptr);
    for (int c = 0; c < nchunks; c++) CK(cudaEventCreateWithFlags(&st.ev[c], cudaEventDisableTiming));
    // kernels of chunk c+1 are queued before the (possibly host-blocking, pageable) copies of chunk c
    if (launch_chunk(0, cl < pl->nlines ? cl : pl->nlines, st.ev[0]) != 0)
        return fail(err, errlen, B200_EINVAL, "cannot launch the topo kernels");
    for (int c = 0; c < nchunks; c++) {
        const int c0 = c * cl, n = (c0 + cl <= pl->nlines) ? cl : pl->nlines - c0;
        if (c + 1 < nchunks) {
            const int d0 = (c + 1) * cl, dn = (d0 + cl <= pl->nlines) ? cl : pl->nlines - d0;
            if (launch_chunk(d0, dn, st.ev[c + 1]) != 0) return fail(err, errlen, B200_EINVAL, "cannot launch the topo kernels");
        }
        CK(copy_chunk(c0, n, st.ev[c]));
    }
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(st.d));
    sink.finish();
    CK(cudaEventElapsedTime(&pl->ms_kernels, pl->ev0, pl->ev1));
    pl->ms_pixels = pl->ms_solve = pl->ms_mask = 0.f; // not separable in the pipelined form
    pl->launches += launches;
    pl->executed = true;
    rc = b200_topo_plan_fetch(pl, nullptr, res, err, errlen);
    if (rc == B200_OK && res) res->ms_total = (float)(now_ms() - t0);
    for (int j = 0; j < njobs && rc == B200_OK; j++) {
        fused[j].gp->executed = true;
        fused[j].gp->ms_kernels = 0.f; // not separable from the topo kernels of the same stream (see the topo result)
        rc = b200_geo_plan_fetch(fused[j].gp.get(), nullptr, jobs[j].res, err, errlen);
        if (rc == B200_OK && jobs[j].res) jobs[j].res->ms_total = (float)(now_ms() - t0);
    }
    return rc;
}

extern "C" int b200_topo_run(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                             const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                             const b200_topo_outputs *out, b200_topo_result *res, char *err, size_t errlen)
{
    return topo_run_impl(p, dem, dem_dtype, orbit, dop, slrng, rho_image, out, res, 0, nullptr, err, errlen);
}

extern "C" int b200_topo_geo2rdr_run(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                                     const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                                     const b200_topo_outputs *out, b200_topo_result *res, int njobs, const b200_geo_job *jobs,
                                     char *err, size_t errlen)
{
    if (njobs < 0 || (njobs > 0 && !jobs)) return fail(err, errlen, B200_EINVAL, "bad geo2rdr job list");
    return topo_run_impl(p, dem, dem_dtype, orbit, dop, slrng, rho_image, out, res, njobs, jobs, err, errlen);
}

// =================================================================================================
// geozero
// =================================================================================================
namespace {

struct GeozeroGrid {
    int min_lat_idx, max_lat_idx, min_lon_idx, max_lon_idx, geo_len, geo_wid;
    double lat_firstr, lon_firstr, dlatr, dlonr;
};

// geozero.f90:118-119, 146-149, 159-170 (real*8 -> integer assignments truncate)
GeozeroGrid geozero_grid(const b200_geozero_params &p)
{
    GeozeroGrid g;
    const double pi = 4.0 * atan(1.0);
    const double deg2rad = pi / 180.0;
    g.dlonr = p.delta_lon * deg2rad;
    g.dlatr = p.delta_lat * deg2rad;
    g.lon_firstr = p.first_lon * deg2rad;
    g.lat_firstr = p.first_lat * deg2rad;
    const double min_latr = p.min_lat * deg2rad, max_latr = p.max_lat * deg2rad;
    const double min_lonr = p.min_lon * deg2rad, max_lonr = p.max_lon * deg2rad;
    g.min_lat_idx = (int)((min_latr - g.lat_firstr) / g.dlatr + 1);
    g.min_lon_idx = (int)((min_lonr - g.lon_firstr) / g.dlonr);
    g.max_lat_idx = (int)((max_latr - g.lat_firstr) / g.dlatr);
    g.max_lon_idx = (int)((max_lonr - g.lon_firstr) / g.dlonr + 1);
    g.geo_len = g.min_lat_idx - g.max_lat_idx;
    g.geo_wid = g.max_lon_idx - g.min_lon_idx;
    return g;
}

int geozero_check(const b200_geozero_params *p, char *err, size_t errlen)
{
    if (!p) return fail(err, errlen, B200_EINVAL, "params is NULL");
    if (p->length < 1 || p->width < 1) return fail(err, errlen, B200_EINVAL, "bad image size %d x %d", p->length, p->width);
    if (p->dem_width < 1 || p->dem_length < 1) return fail(err, errlen, B200_EINVAL, "bad DEM size");
    if (p->delta_lat == 0.0 || p->delta_lon == 0.0) return fail(err, errlen, B200_EINVAL, "DEM posting is zero");
    if (p->prf <= 0 || p->nazlooks < 1 || p->nrnglooks < 1) return fail(err, errlen, B200_EINVAL, "bad prf / looks");
    if (p->look_side != -1 && p->look_side != 1) return fail(err, errlen, B200_EINVAL, "look side must be -1 or +1");
    return B200_OK;
}

} // namespace

struct b200_geozero_plan {
    b200_geozero_params p{};
    GeozeroGrid g{};
    GeozeroConst C{};
    GeozeroGeometry G{nullptr, nullptr, nullptr, nullptr, nullptr};
    float *d_dem = nullptr;
    void *d_raw = nullptr;
    int *d_maxkey = nullptr;
    double *d_poly = nullptr;
    OrbitPolyView op{};
    DeviceOrbit dorb;
    GeoMid *d_mid = nullptr;
    GeozeroStats *d_stats = nullptr;
    float *d_sinc = nullptr;
    void *d_img = nullptr, *d_out = nullptr;
    size_t img_bytes = 0, out_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    float ms_setup = 0.f, ms_solve = 0.f, ms_kernels = 0.f;
    long long num_outside_dem = 0, num_outside_image = 0, num_valid = 0, iterations = 0;
    int launches = 0;

    ~b200_geozero_plan()
    {
        cudaSetDevice(p.device);
        dfree(G.az_idx); dfree(G.rng_idx); dfree(G.dem_crop); dfree(G.row_sc); dfree(G.col_sc);
        dfree(d_dem); dfree(d_raw); dfree(d_maxkey); dfree(d_poly); dfree(dorb.buf); dfree(d_mid); dfree(d_stats);
        dfree(d_sinc); dfree(d_img); dfree(d_out);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_mid) cudaEventDestroy(ev_mid);
        if (stream) cudaStreamDestroy(stream);
    }
};

extern "C" int b200_geozero_grid(const b200_geozero_params *p, int *geo_width, int *geo_length, char *err, size_t errlen)
{
    int rc = geozero_check(p, err, errlen);
    if (rc != B200_OK) return rc;
    const GeozeroGrid g = geozero_grid(*p);
    if (geo_width) *geo_width = g.geo_wid;
    if (geo_length) *geo_length = g.geo_len;
    return B200_OK;
}

static int geozero_plan_build(b200_geozero_plan *pl, const void *dem, int dem_dtype, const b200_orbit *orbit,
                              const b200_poly1d *dop, char *err, size_t errlen)
{
    const b200_geozero_params &p = pl->p;
    int rc;
    // geozero always interpolates the orbit with the Hermite scheme (interpolateWGS84Orbit_f, geozero.f90:229, :341)
    if ((rc = check_orbit(orbit, B200_ORBIT_HERMITE, err, errlen)) != B200_OK) return rc;
    if (!dop || !dop->coeffs || dop->order < 0 || dop->order + 1 > kMaxPoly1dCoeffs)
        return fail(err, errlen, B200_EINVAL, "bad doppler polynomial");
    if (!dem) return fail(err, errlen, B200_EINVAL, "dem is NULL");
    if (dem_dtype != B200_DEM_F32 && dem_dtype != B200_DEM_I16) return fail(err, errlen, B200_EINVAL, "bad dem_dtype");
    pl->g = geozero_grid(p);
    const GeozeroGrid &g = pl->g;
    if (g.geo_len < 1 || g.geo_wid < 1)
        return fail(err, errlen, B200_EINVAL, "empty output grid (%d lines x %d samples): check the bounding box", g.geo_len,
                    g.geo_wid);
    if ((rc = select_device(p.device, err, errlen)) != B200_OK) return rc;
    CK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&pl->ev0));
    CK(cudaEventCreate(&pl->ev1));
    CK(cudaEventCreate(&pl->ev_mid));
    cudaStream_t s = pl->stream;
    CK(cudaEventRecord(pl->ev0, s));

    GeozeroConst &C = pl->C;
    C.elp = make_ellipsoid(p.major, p.e2);
    C.wvl = p.wvl;
    C.tstart = p.t0; // :126-139
    C.dtaz = p.nazlooks / p.prf;
    const double tend = p.t0 + (p.length - 1) * C.dtaz;
    C.tmid = 0.5 * (C.tstart + tend);
    C.rngstart = p.rho0;
    C.dmrg = p.nrnglooks * p.drho;
    C.length = p.length;
    C.width = p.width;
    C.look_side = p.look_side;
    C.lat_firstr = g.lat_firstr;
    C.lon_firstr = g.lon_firstr;
    C.dlatr = g.dlatr;
    C.dlonr = g.dlonr;
    C.max_lat_idx = g.max_lat_idx;
    C.min_lon_idx = g.min_lon_idx;
    C.geo_len = g.geo_len;
    C.geo_wid = g.geo_wid;
    C.demwidth = p.dem_width;
    C.demlength = p.dem_length;
    // doppler-vs-range polynomial and its derivative (:196-224)
    C.fd.order = dop->order;
    C.fd.mean = p.rho0 + dop->mean * p.drho;
    C.fd.norm = dop->norm * p.drho;
    for (int k = 0; k <= dop->order; k++) C.fd.c[k] = dop->coeffs[k] * p.prf;
    if (C.fd.order == 0) {
        C.fdd.order = 0;
        C.fdd.mean = 0.0;
        C.fdd.norm = 1.0;
        C.fdd.c[0] = 0.0;
    } else {
        C.fdd.order = C.fd.order - 1;
        C.fdd.mean = C.fd.mean;
        C.fdd.norm = C.fd.norm;
        for (int k = 1; k <= dop->order; k++) C.fdd.c[k - 1] = k * C.fd.c[k] / C.fd.norm;
    }

    // ---- the part of the DEM under the output grid ----
    int r0 = g.max_lat_idx < 0 ? 0 : g.max_lat_idx, r1 = g.max_lat_idx + g.geo_len;
    if (r1 > p.dem_length) r1 = p.dem_length;
    int c0 = g.min_lon_idx < 0 ? 0 : g.min_lon_idx, c1 = g.min_lon_idx + g.geo_wid;
    if (c1 > p.dem_width) c1 = p.dem_width;
    C.dem_row0 = r0;
    C.dem_col0 = c0;
    C.dem_rows = r1 > r0 ? r1 - r0 : 0;
    C.dem_cols = c1 > c0 ? c1 - c0 : 0;
    // lines of the output grid that fall outside the DEM (:250-253 counts demwidth pixels per such line)
    pl->num_outside_dem = 0;
    for (int line = 0; line < g.geo_len; line++) {
        const int idxlat = g.max_lat_idx + line;
        if (idxlat < 0 || idxlat > p.dem_length - 1) pl->num_outside_dem += p.dem_width;
    }
    const size_t ncell = (size_t)C.dem_rows * (size_t)C.dem_cols;
    if (ncell) {
        const size_t esz = dem_dtype == B200_DEM_I16 ? 2 : 4;
        CK(dmalloc(&pl->d_dem, sizeof(float) * ncell));
        CK(dmalloc(&pl->d_maxkey, sizeof(int)));
        const char *src = (const char *)dem + ((size_t)r0 * (size_t)p.dem_width + (size_t)c0) * esz;
        void *dst = pl->d_dem;
        if (dem_dtype == B200_DEM_I16) {
            CK(dmalloc(&pl->d_raw, esz * ncell));
            dst = pl->d_raw;
        }
        CK(cudaMemcpy2DAsync(dst, (size_t)C.dem_cols * esz, src, (size_t)p.dem_width * esz, (size_t)C.dem_cols * esz,
                             (size_t)C.dem_rows, cudaMemcpyHostToDevice, s));
        int init = (int)0x80000000;
        CK(cudaMemcpyAsync(pl->d_maxkey, &init, sizeof init, cudaMemcpyHostToDevice, s));
        launch_dem_prepare(dst, dem_dtype, pl->d_dem, ncell, pl->d_maxkey, s); // 'read' FLOAT caster (Geozero.py:204)
        pl->launches++;
    } else {
        CK(dmalloc(&pl->d_dem, sizeof(float)));
    }
    C.dem = pl->d_dem;

    // ---- orbit: mid-scene state (:229-236) and the per-window Hermite polynomials ----
    if ((rc = upload_orbit(orbit, pl->dorb, s, err, errlen)) != B200_OK) return rc;
    CK(dmalloc(&pl->d_mid, sizeof(GeoMid)));
    launch_geo_setup(B200_ORBIT_HERMITE, pl->dorb.view, C.tmid, pl->d_mid, s);
    pl->launches++;
    GeoMid mid;
    CK(cudaMemcpyAsync(&mid, pl->d_mid, sizeof mid, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (mid.stat_mid != 0) return fail(err, errlen, B200_EORBIT, "Cannot interpolate orbits at the center of scene.");
    C.xyz_mid = Vec3{mid.xyz[0], mid.xyz[1], mid.xyz[2]};
    C.vel_mid = Vec3{mid.vel[0], mid.vel[1], mid.vel[2]};
    {
        HostOrbitPoly hp;
        if (!build_orbit_poly(B200_ORBIT_HERMITE, orbit->nvec, orbit->t, orbit->pos, orbit->vel, hp))
            return fail(err, errlen, B200_EORBIT, "cannot build the orbit polynomials");
        const size_t nt = (size_t)hp.n, nw = (size_t)hp.nwin, nc = hp.cp.size();
        std::vector<double> blob(nt + 2 * nw + nc);
        memcpy(blob.data(), orbit->t, sizeof(double) * nt);
        memcpy(blob.data() + nt, hp.tc.data(), sizeof(double) * nw);
        memcpy(blob.data() + nt + nw, hp.inv_h.data(), sizeof(double) * nw);
        memcpy(blob.data() + nt + 2 * nw, hp.cp.data(), sizeof(double) * nc);
        CK(dmalloc(&pl->d_poly, sizeof(double) * blob.size()));
        CK(cudaMemcpyAsync(pl->d_poly, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        pl->op = OrbitPolyView{hp.method, hp.n, hp.nwin, hp.ncoef, pl->d_poly, pl->d_poly + nt, pl->d_poly + nt + nw,
                               pl->d_poly + nt + 2 * nw, nullptr};
    }

    // ---- geometry of the grid: solved once ----
    const size_t npix = (size_t)g.geo_len * (size_t)g.geo_wid;
    CK(dmalloc(&pl->G.az_idx, sizeof(double) * npix));
    CK(dmalloc(&pl->G.rng_idx, sizeof(double) * npix));
    CK(dmalloc(&pl->G.dem_crop, sizeof(short) * npix));
    CK(dmalloc(&pl->G.row_sc, sizeof(double) * 2 * (size_t)g.geo_len));
    CK(dmalloc(&pl->G.col_sc, sizeof(double) * 2 * (size_t)g.geo_wid));
    CK(dmalloc(&pl->d_stats, sizeof(GeozeroStats)));
    CK(cudaMemsetAsync(pl->d_stats, 0, sizeof(GeozeroStats), s));
    CK(cudaEventRecord(pl->ev_mid, s));
    launch_geozero_axes(C, pl->G, s);
    if (launch_geozero_solve(C, pl->op, pl->G, pl->d_stats, s) != 0)
        return fail(err, errlen, B200_EINVAL, "cannot launch the geozero solve kernel");
    pl->launches += 2;
    CK(cudaEventRecord(pl->ev1, s));
    GeozeroStats st;
    CK(cudaMemcpyAsync(&st, pl->d_stats, sizeof st, cudaMemcpyDeviceToHost, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&pl->ms_setup, pl->ev0, pl->ev1));
    CK(cudaEventElapsedTime(&pl->ms_solve, pl->ev_mid, pl->ev1));
    pl->iterations = (long long)st.iterations;
    if (pl->d_raw) {
        dfree(pl->d_raw);
        pl->d_raw = nullptr;
    }
    return B200_OK;
}

extern "C" int b200_geozero_plan_create(const b200_geozero_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                                        const b200_poly1d *dop, b200_geozero_plan **plan, char *err, size_t errlen)
{
    if (!plan) return fail(err, errlen, B200_EINVAL, "plan is NULL");
    *plan = nullptr;
    int rc = geozero_check(p, err, errlen);
    if (rc != B200_OK) return rc;
    b200_geozero_plan *pl = new (std::nothrow) b200_geozero_plan;
    if (!pl) return fail(err, errlen, B200_ENOMEM, "out of host memory");
    pl->p = *p;
    rc = geozero_plan_build(pl, dem, dem_dtype, orbit, dop, err, errlen);
    if (rc != B200_OK) {
        delete pl;
        return rc;
    }
    *plan = pl;
    return B200_OK;
}

extern "C" int b200_geozero_plan_geocode(b200_geozero_plan *pl, const void *image, int is_complex, int nbands, int scheme,
                                         int method, void *out, float *ms_kernels, char *err, size_t errlen)
{
    if (!pl || !image || !out) return fail(err, errlen, B200_EINVAL, "plan/image/out is NULL");
    if (nbands < 1) return fail(err, errlen, B200_EINVAL, "nbands must be >= 1");
    if (scheme != B200_SCHEME_BIL && scheme != B200_SCHEME_BIP && scheme != B200_SCHEME_BSQ)
        return fail(err, errlen, B200_EINVAL, "unknown interleaving scheme %d", scheme);
    if (method < B200_GEOZERO_SINC || method > B200_GEOZERO_NEAREST)
        return fail(err, errlen, B200_EINVAL, "Undefined interpolation method."); // geozero.f90:110-112
    const b200_geozero_params &p = pl->p;
    CK(cudaSetDevice(p.device));
    cudaStream_t s = pl->stream;
    const size_t esz = is_complex ? 8 : 4;
    const size_t W = (size_t)p.width, L = (size_t)p.length, gw = (size_t)pl->g.geo_wid, gl = (size_t)pl->g.geo_len, nb = (size_t)nbands;
    const size_t in_bytes = esz * W * L * nb, out_bytes = esz * gw * gl * nb;
    if (pl->img_bytes < in_bytes) {
        dfree(pl->d_img);
        pl->d_img = nullptr;
        CK(dmalloc(&pl->d_img, in_bytes));
        pl->img_bytes = in_bytes;
    }
    if (pl->out_bytes < out_bytes) {
        dfree(pl->d_out);
        pl->d_out = nullptr;
        CK(dmalloc(&pl->d_out, out_bytes));
        pl->out_bytes = out_bytes;
    }
    if (method == B200_GEOZERO_SINC && !pl->d_sinc) { // prepareMethods (geozeroMethods.F:46-64)
        std::vector<float> tab((size_t)kSincSub * kSincLen);
        sinc_make_table(tab.data());
        CK(dmalloc(&pl->d_sinc, sizeof(float) * tab.size()));
        CK(cudaMemcpyAsync(pl->d_sinc, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    HostSource source(s, pl->p.device); // the image being geocoded is usually a mapping of its file
    CK(source.copy(pl->d_img, image, in_bytes));
    CK(cudaEventRecord(pl->ev0, s));
    for (int b = 0; b < nbands; b++) {
        BandView iv, ov;
        if (scheme == B200_SCHEME_BIL) {
            iv = BandView{(size_t)b * W, nb * W, 1};
            ov = BandView{(size_t)b * gw, nb * gw, 1};
        } else if (scheme == B200_SCHEME_BIP) {
            iv = BandView{(size_t)b, nb * W, nb};
            ov = BandView{(size_t)b, nb * gw, nb};
        } else {
            iv = BandView{(size_t)b * W * L, W, 1};
            ov = BandView{(size_t)b * gw * gl, gw, 1};
        }
        // the counters the reference prints are those of one band; keep the last band's
        CK(cudaMemsetAsync(pl->d_stats, 0, sizeof(GeozeroStats), s));
        if (launch_geozero_interp(pl->C, pl->G, method, is_complex, pl->d_img, iv, pl->d_out, ov, pl->d_sinc, pl->d_stats, s) != 0)
            return fail(err, errlen, B200_EINVAL, "cannot launch the geozero gather kernel");
        pl->launches++;
    }
    CK(cudaEventRecord(pl->ev1, s));
    GeozeroStats st;
    CK(cudaMemcpyAsync(&st, pl->d_stats, sizeof st, cudaMemcpyDeviceToHost, s));
    HostSink sink(s, pl->p.device); // ... and so is the geocoded product
    CK(sink.copy(out, pl->d_out, out_bytes));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    sink.finish();
    CK(cudaEventElapsedTime(&pl->ms_kernels, pl->ev0, pl->ev1));
    pl->num_outside_image = (long long)st.outside_image;
    pl->num_valid = (long long)st.valid;
    if (ms_kernels) *ms_kernels = pl->ms_kernels;
    return B200_OK;
}

extern "C" int b200_geozero_plan_fetch(b200_geozero_plan *pl, int16_t *dem_crop, double *az_idx, double *rng_idx,
                                       b200_geozero_result *res, char *err, size_t errlen)
{
    if (!pl) return fail(err, errlen, B200_EINVAL, "plan is NULL");
    CK(cudaSetDevice(pl->p.device));
    cudaStream_t s = pl->stream;
    const size_t npix = (size_t)pl->g.geo_len * (size_t)pl->g.geo_wid;
    if (dem_crop) CK(cudaMemcpyAsync(dem_crop, pl->G.dem_crop, sizeof(short) * npix, cudaMemcpyDeviceToHost, s));
    if (az_idx) CK(cudaMemcpyAsync(az_idx, pl->G.az_idx, sizeof(double) * npix, cudaMemcpyDeviceToHost, s));
    if (rng_idx) CK(cudaMemcpyAsync(rng_idx, pl->G.rng_idx, sizeof(double) * npix, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (res) {
        const b200_geozero_params &p = pl->p;
        res->geo_width = pl->g.geo_wid;
        res->geo_length = pl->g.geo_len;
        res->geo_min_lat = (p.first_lat + pl->g.min_lat_idx * p.delta_lat); // geozero.f90:419-422
        res->geo_max_lat = (p.first_lat + pl->g.max_lat_idx * p.delta_lat);
        res->geo_min_lon = (p.first_lon + pl->g.min_lon_idx * p.delta_lon);
        res->geo_max_lon = (p.first_lon + pl->g.max_lon_idx * p.delta_lon);
        res->num_outside_dem = pl->num_outside_dem;
        res->num_outside_image = pl->num_outside_image;
        res->num_valid = pl->num_valid;
        res->iterations = pl->iterations;
        res->ms_setup = pl->ms_setup;
        res->ms_solve = pl->ms_solve;
        res->ms_kernels = pl->ms_kernels;
        res->ms_total = 0.f;
        res->gpu_launches = pl->launches;
    }
    return B200_OK;
}

extern "C" void b200_geozero_plan_destroy(b200_geozero_plan *pl) { delete pl; }

extern "C" int b200_geozero_run(const b200_geozero_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                                const b200_poly1d *dop, const void *image, int is_complex, int nbands, int scheme, int method,
                                void *out, int16_t *dem_crop, b200_geozero_result *res, char *err, size_t errlen)
{
    const double t0 = now_ms();
    b200_geozero_plan *pl = nullptr;
    int rc = b200_geozero_plan_create(p, dem, dem_dtype, orbit, dop, &pl, err, errlen);
    if (rc != B200_OK) return rc;
    rc = b200_geozero_plan_geocode(pl, image, is_complex, nbands, scheme, method, out, nullptr, err, errlen);
    if (rc == B200_OK) rc = b200_geozero_plan_fetch(pl, dem_crop, nullptr, nullptr, res, err, errlen);
    delete pl;
    if (rc == B200_OK && res) res->ms_total = (float)(now_ms() - t0);
    return rc;
}

// =================================================================================================
// resamp_slc
// =================================================================================================
namespace {

// normalised sinc table of resamp_slcMethods.f:57-83 (every sub-sample phase scaled to unit sum in real*8, then real*4)
void resamp_sinc_table(float *fintp)
{
    const double pi = 4.0 * atan(1.0);
    const int n = kSincSub * kSincLen;
    std::vector<double> r_filter((size_t)n + 1, 0.0);
    const double r_soff = n / 2.0;
    for (int i = 0; i < n; i++) { // sinc_coef(beta = 1, relfiltlen = 8, decfactor = 8192, pedestal = 0, weight = 1)
        const double r_wa = i - r_soff;
        const double r_s = r_wa * 1.0 / (1.0 * kSincSub);
        const double r_fct = (r_s != 0.0) ? sin(pi * r_s) / (pi * r_s) : 1.0;
        const double r_wgt = (1.0 - 0.5) + 0.5 * cos((pi * r_wa) / r_soff);
        r_filter[i] = r_fct * r_wgt;
    }
    for (int i = 0; i < kSincSub; i++) {
        double ssum = 0.0;
        for (int j = 0; j < kSincLen; j++) ssum = ssum + r_filter[i + j * kSincSub];
        for (int j = 0; j < kSincLen; j++) r_filter[i + j * kSincSub] = r_filter[i + j * kSincSub] / ssum;
    }
    for (int i = 0; i < kSincLen; i++)
        for (int j = 0; j < kSincSub; j++) fintp[i + j * kSincLen] = (float)r_filter[j + i * kSincSub];
}

int fill_poly2d_or_zero(const b200_poly2d *src, Poly2dDev &dst, const char *what, char *err, size_t errlen)
{
    if (src) return fill_poly2d(src, dst, what, err, errlen);
    memset(&dst, 0, sizeof dst); // order 0, coefficient 0 (Resamp_slc.py:86-140)
    dst.norm_range = dst.norm_azimuth = dst.inv_norm_range = dst.inv_norm_azimuth = 1.0;
    return B200_OK;
}

bool poly2d_is_zero(const Poly2dDev &p)
{
    const int n = (p.range_order + 1) * (p.azimuth_order + 1);
    for (int i = 0; i < n; i++)
        if (p.c[i] != 0.0) return false;
    return true;
}

struct ResampBuffers {
    float2 *d_in = nullptr, *d_out = nullptr;
    void *d_raz = nullptr, *d_rrg = nullptr;
    float *d_sinc = nullptr;
    ResampStats *d_stats = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~ResampBuffers()
    {
        dfree(d_in); dfree(d_out); dfree(d_raz); dfree(d_rrg); dfree(d_sinc); dfree(d_stats);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

} // namespace

// resid_on_device: resid_az / resid_rg are device pointers on p->device (the offsets a geo2rdr plan left in HBM)
static int resamp_core(const b200_resamp_params *p, const b200_poly2d *rg_carrier, const b200_poly2d *az_carrier,
                       const b200_poly2d *rg_offsets, const b200_poly2d *az_offsets, const b200_poly2d *doppler,
                       const float *slc_in, const void *resid_az, const void *resid_rg, int resid_dtype, bool resid_on_device,
                       float *slc_out, b200_resamp_result *res, char *err, size_t errlen)
{
    const double t0 = now_ms();
    if (!p || !slc_in || !slc_out) return fail(err, errlen, B200_EINVAL, "params / input / output image is NULL");
    if (p->in_width < 1 || p->in_length < 1 || p->out_width < 1 || p->out_length < 1)
        return fail(err, errlen, B200_EINVAL, "bad image sizes");
    if (resid_dtype != B200_RESID_F64 && resid_dtype != B200_RESID_F32) return fail(err, errlen, B200_EINVAL, "bad resid_dtype");
    ResampConst C{};
    C.inwidth = p->in_width; C.inlength = p->in_length; C.outwidth = p->out_width; C.outlength = p->out_length;
    C.wvl = p->wvl; C.slr = p->slr; C.r0 = p->r0; C.refwvl = p->ref_wvl; C.refr0 = p->ref_r0; C.refslr = p->ref_slr;
    C.flatten = p->flatten;
    C.pi = 4.0 * atan(1.0);
    int rc;
    if ((rc = fill_poly2d_or_zero(rg_carrier, C.rg_carrier, "range carrier", err, errlen)) != B200_OK) return rc;
    if ((rc = fill_poly2d_or_zero(az_carrier, C.az_carrier, "azimuth carrier", err, errlen)) != B200_OK) return rc;
    if ((rc = fill_poly2d_or_zero(rg_offsets, C.rg_off, "range offsets", err, errlen)) != B200_OK) return rc;
    if ((rc = fill_poly2d_or_zero(az_offsets, C.az_off, "azimuth offsets", err, errlen)) != B200_OK) return rc;
    if ((rc = fill_poly2d_or_zero(doppler, C.dop, "doppler", err, errlen)) != B200_OK) return rc;
    C.has_carrier = (poly2d_is_zero(C.rg_carrier) && poly2d_is_zero(C.az_carrier)) ? 0 : 1;
    if (p->flatten && (p->wvl == 0.0 || p->ref_wvl == 0.0)) return fail(err, errlen, B200_EINVAL, "flattening needs the wavelengths");
    if ((rc = select_device(p->device, err, errlen)) != B200_OK) return rc;

    ResampBuffers B;
    CK(cudaStreamCreateWithFlags(&B.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&B.ev0));
    CK(cudaEventCreate(&B.ev1));
    cudaStream_t s = B.stream;
    const size_t nin = (size_t)p->in_width * (size_t)p->in_length, nout = (size_t)p->out_width * (size_t)p->out_length;
    const size_t rsz = resid_dtype == B200_RESID_F32 ? 4 : 8;
    CK(dmalloc(&B.d_in, sizeof(float2) * nin));
    CK(dmalloc(&B.d_out, sizeof(float2) * nout));
    CK(dmalloc(&B.d_sinc, sizeof(float) * kSincSub * kSincLen));
    CK(dmalloc(&B.d_stats, sizeof(ResampStats)));
    std::vector<float> tab((size_t)kSincSub * kSincLen);
    resamp_sinc_table(tab.data());
    CK(cudaMemcpyAsync(B.d_sinc, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, s));
    HostSource source(s, p->device); // the SLC and the .off rasters arrive as mappings of their files (Resamp_slc.py:86-160)
    CK(source.copy(B.d_in, slc_in, sizeof(float2) * nin));
    const void *d_raz = resid_az, *d_rrg = resid_rg;
    if (!resid_on_device) {
        if (resid_az) {
            CK(dmalloc(&B.d_raz, rsz * nout));
            CK(source.copy(B.d_raz, resid_az, rsz * nout));
            d_raz = B.d_raz;
        }
        if (resid_rg) {
            CK(dmalloc(&B.d_rrg, rsz * nout));
            CK(source.copy(B.d_rrg, resid_rg, rsz * nout));
            d_rrg = B.d_rrg;
        }
    }
    CK(cudaMemsetAsync(B.d_stats, 0, sizeof(ResampStats), s));
    CK(cudaEventRecord(B.ev0, s));
    int launches = 0;
    // with both carriers identically zero the up-front pass multiplies every sample by (1, -0): the identity
    if (C.has_carrier) {
        launch_resamp_carrier(C, B.d_in, B.d_in, s);
        launches++;
    }
    if (launch_resamp_slc(C, B.d_in, d_raz, d_rrg, resid_dtype == B200_RESID_F32, B.d_sinc, B.d_out, B.d_stats, s) != 0)
        return fail(err, errlen, B200_EINVAL, "cannot launch the resampling kernel");
    launches++;
    CK(cudaEventRecord(B.ev1, s));
    ResampStats st;
    CK(cudaMemcpyAsync(&st, B.d_stats, sizeof st, cudaMemcpyDeviceToHost, s));
    HostSink sink(s, p->device);
    CK(sink.copy(slc_out, B.d_out, sizeof(float2) * nout));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    sink.finish();
    if (res) {
        res->num_valid = (long long)st.valid;
        CK(cudaEventElapsedTime(&res->ms_kernels, B.ev0, B.ev1));
        res->gpu_launches = launches;
        res->ms_total = (float)(now_ms() - t0);
    }
    return B200_OK;
}

extern "C" int b200_resamp_slc_run(const b200_resamp_params *p, const b200_poly2d *rg_carrier, const b200_poly2d *az_carrier,
                                   const b200_poly2d *rg_offsets, const b200_poly2d *az_offsets, const b200_poly2d *doppler,
                                   const float *slc_in, const void *resid_az, const void *resid_rg, int resid_dtype,
                                   float *slc_out, b200_resamp_result *res, char *err, size_t errlen)
{
    return resamp_core(p, rg_carrier, az_carrier, rg_offsets, az_offsets, doppler, slc_in, resid_az, resid_rg, resid_dtype, false,
                       slc_out, res, err, errlen);
}

extern "C" int b200_resamp_slc_from_geo_plan(const b200_resamp_params *p, b200_geo_plan *geo, const b200_poly2d *rg_carrier,
                                             const b200_poly2d *az_carrier, const b200_poly2d *rg_offsets,
                                             const b200_poly2d *az_offsets, const b200_poly2d *doppler, const float *slc_in,
                                             float *slc_out, b200_resamp_result *res, char *err, size_t errlen)
{
    if (!p || !geo) return fail(err, errlen, B200_EINVAL, "params / geo2rdr plan is NULL");
    if (!geo->executed || !geo->d_out[2] || !geo->d_out[3])
        return fail(err, errlen, B200_EINVAL, "the geo2rdr plan must have been executed with azimuth and range offsets requested");
    if (geo->p.device != p->device) return fail(err, errlen, B200_EINVAL, "the geo2rdr plan lives on another device");
    if (p->out_width != geo->p.dem_width || p->out_length != geo->nlines)
        return fail(err, errlen, B200_EINVAL, "output grid %d x %d differs from the plan's offset rasters %d x %d", p->out_length,
                    p->out_width, geo->nlines, geo->p.dem_width);
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(geo->stream));
    return resamp_core(p, rg_carrier, az_carrier, rg_offsets, az_offsets, doppler, slc_in, geo->d_out[2], geo->d_out[3],
                       geo->out_f32 ? B200_RESID_F32 : B200_RESID_F64, true, slc_out, res, err, errlen);
}

// =================================================================================================
// multilooking of the geometry layers, mask projection (the other consumers)
// =================================================================================================
namespace {
struct PostStreams { // H2D | kernels | D2H
    cudaStream_t h = nullptr, k = nullptr, d = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<void *> bufs;
    int init()
    {
        if (cudaStreamCreateWithFlags(&h, cudaStreamNonBlocking) != cudaSuccess) return -1;
        if (cudaStreamCreateWithFlags(&k, cudaStreamNonBlocking) != cudaSuccess) return -1;
        if (cudaStreamCreateWithFlags(&d, cudaStreamNonBlocking) != cudaSuccess) return -1;
        if (cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess) return -1;
        return 0;
    }
    cudaEvent_t event()
    {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ev.push_back(e);
        return e;
    }
    ~PostStreams()
    {
        for (cudaStream_t s : {h, k, d})
            if (s) cudaStreamSynchronize(s);
        for (void *b : bufs) dfree(b);
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (cudaStream_t s : {h, k, d})
            if (s) cudaStreamDestroy(s);
    }
};
} // namespace

extern "C" int b200_looks_run(const void *in, void *out, int dtype, int length, int width, int bands, int scheme, int down_looks,
                              int across_looks, int method, int device, b200_looks_result *res, char *err, size_t errlen)
{
    const double t0 = now_ms();
    const size_t esz = type_size(dtype);
    if (!in || !out) return fail(err, errlen, B200_EINVAL, "input / output image is NULL");
    if (esz == 0) return fail(err, errlen, B200_EINVAL, "Error. Unrecognized data type %d", dtype); // looksmodule.cpp:124-127
    if (length < 1 || width < 1 || bands < 1) return fail(err, errlen, B200_EINVAL, "bad image %d x %d x %d", length, width, bands);
    if (scheme != B200_SCHEME_BIL && scheme != B200_SCHEME_BIP && scheme != B200_SCHEME_BSQ)
        return fail(err, errlen, B200_EINVAL, "bad interleaving scheme %d", scheme);
    if (down_looks < 1 || across_looks < 1) return fail(err, errlen, B200_EINVAL, "looks must be >= 1 (%d down, %d across)", down_looks, across_looks);
    if (method != B200_LOOKS_AVERAGE && method != B200_LOOKS_NEAREST) return fail(err, errlen, B200_EINVAL, "bad method %d", method);
    LooksGeom G{length, width, bands, scheme, down_looks, across_looks, length / down_looks, width / across_looks, 0, 0};
    if (res) {
        res->out_length = G.out_length;
        res->out_width = G.out_width;
        res->ms_kernels = res->ms_total = 0.f;
        res->gpu_launches = 0;
    }
    if (method == B200_LOOKS_AVERAGE) {
        const long long per_out = (long long)across_looks * (scheme == B200_SCHEME_BIP ? bands : 1) * (dtype == B200_T_CFLOAT ? 2 : 1);
        if (per_out > kLooksMaxTile)
            return fail(err, errlen, B200_EINVAL, "across_looks x interleaved bands = %lld exceeds the %d column sums a tile holds", per_out,
                        kLooksMaxTile);
    }
    int rc = select_device(device, err, errlen);
    if (rc != B200_OK) return rc;
    if (G.out_length == 0 || G.out_width == 0) return B200_OK; // nothing to write (the reference's loops do not execute)
    PostStreams st;
    if (st.init() != 0) return fail(err, errlen, B200_ECUDA, "cannot create streams");
    const size_t in_bytes = (size_t)length * width * bands * esz, out_bytes = (size_t)G.out_length * G.out_width * bands * esz;
    void *d_in = nullptr, *d_out = nullptr;
    CK(dmalloc(&d_in, in_bytes));
    st.bufs.push_back(d_in);
    CK(dmalloc(&d_out, out_bytes));
    st.bufs.push_back(d_out);
    // blocks of output lines: H2D(c+1) | kernel(c) | D2H(c-1); a block of a band-sequential image is one piece per band
    long long cl = 16000000LL / ((long long)width * bands * down_looks);
    if (cl < 1) cl = 1;
    const int pieces = scheme == B200_SCHEME_BSQ ? bands : 1;
    int launches = 0;
    CK(cudaEventRecord(st.ev0, st.k));
    for (int c0 = 0; c0 < G.out_length; c0 += (int)cl) {
        const int n = (c0 + cl <= G.out_length) ? (int)cl : G.out_length - c0;
        for (int b = 0; b < pieces; b++) {
            const size_t per_line = (size_t)width * (pieces > 1 ? 1 : bands) * esz;
            const size_t o = ((size_t)b * length + (size_t)c0 * down_looks) * per_line;
            CK(cudaMemcpyAsync((char *)d_in + o, (const char *)in + o, (size_t)n * down_looks * per_line, cudaMemcpyHostToDevice, st.h));
        }
        cudaEvent_t eh = st.event(), ek = st.event();
        CK(cudaEventRecord(eh, st.h));
        CK(cudaStreamWaitEvent(st.k, eh, 0));
        G.line0 = c0;
        G.nlines = n;
        const int lr = launch_looks(G, dtype, method, d_in, d_out, st.k);
        if (lr != 0) return fail(err, errlen, B200_EINVAL, "cannot launch the looks kernel (%d)", lr);
        launches++;
        CK(cudaEventRecord(ek, st.k));
        CK(cudaStreamWaitEvent(st.d, ek, 0));
        for (int b = 0; b < pieces; b++) {
            const size_t per_line = (size_t)G.out_width * (pieces > 1 ? 1 : bands) * esz;
            const size_t o = ((size_t)b * G.out_length + (size_t)c0) * per_line;
            CK(cudaMemcpyAsync((char *)out + o, (char *)d_out + o, (size_t)n * per_line, cudaMemcpyDeviceToHost, st.d));
        }
    }
    CK(cudaEventRecord(st.ev1, st.k));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st.k));
    CK(cudaStreamSynchronize(st.d));
    if (res) {
        CK(cudaEventElapsedTime(&res->ms_kernels, st.ev0, st.ev1));
        res->gpu_launches = launches;
        res->ms_total = (float)(now_ms() - t0);
    }
    return B200_OK;
}

extern "C" int b200_topo_plan_looks(b200_topo_plan *pl, int layer, int down_looks, int across_looks, int method, void *out,
                                    b200_looks_result *res, char *err, size_t errlen)
{
    const double t0 = now_ms();
    if (!pl || !out) return fail(err, errlen, B200_EINVAL, "plan / out is NULL");
    if (!pl->executed) return fail(err, errlen, B200_EINVAL, "plan was not executed");
    if (down_looks < 1 || across_looks < 1) return fail(err, errlen, B200_EINVAL, "looks must be >= 1 (%d down, %d across)", down_looks, across_looks);
    if (method != B200_LOOKS_AVERAGE && method != B200_LOOKS_NEAREST) return fail(err, errlen, B200_EINVAL, "bad method %d", method);
    const void *src = nullptr;
    int dtype = B200_T_DOUBLE, bands = 1;
    switch (layer) {
    case B200_LAYER_LAT: src = pl->layers.lat; break;
    case B200_LAYER_LON: src = pl->layers.lon; break;
    case B200_LAYER_HGT: src = pl->layers.hgt; break;
    case B200_LAYER_LOS: src = pl->layers.los; dtype = B200_T_FLOAT; bands = 2; break;
    case B200_LAYER_INC: src = pl->layers.inc; dtype = B200_T_FLOAT; bands = 2; break;
    case B200_LAYER_MASK: src = pl->layers.mask; dtype = B200_T_BYTE; break;
    default: return fail(err, errlen, B200_EINVAL, "bad layer %d", layer);
    }
    if (!src) return fail(err, errlen, B200_EINVAL, "layer %d was not requested from this plan", layer);
    if (method == B200_LOOKS_AVERAGE && across_looks > kLooksMaxTile)
        return fail(err, errlen, B200_EINVAL, "across_looks = %d exceeds the %d column sums a tile holds", across_looks, kLooksMaxTile);
    LooksGeom G{pl->nlines, pl->p.width, bands, B200_SCHEME_BIL, down_looks, across_looks, pl->nlines / down_looks,
                pl->p.width / across_looks, 0, pl->nlines / down_looks};
    if (res) {
        res->out_length = G.out_length;
        res->out_width = G.out_width;
        res->ms_kernels = res->ms_total = 0.f;
        res->gpu_launches = 0;
    }
    if (G.out_length == 0 || G.out_width == 0) return B200_OK;
    CK(cudaSetDevice(pl->p.device));
    cudaStream_t s = pl->stream;
    const size_t out_bytes = (size_t)G.out_length * G.out_width * bands * type_size(dtype);
    void *d_out = nullptr;
    CK(dmalloc(&d_out, out_bytes));
    struct Guard {
        void *p;
        ~Guard() { dfree(p); }
    } guard{d_out};
    CK(cudaEventRecord(pl->ev0, s));
    if (launch_looks(G, dtype, method, src, d_out, s) != 0) return fail(err, errlen, B200_EINVAL, "cannot launch the looks kernel");
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    if (res) {
        CK(cudaEventElapsedTime(&res->ms_kernels, pl->ev0, pl->ev1));
        res->gpu_launches = 1;
        res->ms_total = (float)(now_ms() - t0);
    }
    return B200_OK;
}

extern "C" int b200_mask_to_radar_run(const void *mask, int dtype, int mask_length, int mask_width, double start_lat,
                                      double delta_lat, double start_lon, double delta_lon, const void *lat, const void *lon,
                                      int coord_f32, size_t npix, void *out, int device, b200_mask_result *res, char *err,
                                      size_t errlen)
{
    const double t0 = now_ms();
    if (!mask || !lat || !lon || !out) return fail(err, errlen, B200_EINVAL, "mask / lat / lon / out is NULL");
    if (dtype != B200_T_BYTE && dtype != B200_T_SHORT && dtype != B200_T_INT && dtype != B200_T_FLOAT)
        return fail(err, errlen, B200_EINVAL, "mask data type %d is not one of BYTE, SHORT, INT, FLOAT", dtype);
    if (mask_length < 1 || mask_width < 1) return fail(err, errlen, B200_EINVAL, "bad mask grid %d x %d", mask_length, mask_width);
    if (res) {
        res->ms_kernels = res->ms_total = 0.f;
        res->gpu_launches = 0;
    }
    int rc = select_device(device, err, errlen);
    if (rc != B200_OK) return rc;
    if (npix == 0) return B200_OK;
    PostStreams st;
    if (st.init() != 0) return fail(err, errlen, B200_ECUDA, "cannot create streams");
    const size_t esz = type_size(dtype), csz = coord_f32 ? 4 : 8;
    const size_t mbytes = (size_t)mask_length * mask_width * esz;
    void *d_mask = nullptr, *d_lat = nullptr, *d_lon = nullptr, *d_out = nullptr;
    CK(dmalloc(&d_mask, mbytes));
    st.bufs.push_back(d_mask);
    CK(dmalloc(&d_lat, npix * csz));
    st.bufs.push_back(d_lat);
    CK(dmalloc(&d_lon, npix * csz));
    st.bufs.push_back(d_lon);
    CK(dmalloc(&d_out, npix * esz));
    st.bufs.push_back(d_out);
    CK(cudaMemcpyAsync(d_mask, mask, mbytes, cudaMemcpyHostToDevice, st.h));
    const MaskProj M{mask_length, mask_width, start_lat, delta_lat, start_lon, delta_lon};
    const size_t chunk = 16000000;
    int launches = 0;
    CK(cudaEventRecord(st.ev0, st.k));
    for (size_t p0 = 0; p0 < npix; p0 += chunk) {
        const size_t n = p0 + chunk <= npix ? chunk : npix - p0;
        CK(cudaMemcpyAsync((char *)d_lat + p0 * csz, (const char *)lat + p0 * csz, n * csz, cudaMemcpyHostToDevice, st.h));
        CK(cudaMemcpyAsync((char *)d_lon + p0 * csz, (const char *)lon + p0 * csz, n * csz, cudaMemcpyHostToDevice, st.h));
        cudaEvent_t eh = st.event(), ek = st.event();
        CK(cudaEventRecord(eh, st.h));
        CK(cudaStreamWaitEvent(st.k, eh, 0));
        if (launch_mask_to_radar(M, dtype, d_mask, (char *)d_lat + p0 * csz, (char *)d_lon + p0 * csz, coord_f32, n,
                                 (char *)d_out + p0 * esz, st.k) != 0)
            return fail(err, errlen, B200_EINVAL, "cannot launch the mask projection kernel");
        launches++;
        CK(cudaEventRecord(ek, st.k));
        CK(cudaStreamWaitEvent(st.d, ek, 0));
        CK(cudaMemcpyAsync((char *)out + p0 * esz, (char *)d_out + p0 * esz, n * esz, cudaMemcpyDeviceToHost, st.d));
    }
    CK(cudaEventRecord(st.ev1, st.k));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st.k));
    CK(cudaStreamSynchronize(st.d));
    if (res) {
        CK(cudaEventElapsedTime(&res->ms_kernels, st.ev0, st.ev1));
        res->gpu_launches = launches;
        res->ms_total = (float)(now_ms() - t0);
    }
    return B200_OK;
}

// =================================================================================================
// utilities
// =================================================================================================
extern "C" int b200_abi_version(void) { return B200GEOM_ABI_VERSION; }

extern "C" void b200_release_cached_memory(void)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_release_locked(-1);
}

extern "C" int b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int b200_device_name(int device, char *buf, size_t len)
{
    cudaDeviceProp prop;
    if (!buf || !len) return B200_EINVAL;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        buf[0] = 0;
        return B200_ENODEVICE;
    }
    snprintf(buf, len, "%s", prop.name);
    return B200_OK;
}

extern "C" void *b200_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void b200_free_pinned(void *p)
{
    if (p) cudaFreeHost(p);
}

// Device -> page-locked host copy of `bytes` from a scratch device buffer, in chunks of chunk_bytes on one stream: the
// floor of any end-to-end call that has to deliver that many bytes of results (bench.py runs it on all ranks at once)
extern "C" int b200_host_file_register(const void *base, size_t bytes, int fd, long long file_offset, char *err, size_t errlen)
{
    if (!base || !bytes || fd < 0 || file_offset < 0) return fail(err, errlen, B200_EINVAL, "b200_host_file_register: bad arguments");
    const int mine = dup(fd);
    if (mine < 0) return fail(err, errlen, B200_EINVAL, "b200_host_file_register: cannot duplicate descriptor %d (errno %d)", fd, errno);
    std::unique_lock<std::shared_mutex> lk(g_file_mu);
    for (const FileRange &r : g_files)
        if ((const char *)base < r.base + r.bytes && r.base < (const char *)base + bytes) {
            lk.unlock();
            close(mine);
            return fail(err, errlen, B200_EINVAL, "b200_host_file_register: the range overlaps a registered one");
        }
    g_files.push_back(FileRange{(const char *)base, bytes, mine, file_offset});
    return B200_OK;
}

extern "C" unsigned long long b200_host_file_bytes(void) { return g_file_bytes.load(); }
extern "C" unsigned long long b200_host_file_bytes_read(void) { return g_file_bytes_read.load(); }

extern "C" int b200_host_file_unregister(const void *base)
{
    std::unique_lock<std::shared_mutex> lk(g_file_mu);
    for (size_t i = 0; i < g_files.size(); i++)
        if (g_files[i].base == (const char *)base) {
            close(g_files[i].fd);
            g_files.erase(g_files.begin() + (long)i);
            return B200_OK;
        }
    return B200_EINVAL;
}

extern "C" int b200_d2h_floor(int device, void *host, size_t bytes, size_t chunk_bytes, float *ms, char *err, size_t errlen)
{
    int rc = select_device(device, err, errlen);
    if (rc != B200_OK) return rc;
    if (!host || !bytes || !ms) return fail(err, errlen, B200_EINVAL, "host buffer, size and result pointer are mandatory");
    if (chunk_bytes == 0 || chunk_bytes > bytes) chunk_bytes = bytes;
    if (chunk_bytes > ((size_t)1 << 30)) chunk_bytes = (size_t)1 << 30;
    void *d = nullptr;
    CK(dmalloc(&d, chunk_bytes));
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d, 0, chunk_bytes, s);
    if (ce == cudaSuccess) ce = cudaEventRecord(e0, s);
    for (size_t o = 0; o < bytes && ce == cudaSuccess; o += chunk_bytes)
        ce = cudaMemcpyAsync((char *)host + o, d, bytes - o < chunk_bytes ? bytes - o : chunk_bytes, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaEventRecord(e1, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (s) cudaStreamDestroy(s);
    dfree(d);
    if (ce != cudaSuccess) return fail(err, errlen, B200_ECUDA, "device-to-host copy failed: %s", cudaGetErrorString(ce));
    return B200_OK;
}

// DFMA-saturating microbenchmark: 8 independent FMA chains per thread
__global__ void k_fp64_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int b200_fp64_peak(int device, double *tflops, char *err, size_t errlen)
{
    int rc = select_device(device, err, errlen);
    if (rc != B200_OK) return rc;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *d = nullptr;
    CK(dmalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dfree(d);
    if (tflops) *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return B200_OK;
}

__global__ void k_primitive(int what, Ellipsoid e, OrbitView orb, const double *in, double *out)
{
    if (threadIdx.x != 0) return;
    if (what == 0) {
        Vec3 v = llh_to_xyz(e, in[0], in[1], in[2]);
        out[0] = v.x; out[1] = v.y; out[2] = v.z;
    } else if (what == 1) {
        xyz_to_llh(e, Vec3{in[0], in[1], in[2]}, out[0], out[1], out[2]);
    } else {
        Vec3 p = v3(0, 0, 0), v = v3(0, 0, 0);
        int method = what == 2 ? 0 : (what == 3 ? 2 : 1);
        int stat = orbit_interp(method, orb, in[0], p, v);
        out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = v.x; out[4] = v.y; out[5] = v.z; out[6] = (double)stat;
    }
}

extern "C" int b200_device_primitive(int device, int what, double a, double e2, const b200_orbit *orbit, const double *in,
                                     double *out, char *err, size_t errlen)
{
    int rc = select_device(device, err, errlen);
    if (rc != B200_OK) return rc;
    if (what < 0 || what > 4 || !in || !out) return fail(err, errlen, B200_EINVAL, "bad primitive request");
    DeviceOrbit dorb;
    if (what >= 2) {
        if (!orbit) return fail(err, errlen, B200_EINVAL, "orbit is NULL");
        if ((rc = upload_orbit(orbit, dorb, 0, err, errlen)) != B200_OK) return rc;
    }
    double *d = nullptr;
    CK(dmalloc(&d, sizeof(double) * 16));
    CK(cudaMemcpy(d, in, sizeof(double) * 3, cudaMemcpyHostToDevice));
    k_primitive<<<1, 32>>>(what, make_ellipsoid(a, e2), dorb.view, d, d + 8);
    cudaError_t e = cudaMemcpy(out, d + 8, sizeof(double) * (what >= 2 ? 7 : 3), cudaMemcpyDeviceToHost);
    dfree(d);
    dfree(dorb.buf);
    if (e != cudaSuccess) return fail(err, errlen, B200_ECUDA, "primitive kernel failed: %s", cudaGetErrorString(e));
    return B200_OK;
}
