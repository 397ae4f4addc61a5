// Runtime, storage, initialisers, host materialisation and metadata-only views.
// Reference behaviour restated from arraylib/src/arraylib.c (line ranges cited per function);
// the design (device pool, stream ordering, error channel) is new.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <thread>
#include <unordered_map>

#include "al_internal.h"

namespace al {

static Runtime g_rt;
static thread_local std::string g_err;
static std::recursive_mutex g_api_mutex;

Runtime& rt() { return g_rt; }
// A host thread that enters the library for the first time has CUDA device 0 current, whatever device the process
// was bound to: events created and recorded from such a thread then belong to the wrong device ("invalid resource
// handle" on ranks > 0 as soon as a second Python thread calls to_numpy). Every entry point re-binds its thread.
static thread_local int t_bound_device = -1;
void bind_thread() {
    if (g_rt.device >= 0 && t_bound_device != g_rt.device) {
        cudaSetDevice(g_rt.device);
        t_bound_device = g_rt.device;
    }
}
std::recursive_mutex& api_mutex() { return g_api_mutex; }
const std::string& last_error_string() { return g_err; }
void restore_error(const std::string& s) { g_err = s; }

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
void clear_error() { g_err.clear(); }

bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    const char* cls = (e == cudaErrorMemoryAllocation) ? "MemoryError" : "RuntimeError";
    set_error("%s: CUDA failure %s in %s", cls, cudaGetErrorString(e), what);
    return false;
}

bool check_launch(const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return true;
    set_error("RuntimeError: launch of %s failed: %s", name, cudaGetErrorString(e));
    return false;
}

bool ensure_init() { return g_rt.device >= 0 || al_init(-1) == 0; }

size_t count_of(const size_t* shape, size_t ndim) {
    size_t n = 1;
    for (size_t i = 0; i < ndim; i++) n *= shape[i] ? shape[i] : 1;
    return n;
}

// Asynchronous uploads (array_fill_async) run on the H2D stream. The compute stream does NOT wait for an upload
// when it is enqueued -- that would hold the next kernel back until every upload queued so far has landed --
// but when the uploaded buffer is first USED: every entry point validates its operands, and valid() makes the
// compute stream wait for the operand's own upload event. So `upload a, b, X; a + b` starts as soon as a and b
// are resident while X is still travelling (and the D2H of a + b overlaps it: PCIe is full duplex).
static std::unordered_map<const Data*, cudaEvent_t> g_uploads;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}

static void await_upload(const Data* d) {
    if (g_uploads.empty()) return;
    auto it = g_uploads.find(d);
    if (it == g_uploads.end()) return;
    cudaStreamWaitEvent(g_rt.stream, it->second, 0);
    g_event_pool.push_back(it->second);
    g_uploads.erase(it);
}

static void free_layout(Layout* lay);

// arraylib.c:223-231 without the entry-point prologue: internal callers (error paths, temporaries, the download
// reaper) release arrays through this so that the thread's error message survives the clean-up.
void release_array(NDArray* a) {
    if (!a) return;
    if (a->data && __atomic_sub_fetch(&a->data->refs, 1, __ATOMIC_ACQ_REL) == 0) {
        await_upload(a->data);  // never used: the block must still not be recycled before its upload has landed
        g_rt.bytes_in_use -= a->data->size * sizeof(float);
        device_free(a->data->mem);
        free(a->data);
    }
    free_layout(a->lay);
    free(a);
}

// Blocks that an in-flight asynchronous download still reads: they hold one extra reference until their copy has
// completed (an event recorded behind it on the D2H stream), so the allocator cannot hand them to a new owner
// early. Finished entries are reaped whenever a new download is queued, whenever the allocator has to go to the
// driver, and in al_sync() -- a pipeline that never calls al_sync() holds at most the downloads in flight.
struct PendingDownload {
    NDArray* hold;
    cudaEvent_t done;
};
static std::vector<PendingDownload> g_pending_downloads;

static void reap_downloads(bool wait_for_all) {
    if (g_pending_downloads.empty()) return;
    size_t kept = 0;
    for (size_t i = 0; i < g_pending_downloads.size(); i++) {
        PendingDownload& p = g_pending_downloads[i];
        bool finished = wait_for_all;
        if (!finished) {
            cudaError_t e = cudaEventQuery(p.done);
            if (e != cudaSuccess) cudaGetLastError();  // cudaErrorNotReady is not sticky, but keep the slot clean
            finished = (e == cudaSuccess);
        }
        if (finished) {
            release_array(p.hold);
            g_event_pool.push_back(p.done);
        } else {
            g_pending_downloads[kept++] = p;
        }
    }
    g_pending_downloads.resize(kept);
}

// In-place writers (setters, array_write_base, in-place collectives) run on the compute stream; an asynchronous
// download of the same buffer runs on the D2H stream and is only ordered AFTER earlier compute work. Make the
// compute stream wait for the pending downloads of `d` so that the host never sees a mix of old and new values.
void await_downloads(const Data* d) {
    for (const PendingDownload& p : g_pending_downloads)
        if (p.hold->data == d) cudaStreamWaitEvent(g_rt.stream, p.done, 0);
}

bool valid(const NDArray* a, const char* who) {
    if (a && a->data && a->lay && a->lay->ndim > 0 && a->ptr) {
        await_upload(a->data);
        return true;
    }
    set_error("TypeError: %s got a NULL or malformed array", who);
    return false;
}

bool valid_mut(const NDArray* a, const char* who) {
    if (!valid(a, who)) return false;
    await_downloads(a->data);
    return true;
}

size_t view_offset(const NDArray* a) { return (size_t)(a->ptr - a->data->mem); }

// arraylib.c:91-98
bool is_contiguous(const NDArray* a) {
    size_t expect = 1;
    for (size_t i = a->lay->ndim; i-- > 0;) {
        if (a->lay->stride[i] != expect) return false;
        expect *= a->lay->shape[i];
    }
    return true;
}

// ---- device memory: size-class cache over cudaMalloc ---------------------------------------------
// Every kernel and copy of the library runs on ONE stream, so a block freed by the host can be
// handed to the next allocation immediately: whatever still reads or writes it was enqueued
// earlier on that same stream. Reusing blocks of the exact (rounded) size avoids both the
// driver's synchronous cudaMalloc/cudaFree and the VA remapping the stream-ordered pool performs
// when multi-GiB requests of different sizes alternate (measured: 67 ms of idle GPU per bench
// step with cudaMallocAsync versus none with this cache).
static std::multimap<size_t, void*> g_free_blocks;
static std::unordered_map<void*, size_t> g_block_size;
static size_t g_cached_bytes = 0;

static size_t round_block(size_t bytes) {
    const size_t q = bytes >= (1u << 20) ? (size_t)2 << 20 : 512;
    return (bytes + q - 1) / q * q;
}

static void release_cached_blocks() {
    if (g_free_blocks.empty()) return;
    cudaStreamSynchronize(g_rt.stream);
    for (auto& kv : g_free_blocks) {
        cudaFree(kv.second);
        g_block_size.erase(kv.second);
    }
    g_free_blocks.clear();
    g_cached_bytes = 0;
}

// The cache may hold at most `cache_limit` bytes of idle blocks (default: half of the device's memory, so a
// workload whose sizes drift cannot strand the whole HBM in blocks nobody asks for again; `cache_limit_mb`
// tuning knob). Over the limit the LARGEST idle blocks go back to the driver first: they are the ones another
// size class cannot reuse. cudaFree waits for the device, which is what makes the early release safe.
static size_t g_cache_limit = 0;

static void trim_cache_to_limit() {
    if (g_cache_limit == 0) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        g_cache_limit = total_b / 2;
    }
    while (g_cached_bytes > g_cache_limit && !g_free_blocks.empty()) {
        auto last = std::prev(g_free_blocks.end());
        cudaFree(last->second);
        g_block_size.erase(last->second);
        g_cached_bytes -= last->first;
        g_free_blocks.erase(last);
    }
}

void set_cache_limit(size_t bytes) {
    g_cache_limit = bytes;  // 0 = back to the default (half of the device's memory)
    trim_cache_to_limit();
}
size_t cached_bytes() { return g_cached_bytes; }

void* device_alloc(size_t bytes) {
    const size_t want = round_block(bytes ? bytes : 1);
    auto it = g_free_blocks.lower_bound(want);
    if (it != g_free_blocks.end() && it->first <= want + want / 8) {  // at most 12.5% slack
        void* p = it->second;
        g_cached_bytes -= it->first;
        g_free_blocks.erase(it);
        return p;
    }
    reap_downloads(false);  // finished downloads hand their blocks back to the cache first
    it = g_free_blocks.lower_bound(want);
    if (it != g_free_blocks.end() && it->first <= want + want / 8) {
        void* p = it->second;
        g_cached_bytes -= it->first;
        g_free_blocks.erase(it);
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // give the cache back to the driver and retry once
        cudaGetLastError();
        release_cached_blocks();
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("MemoryError: failed to allocate %zu bytes of device memory (%s)", want, cudaGetErrorString(e));
        return nullptr;
    }
    g_block_size[p] = want;
    return p;
}

void device_free(void* p) {
    if (!p) return;
    auto it = g_block_size.find(p);
    if (it == g_block_size.end()) return;
    g_free_blocks.emplace(it->second, p);
    g_cached_bytes += it->second;
    trim_cache_to_limit();
}

static Layout* make_layout(const size_t* shape, const size_t* stride, size_t ndim) {
    Layout* lay = (Layout*)malloc(sizeof(Layout));
    lay->ndim = ndim;
    lay->shape = (size_t*)malloc(sizeof(size_t) * ndim);
    lay->stride = (size_t*)malloc(sizeof(size_t) * ndim);
    memcpy(lay->shape, shape, sizeof(size_t) * ndim);
    if (stride) {
        memcpy(lay->stride, stride, sizeof(size_t) * ndim);
    } else {  // row-major (arraylib.c:102-107)
        size_t run = 1;
        for (size_t i = ndim; i-- > 0;) {
            lay->stride[i] = run;
            run *= shape[i];
        }
    }
    return lay;
}

static void free_layout(Layout* lay) {
    if (!lay) return;
    free(lay->shape);
    free(lay->stride);
    free(lay);
}

// arraylib.c:129-137 + 212-221, with the buffer taken from the stream-ordered device pool.
NDArray* new_array(const size_t* shape, size_t ndim) {
    if (!ensure_init()) return nullptr;
    AL_REQUIRE(ndim > 0 && shape, "ValueError: non-positive ndim");
    for (size_t i = 0; i < ndim; i++)
        AL_REQUIRE(shape[i] > 0, "ValueError: zero extent in shape (axis %zu)", i);
    size_t n = count_of(shape, ndim);
    float* mem = (float*)device_alloc(n * sizeof(float));
    if (!mem) return nullptr;
    Data* d = (Data*)malloc(sizeof(Data));
    d->mem = mem;
    d->size = n;
    d->refs = 1;
    g_rt.bytes_in_use += n * sizeof(float);
    NDArray* a = (NDArray*)malloc(sizeof(NDArray));
    a->data = d;
    a->ptr = mem;
    a->lay = make_layout(shape, nullptr, ndim);
    a->view = false;
    return a;
}

// arraylib.c:237-246 / 258-269: a second NDArray over the same Data.
NDArray* new_view(const NDArray* src, const size_t* shape, const size_t* stride, size_t ndim,
                  size_t extra_offset) {
    NDArray* a = (NDArray*)malloc(sizeof(NDArray));
    a->data = src->data;
    __atomic_add_fetch(&src->data->refs, 1, __ATOMIC_RELAXED);
    a->ptr = src->ptr + extra_offset;
    a->lay = make_layout(shape, stride, ndim);
    a->view = true;
    return a;
}

// ---- fills -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_const_kernel(float* __restrict__ p, size_t n, float v) {
    // scalar head up to the first 16-byte boundary, float4 body, scalar tail
    size_t head = ((16 - ((uintptr_t)p & 15)) & 15) >> 2;
    if (head > n) head = n;
    const size_t n4 = (n - head) >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4* body = reinterpret_cast<float4*>(p + head);
    const float4 vv = make_float4(v, v, v, v);
    for (size_t i = tid; i < n4; i += stride) body[i] = vv;
    if (tid < head) p[tid] = v;
    for (size_t i = head + (n4 << 2) + tid; i < n; i += stride) p[i] = v;
}

// value(i) = float(start + (first + i) * step) with the product formed in exact integers
__global__ void __launch_bounds__(256)
fill_iota_kernel(float* __restrict__ p, size_t n, unsigned long long start, unsigned long long step) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = (float)(start + i * step);
}

// linspace body (arraylib.c:313-314): p[i] = start + p[i] * dx, fused like the reference build
__global__ void __launch_bounds__(256) axpb_kernel(float* __restrict__ p, size_t n, float dx, float start) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = fmaf(p[i], dx, start);
}

static int fill_grid(size_t work) {
    size_t blocks = (work + 255) / 256;
    size_t cap = (size_t)kNumSMs * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

bool fill_const(float* p, size_t n, float v) {
    fill_const_kernel<<<fill_grid(n / 4 + 4), 256, 0, g_rt.stream>>>(p, n, v);
    note_launch("fill_const");
    return check_launch("fill_const");
}


// ---- pageable host memory <-> device through a pinned ring ---------------------------------------------------
// The reference's from_numpy / to_numpy hand over ordinary (pageable) numpy buffers (core.py:340-363). A plain
// cudaMemcpy of pageable memory is staged by the driver on ONE host thread (measured here: ~17 GB/s on the wire
// against 55 GB/s for pinned memory). Large pageable transfers therefore go through the library's own ring of
// pinned slots: helper threads copy between the caller's buffer and a slot while the DMA engine moves the
// previous slot, uploads and downloads use separate rings and streams (PCIe is full duplex), and the API lock is
// NOT held while bytes move, so one host thread can upload while another downloads.
extern "C" void al_host_copy_stream(char* dst, const char* src, size_t n);  // hostcopy.cpp: streaming (non-temporal) stores
static inline void host_copy(char* dst, const char* src, size_t n) {
    if (g_rt.host_copy_nt) al_host_copy_stream(dst, src, n);
    else memcpy(dst, src, n);
}

struct CopyPool {
    struct Job {
        char* dst;
        const char* src;
        size_t n;
        std::atomic<int>* pending;
    };
    std::mutex m;
    std::condition_variable cv;
    std::deque<Job> q;
    int helpers = 0;

    void worker() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return !q.empty(); });
                j = q.front();
                q.pop_front();
            }
            host_copy(j.dst, j.src, j.n);
            j.pending->fetch_sub(1, std::memory_order_release);
        }
    }
    void start() {
        if (helpers) return;
        // half of this rank's share of the host cores (the ranks of one box split them): 8 on a 16-core single-GPU
        // box, 2 per rank with 8 ranks on 32 cores
        unsigned hw = std::thread::hardware_concurrency();
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        int ranks = lw ? atoi(lw) : 1;
        if (ranks < 1) ranks = 1;
        const char* env = getenv("ARRAYLIB_COPY_THREADS");
        helpers = env ? atoi(env) : (int)(hw / (2 * (unsigned)ranks));
        if (helpers < 1) helpers = 1;
        if (helpers > 12) helpers = 12;
        for (int i = 0; i < helpers; i++) std::thread([this] { worker(); }).detach();
    }
    // dst/src may be pageable or pinned; returns when every byte has been copied
    void copy(char* dst, const char* src, size_t n) {
        const size_t piece = (size_t)4 << 20;
        size_t parts = (n + piece - 1) / piece;
        if (parts > (size_t)helpers + 1) parts = (size_t)helpers + 1;
        if (parts <= 1) {
            memcpy(dst, src, n);
            return;
        }
        const size_t each = ((n / parts) + 4095) & ~(size_t)4095;
        std::atomic<int> pending{0};
        size_t off = each;  // the caller's own share is [0, each)
        {
            std::lock_guard<std::mutex> lk(m);
            while (off < n) {
                const size_t len = std::min(each, n - off);
                pending.fetch_add(1, std::memory_order_relaxed);
                q.push_back({dst + off, src + off, len, &pending});
                off += len;
            }
        }
        cv.notify_all();
        host_copy(dst, src, std::min(each, n));
        while (pending.load(std::memory_order_acquire) > 0) std::this_thread::yield();
    }
};
static CopyPool* g_copy_pool = nullptr;  // leaked on purpose: detached helpers may outlive static destruction

constexpr int kRingSlots = 3;
// slot size of the staging rings: ARRAYLIB_RING_SLOT_MB (1..256), default 32 MiB
static size_t ring_slot_bytes() {
    static size_t bytes = 0;
    if (bytes == 0) {
        const char* env = getenv("ARRAYLIB_RING_SLOT_MB");
        long mb = env ? atol(env) : 32;
        if (mb < 1) mb = 1;
        if (mb > 256) mb = 256;
        bytes = (size_t)mb << 20;
    }
    return bytes;
}
#define kRingSlotBytes (ring_slot_bytes())
constexpr size_t kStagedMinBytes = (size_t)4 << 20;  // smaller transfers: the driver's own staging is as good
struct PinnedRing {
    std::mutex busy;  // one transfer per direction at a time
    char* slot[kRingSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t done[kRingSlots] = {nullptr, nullptr, nullptr};
    bool in_flight[kRingSlots] = {false, false, false};  // a DMA out of this slot may still be running (uploads return early)
    bool ready() {
        if (slot[0]) return true;
        for (int i = 0; i < kRingSlots; i++) {
            if (cudaHostAlloc((void**)&slot[i], kRingSlotBytes, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                for (int j = 0; j <= i; j++) {
                    if (slot[j]) cudaFreeHost(slot[j]);
                    slot[j] = nullptr;
                }
                return false;
            }
        }
        return true;
    }
};
static PinnedRing g_up_ring, g_down_ring;

static bool is_pageable(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

static bool staging_enabled() { return g_rt.host_staging != 0; }

// Upload `bytes` from pageable `src` to `dev` on the H2D stream. The caller has already made the H2D stream wait
// for the compute stream. Returns once the SOURCE has been read completely (the last DMA may still be in flight).
static bool staged_upload(float* dev, const char* src, size_t bytes) {
    std::lock_guard<std::mutex> ring(g_up_ring.busy);
    if (!g_up_ring.ready()) return set_error("MemoryError: cannot allocate the pinned upload ring"), false;
    size_t off = 0;
    for (int c = 0; off < bytes; c++) {
        const int s = c % kRingSlots;
        const size_t len = std::min(kRingSlotBytes, bytes - off);
        // the slot may still be read by an earlier DMA -- of this transfer, or of the PREVIOUS one, which returned as
        // soon as its source had been read
        if (g_up_ring.in_flight[s] && !cuda_ok(cudaEventSynchronize(g_up_ring.done[s]), "cudaEventSynchronize(upload ring)")) return false;
        g_copy_pool->copy(g_up_ring.slot[s], src + off, len);
        if (!cuda_ok(cudaMemcpyAsync((char*)dev + off, g_up_ring.slot[s], len, cudaMemcpyHostToDevice, g_rt.h2d_stream),
                     "cudaMemcpyAsync(H2D ring)") ||
            !cuda_ok(cudaEventRecord(g_up_ring.done[s], g_rt.h2d_stream), "cudaEventRecord"))
            return false;
        g_up_ring.in_flight[s] = true;
        off += len;
    }
    return true;
}

// Download `bytes` from `dev` into pageable `dst` on the D2H stream (which already waits for the compute stream).
// Returns when `dst` is complete.
static bool staged_download(char* dst, const float* dev, size_t bytes) {
    std::lock_guard<std::mutex> ring(g_down_ring.busy);
    if (!g_down_ring.ready()) return set_error("MemoryError: cannot allocate the pinned download ring"), false;
    const size_t chunks = (bytes + kRingSlotBytes - 1) / kRingSlotBytes;
    auto issue = [&](size_t c) {
        const size_t off = c * kRingSlotBytes, len = std::min(kRingSlotBytes, bytes - off);
        const int s = (int)(c % kRingSlots);
        return cuda_ok(cudaMemcpyAsync(g_down_ring.slot[s], (const char*)dev + off, len, cudaMemcpyDeviceToHost, g_rt.d2h_stream),
                       "cudaMemcpyAsync(D2H ring)") &&
               cuda_ok(cudaEventRecord(g_down_ring.done[s], g_rt.d2h_stream), "cudaEventRecord");
    };
    for (size_t c = 0; c < chunks && c < (size_t)kRingSlots; c++)
        if (!issue(c)) return false;
    for (size_t c = 0; c < chunks; c++) {
        const size_t off = c * kRingSlotBytes, len = std::min(kRingSlotBytes, bytes - off);
        const int s = (int)(c % kRingSlots);
        if (!cuda_ok(cudaEventSynchronize(g_down_ring.done[s]), "cudaEventSynchronize(download ring)")) return false;
        g_copy_pool->copy(dst + off, g_down_ring.slot[s], len);
        if (c + kRingSlots < chunks && !issue(c + kRingSlots)) return false;
    }
    return true;
}

static bool use_staging(const void* host, size_t bytes) {
    if (!staging_enabled() || bytes < kStagedMinBytes || !is_pageable(host)) return false;
    if (!g_copy_pool) {
        g_copy_pool = new CopyPool();
        g_copy_pool->start();
    }
    return true;
}

}  // namespace al

using namespace al;

extern "C" {

// ---- runtime -------------------------------------------------------------------------------------
int al_init(int device) {
    AL_ENTER;
    if (g_rt.device >= 0) return 0;
    if (device < 0) {
        const char* e = getenv("ARRAYLIB_DEVICE");
        if (!e) e = getenv("LOCAL_RANK");
        device = e ? atoi(e) : 0;
    }
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("RuntimeError: no CUDA device available (%s); arraylib_b200 has no CPU fallback",
                  err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
        return -1;
    }
    if (device >= count) device = device % count;
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return -1;
    if (!cuda_ok(cudaStreamCreateWithFlags(&g_rt.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return -1;
    if (!cuda_ok(cudaStreamCreateWithFlags(&g_rt.h2d_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return -1;
    if (!cuda_ok(cudaStreamCreateWithFlags(&g_rt.d2h_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return -1;
    if (!cuda_ok(cudaEventCreateWithFlags(&g_rt.ev_copy, cudaEventDisableTiming), "cudaEventCreate")) return -1;
    if (!cuda_ok(cudaEventCreate(&g_rt.ev_start), "cudaEventCreate")) return -1;
    if (!cuda_ok(cudaEventCreate(&g_rt.ev_stop), "cudaEventCreate")) return -1;
    g_rt.device = device;
    t_bound_device = device;
    // ARRAYLIB_TUNING="key=value,key=value": the al_set_tuning knobs from the environment
    if (const char* env = getenv("ARRAYLIB_TUNING")) {
        std::string spec(env);
        size_t pos = 0;
        while (pos < spec.size()) {
            size_t end = spec.find(',', pos);
            if (end == std::string::npos) end = spec.size();
            const std::string item = spec.substr(pos, end - pos);
            const size_t eq = item.find('=');
            if (eq != std::string::npos && al_set_tuning(item.substr(0, eq).c_str(), atol(item.c_str() + eq + 1)) != 0)
                fprintf(stderr, "arraylib_b200: ignoring ARRAYLIB_TUNING entry '%s'\n", item.c_str());
            pos = end + 1;
        }
    }
    clear_error();
    return 0;
}

int al_device(void) { return g_rt.device; }

int al_device_count(void) {  // CUDA devices visible to this process (0 when there is no usable driver)
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int al_sync(void) {
    AL_ENTER;
    if (g_rt.device < 0) return 0;
    bool ok = cuda_ok(cudaStreamSynchronize(g_rt.h2d_stream), "cudaStreamSynchronize(h2d)") &&
              cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize") &&
              cuda_ok(cudaStreamSynchronize(g_rt.d2h_stream), "cudaStreamSynchronize(d2h)");
    reap_downloads(true);  // the D2H stream is idle (or broken): nothing reads the held blocks any more
    for (auto& kv : g_uploads) g_event_pool.push_back(kv.second);  // the H2D stream is idle: nothing left to wait for
    g_uploads.clear();
    if (ok && comm_error_pending()) ok = false;  // a peer-memory collective gave up waiting for another rank
    return ok ? 0 : -1;
}

void* al_stream(void) { return (void*)g_rt.stream; }
const char* al_last_error(void) { return g_err.c_str(); }
const char* al_version(void) { return "arraylib_b200 0.1 (sm_100a)"; }
uint64_t al_launch_count(void) { return g_rt.launches; }
const char* al_last_kernel(void) { return g_rt.last_kernel; }
uint64_t al_bytes_in_use(void) { return g_rt.bytes_in_use; }
uint64_t al_bytes_cached(void) { return cached_bytes(); }

int al_timer_start(void) {
    AL_ENTER;
    if (!ensure_init()) return -1;
    return cuda_ok(cudaEventRecord(g_rt.ev_start, g_rt.stream), "cudaEventRecord") ? 0 : -1;
}

int al_timer_stop(float* ms) {
    AL_ENTER;
    if (!ensure_init()) return -1;
    if (!cuda_ok(cudaEventRecord(g_rt.ev_stop, g_rt.stream), "cudaEventRecord")) return -1;
    if (!cuda_ok(cudaEventSynchronize(g_rt.ev_stop), "cudaEventSynchronize")) return -1;
    return cuda_ok(cudaEventElapsedTime(ms, g_rt.ev_start, g_rt.ev_stop), "cudaEventElapsedTime") ? 0 : -1;
}

// A small pool of events so that a caller (bench.py) can bracket individual operations on the
// library stream without synchronising between them: record now, read elapsed times after al_sync.
static std::vector<cudaEvent_t> g_events;

int al_event_record(int idx) {
    AL_ENTER;
    if (!ensure_init()) return -1;
    if (idx < 0 || idx >= 1 << 16) return set_error("IndexError: event index %d out of range", idx), -1;
    while ((int)g_events.size() <= idx) {
        cudaEvent_t e;
        if (!cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return -1;
        g_events.push_back(e);
    }
    return cuda_ok(cudaEventRecord(g_events[idx], g_rt.stream), "cudaEventRecord") ? 0 : -1;
}

int al_event_elapsed(int first, int second, float* ms) {
    AL_ENTER;
    if (first < 0 || second < 0 || first >= (int)g_events.size() || second >= (int)g_events.size())
        return set_error("IndexError: event was never recorded"), -1;
    if (!cuda_ok(cudaEventSynchronize(g_events[second]), "cudaEventSynchronize")) return -1;
    return cuda_ok(cudaEventElapsedTime(ms, g_events[first], g_events[second]), "cudaEventElapsedTime") ? 0 : -1;
}

int al_trim_pool(void) {
    AL_ENTER;
    if (!ensure_init()) return -1;
    release_cached_blocks();
    return 0;
}

void* al_host_alloc(size_t bytes) {
    AL_ENTER;
    if (!ensure_init()) return nullptr;
    void* p = nullptr;
    if (!cuda_ok(cudaHostAlloc(&p, bytes, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
    return p;
}

int al_host_free(void* p) {
    AL_ENTER;
    return cuda_ok(cudaFreeHost(p), "cudaFreeHost") ? 0 : -1;
}

int al_set_tuning(const char* key, long value) {
    AL_ENTER;
    std::string k = key ? key : "";
    if (k == "stream_store_min_bytes") g_rt.stream_store_min_bytes = value;
    else if (k == "ew_unroll") (void)value;  // swept once (profiles/r1_micro_stream_variants.log), now fixed at 4
    else if (k == "load_policy") g_rt.load_policy = value;
    else if (k == "cache_limit_mb") set_cache_limit((size_t)(value < 0 ? 0 : value) << 20);
    else if (k == "force_general") g_rt.force_general = value;
    else if (k == "reduce_splits") g_rt.reduce_splits = value;
    else if (k == "reduce_sequential") g_rt.reduce_sequential = value;
    else if (k == "skip_zero_sign_fix") g_rt.skip_zero_sign_fix = value;
    else if (k == "host_staging") g_rt.host_staging = value;
    else if (k == "host_copy_nt") g_rt.host_copy_nt = value;
    else if (k == "reduce_loads_per_lane") g_rt.reduce_loads_per_lane = value;
    else if (k == "disable_window") g_rt.disable_window = value;
    else if (k == "window_shfl") g_rt.window_shfl = value;
    else if (k == "window_ctas_per_sm") g_rt.window_ctas_per_sm = value;
    else if (k == "window_whole_blocks") g_rt.window_whole_blocks = value;
    else if (k == "tile64") g_rt.tile64 = value;
    else if (k == "tile_tq") g_rt.tile_tq = value;
    else if (k == "interleave") g_rt.interleave = value;
    else if (k == "stream_tma") g_rt.stream_tma = value;
    else if (k == "chain_block") g_rt.chain_block = value;
    else if (k == "peer_collective") g_rt.peer_collective = value;
    else if (k == "reduce_peel") g_rt.reduce_peel = value;
    else if (k == "outer_tile") g_rt.outer_tile = value;
    else if (k == "period_plan") g_rt.period_plan = value;
    else if (k == "window_any_l16") g_rt.window_any_l16 = value;
    else if (k == "heavy_stream") g_rt.heavy_stream = value;
    else if (k == "heavy_ctas_per_sm") g_rt.heavy_ctas_per_sm = value > 0 ? value : 0;
    else {
        set_error("KeyError: unknown tuning key '%s'", k.c_str());
        return -1;
    }
    return 0;
}

// ---- low-level helpers the reference header also declares (arraylib.h "MEMORY ALLOCATORS", "DATA",
//      "LAYOUT"); the reference's Python never calls them, they exist so that its cdef binds completely --
void* alloc(size_t size) {  // arraylib.c:55-60 (host memory)
    AL_ENTER;
    if (size == 0) return set_error("ValueError: non-positive size in malloc."), nullptr;
    void* p = malloc(size);
    if (!p) set_error("MemoryError: failed to allocate memory.");
    return p;
}

Data* data_empty(size_t size) {  // arraylib.c:129-137, with `mem` in device memory
    AL_ENTER;
    if (!ensure_init()) return nullptr;
    if (size == 0) return set_error("ValueError: non-postive size in data_empty."), nullptr;
    float* mem = (float*)device_alloc(size * sizeof(float));
    if (!mem) return nullptr;
    Data* d = (Data*)malloc(sizeof(Data));
    d->mem = mem;
    d->size = size;
    d->refs = 1;
    g_rt.bytes_in_use += size * sizeof(float);
    return d;
}

Layout* layout_alloc(size_t ndim) {  // arraylib.c:143-149
    AL_ENTER;
    if (ndim == 0) return set_error("ValueError: non-positive ndim"), nullptr;
    Layout* lay = (Layout*)malloc(sizeof(Layout));
    lay->ndim = ndim;
    lay->shape = (size_t*)calloc(ndim, sizeof(size_t));
    lay->stride = (size_t*)calloc(ndim, sizeof(size_t));
    return lay;
}

Layout* layout_copy(Layout* dst, const Layout* src) {  // arraylib.c:151-156 (dst must hold src->ndim entries)
    AL_ENTER;
    if (!dst || !src) return set_error("TypeError: layout_copy on NULL"), nullptr;
    dst->ndim = src->ndim;
    memcpy(dst->shape, src->shape, sizeof(size_t) * src->ndim);
    memcpy(dst->stride, src->stride, sizeof(size_t) * src->ndim);
    return dst;
}

void layout_free(Layout* lay) { free_layout(lay); }  // arraylib.c:158-162

// arraylib.c:180-206: right-aligned broadcast of n layouts to a common rank; stride 0 on broadcast axes.
Layout** layout_broadcast(const Layout** lays, size_t nlay) {
    AL_ENTER;
    if (!lays || nlay == 0) return set_error("TypeError: layout_broadcast on NULL"), nullptr;
    size_t nd = 0;
    for (size_t j = 0; j < nlay; j++) nd = lays[j]->ndim > nd ? lays[j]->ndim : nd;
    Layout** out = (Layout**)malloc(sizeof(Layout*) * nlay);
    for (size_t j = 0; j < nlay; j++) out[j] = layout_alloc(nd);
    for (size_t k = 0; k < nd; k++) {
        size_t e = 0;
        for (size_t j = 0; j < nlay; j++) {
            const size_t ej = k < lays[j]->ndim ? lays[j]->shape[lays[j]->ndim - 1 - k] : 1;
            e = ej > e ? ej : e;
        }
        for (size_t j = 0; j < nlay; j++) {
            const bool present = k < lays[j]->ndim && lays[j]->shape[lays[j]->ndim - 1 - k] == e;
            out[j]->shape[nd - 1 - k] = e;
            out[j]->stride[nd - 1 - k] = present ? lays[j]->stride[lays[j]->ndim - 1 - k] : 0;
        }
    }
    return out;
}

// arraylib.c:491-501, 569-589 take a C function pointer: a host callback cannot run on the device.
NDArray* array_scalar_op(binop, const NDArray*, f32) {
    AL_ENTER;
    set_error("NotImplementedError: host callbacks cannot run on the device; use array_scalar_op_named with an al_binop entry");
    return nullptr;
}
NDArray* array_array_op(binop, const NDArray*, const NDArray*) {
    AL_ENTER;
    set_error("NotImplementedError: host callbacks cannot run on the device; use array_array_op_named with an al_binop entry");
    return nullptr;
}

// ---- storage -------------------------------------------------------------------------------------
NDArray* array_empty(const size_t* shape, size_t ndim) {
    AL_ENTER;
    return new_array(shape, ndim);
}

// arraylib.c:223-231. The device block is returned to the pool in stream order, so kernels
// already enqueued against it finish first.
void array_free(NDArray* a) {
    AL_ENTER;
    if (!a) {
        set_error("TypeError: array_free on NULL");
        return;
    }
    release_array(a);
}

NDArray* array_shallow_copy(const NDArray* src) {
    AL_ENTER;
    if (!valid(src, "array_shallow_copy")) return nullptr;
    return new_view(src, src->lay->shape, src->lay->stride, src->lay->ndim, 0);
}

// ---- initialisers --------------------------------------------------------------------------------
NDArray* array_full(const size_t* shape, size_t ndim, f32 value) {
    AL_ENTER;
    NDArray* a = new_array(shape, ndim);
    if (!a) return nullptr;
    if (!fill_const(a->ptr, a->data->size, value)) {
        array_free(a);
        return nullptr;
    }
    return a;
}

NDArray* array_zeros(const size_t* shape, size_t ndim) { return array_full(shape, ndim, 0.0f); }
NDArray* array_ones(const size_t* shape, size_t ndim) { return array_full(shape, ndim, 1.0f); }

// arraylib.c:282-288 -- `elems` is a host buffer; this is the H2D upload path.
NDArray* array_fill(f32* elems, const size_t* shape, size_t ndim) {
    std::unique_lock<std::recursive_mutex> api(api_mutex());
    bind_thread();
    clear_error();
    AL_REQUIRE(elems, "TypeError: array_fill got a NULL host buffer");
    NDArray* a = new_array(shape, ndim);
    if (!a) return nullptr;
    const size_t bytes = a->data->size * sizeof(float);
    bool ok;
    if (use_staging(elems, bytes)) {
        // the block may have just been released by work still queued on the compute stream
        ok = cuda_ok(cudaEventRecord(g_rt.ev_copy, g_rt.stream), "cudaEventRecord") &&
             cuda_ok(cudaStreamWaitEvent(g_rt.h2d_stream, g_rt.ev_copy, 0), "cudaStreamWaitEvent");
        if (ok) {
            api.unlock();  // other threads may launch kernels / download while the bytes move
            ok = staged_upload(a->ptr, (const char*)elems, bytes);
            api.lock();
        }
        if (ok) {  // like array_fill_async: the compute stream waits for the last DMA when the array is first used
            cudaEvent_t landed = take_event();
            ok = landed && cuda_ok(cudaEventRecord(landed, g_rt.h2d_stream), "cudaEventRecord");
            if (ok) g_uploads[a->data] = landed;
            else if (landed) g_event_pool.push_back(landed);
        }
    } else {
        ok = cuda_ok(cudaMemcpyAsync(a->ptr, elems, bytes, cudaMemcpyHostToDevice, g_rt.stream), "cudaMemcpyAsync(H2D)") &&
             cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize");
    }
    if (!ok) {
        release_array(a);
        return nullptr;
    }
    return a;
}

// Asynchronous upload (extension): the copy runs on a dedicated H2D stream; the compute stream waits for it
// when the array is first used, so the call returns at once and overlaps with downloads, kernels and the
// uploads of other arrays. `elems` must be pinned (al_host_alloc) and stay untouched until al_sync().
NDArray* array_fill_async(f32* elems, const size_t* shape, size_t ndim) {
    AL_ENTER;
    AL_REQUIRE(elems, "TypeError: array_fill_async got a NULL host buffer");
    NDArray* a = new_array(shape, ndim);
    if (!a) return nullptr;
    // the block may have just been released by work still queued on the compute stream
    bool ok = cuda_ok(cudaEventRecord(g_rt.ev_copy, g_rt.stream), "cudaEventRecord") &&
              cuda_ok(cudaStreamWaitEvent(g_rt.h2d_stream, g_rt.ev_copy, 0), "cudaStreamWaitEvent") &&
              cuda_ok(cudaMemcpyAsync(a->ptr, elems, a->data->size * sizeof(float), cudaMemcpyHostToDevice,
                                      g_rt.h2d_stream), "cudaMemcpyAsync(H2D)");
    if (ok) {  // the compute stream waits for this event when the array is first used (valid())
        cudaEvent_t landed = take_event();
        ok = landed && cuda_ok(cudaEventRecord(landed, g_rt.h2d_stream), "cudaEventRecord");
        if (ok) g_uploads[a->data] = landed;
        else if (landed) g_event_pool.push_back(landed);
    }
    if (!ok) {
        std::string keep = g_err;
        array_free(a);
        g_err = keep;
        return nullptr;
    }
    return a;
}

// Asynchronous download (extension): ordered after everything enqueued so far, executed on a
// dedicated D2H stream. `dst` must be pinned and is valid after al_sync().
int array_to_host_async(const NDArray* src, f32* dst) {
    AL_ENTER;
    if (!valid(src, "array_to_host_async")) return -1;
    if (!dst) return set_error("TypeError: array_to_host_async got a NULL destination"), -1;
    reap_downloads(false);
    size_t n = count_of(src->lay->shape, src->lay->ndim);
    NDArray* hold = is_contiguous(src) ? array_shallow_copy(src) : array_deep_copy(src);
    if (!hold) return -1;
    bool ok = cuda_ok(cudaEventRecord(g_rt.ev_copy, g_rt.stream), "cudaEventRecord") &&
              cuda_ok(cudaStreamWaitEvent(g_rt.d2h_stream, g_rt.ev_copy, 0), "cudaStreamWaitEvent") &&
              cuda_ok(cudaMemcpyAsync(dst, hold->ptr, n * sizeof(float), cudaMemcpyDeviceToHost, g_rt.d2h_stream),
                      "cudaMemcpyAsync(D2H)");
    cudaEvent_t done = take_event();
    if (ok) ok = done && cuda_ok(cudaEventRecord(done, g_rt.d2h_stream), "cudaEventRecord");
    if (ok) {
        g_pending_downloads.push_back({hold, done});  // released once `done` has fired (reap_downloads)
    } else {  // nothing was queued that still reads the block
        const std::string keep = g_err;
        if (done) g_event_pool.push_back(done);
        array_free(hold);
        g_err = keep;
    }
    return ok ? 0 : -1;
}

// arraylib.c:298-308. While every value and every partial sum stays <= 2^24 the reference's
// size_t -> f32 -> size_t round trip is exact and the sequence is start + i*step: that prefix is
// produced on the device in closed form. Past 2^24 the f32 add rounds (and, for step 1, sticks at
// 16777216 forever), so the tail replays the reference's recurrence step by step on the host and
// is uploaded; a fixed point is detected and filled on the device instead.
NDArray* array_arange(f32 start, f32 end, f32 step) {
    AL_ENTER;
    AL_REQUIRE(start < end && start >= 0, "ValueError: invalid array range.");
    size_t span = (size_t)(end - start), st = (size_t)step;
    AL_REQUIRE(st > 0, "ValueError: arange step must be >= 1");
    size_t total = (span + st - 1) / st;
    AL_REQUIRE(total > 0, "ValueError: empty arange");
    size_t shape1[1] = {total};
    NDArray* a = new_array(shape1, 1);
    if (!a) return nullptr;
    const size_t lim = (size_t)1 << 24;
    size_t s0 = (size_t)start;
    size_t n_exact = 0;
    if ((float)st == step && s0 <= lim) {
        n_exact = (lim - s0) / st + 1;
        if (n_exact > total) n_exact = total;
    }
    bool ok = true;
    if (n_exact) {
        fill_iota_kernel<<<fill_grid(n_exact), 256, 0, g_rt.stream>>>(a->ptr, n_exact, s0, st);
        note_launch("fill_iota");
        ok = check_launch("fill_iota");
    }
    if (ok && n_exact < total) {
        size_t running = n_exact ? s0 + (n_exact - 1) * st : s0;
        size_t i = n_exact ? n_exact - 1 : 0;  // element i holds `running`
        std::vector<float> host;
        const size_t chunk = (size_t)1 << 22;
        size_t first = i;  // index of host[0]
        for (; i < total && ok; i++) {
            host.push_back((float)running);
            size_t next = (size_t)((float)running + step);
            bool stuck = (next == running);
            running = next;
            if (host.size() == chunk || i + 1 == total || stuck) {
                ok = cuda_ok(cudaMemcpyAsync(a->ptr + first, host.data(), host.size() * sizeof(float),
                                             cudaMemcpyHostToDevice, g_rt.stream), "cudaMemcpyAsync(H2D)") &&
                     cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize");
                first = i + 1;
                host.clear();
                if (ok && stuck && i + 1 < total) {
                    ok = fill_const(a->ptr + i + 1, total - (i + 1), (float)running);
                    break;
                }
            }
        }
    }
    if (!ok) {
        std::string keep = g_err;
        array_free(a);
        g_err = keep;
        return nullptr;
    }
    return a;
}

// arraylib.c:310-316
NDArray* array_linspace(f32 start, f32 end, f32 n) {
    AL_ENTER;
    f32 dx = (end - start) / (n - 1);
    NDArray* a = array_arange(0, n, 1);
    if (!a) return nullptr;
    axpb_kernel<<<fill_grid(a->data->size), 256, 0, g_rt.stream>>>(a->ptr, a->data->size, dx, start);
    note_launch("linspace_axpb");
    if (!check_launch("linspace_axpb")) {
        array_free(a);
        return nullptr;
    }
    return a;
}

// ---- host materialisation ------------------------------------------------------------------------
int array_to_host(const NDArray* src, f32* dst) {
    std::unique_lock<std::recursive_mutex> api(api_mutex());
    bind_thread();
    clear_error();
    if (!valid(src, "array_to_host")) return -1;
    if (!dst) {
        set_error("TypeError: array_to_host got a NULL destination");
        return -1;
    }
    const size_t n = count_of(src->lay->shape, src->lay->ndim);
    // a contiguous source is read in place (through an extra reference that keeps the block alive), anything else
    // is gathered into a temporary first
    NDArray* from = is_contiguous(src) ? new_view(src, src->lay->shape, src->lay->stride, src->lay->ndim, 0) : array_deep_copy(src);
    if (!from) return -1;
    bool ok;
    if (use_staging(dst, n * sizeof(float))) {
        ok = cuda_ok(cudaEventRecord(g_rt.ev_copy, g_rt.stream), "cudaEventRecord") &&
             cuda_ok(cudaStreamWaitEvent(g_rt.d2h_stream, g_rt.ev_copy, 0), "cudaStreamWaitEvent");
        if (ok) {
            const float* dev = from->ptr;
            api.unlock();
            ok = staged_download((char*)dst, dev, n * sizeof(float));
            api.lock();
        }
    } else {
        ok = cuda_ok(cudaMemcpyAsync(dst, from->ptr, n * sizeof(float), cudaMemcpyDeviceToHost, g_rt.stream),
                     "cudaMemcpyAsync(D2H)") &&
             cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize");
    }
    release_array(from);
    return ok ? 0 : -1;
}

int array_read_base(const NDArray* src, size_t offset, size_t count, f32* dst) {
    AL_ENTER;
    if (!valid(src, "array_read_base")) return -1;
    if (offset + count > src->data->size) {
        set_error("IndexError: base read [%zu, %zu) exceeds %zu elements", offset, offset + count, src->data->size);
        return -1;
    }
    bool ok = cuda_ok(cudaMemcpyAsync(dst, src->data->mem + offset, count * sizeof(float), cudaMemcpyDeviceToHost,
                                      g_rt.stream), "cudaMemcpyAsync(D2H)") &&
              cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize");
    return ok ? 0 : -1;
}

int array_write_base(NDArray* dst, size_t offset, size_t count, const f32* src) {
    AL_ENTER;
    if (!valid_mut(dst, "array_write_base")) return -1;
    if (offset + count > dst->data->size) {
        set_error("IndexError: base write [%zu, %zu) exceeds %zu elements", offset, offset + count, dst->data->size);
        return -1;
    }
    bool ok = cuda_ok(cudaMemcpyAsync(dst->data->mem + offset, src, count * sizeof(float), cudaMemcpyHostToDevice,
                                      g_rt.stream), "cudaMemcpyAsync(H2D)") &&
              cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize");
    return ok ? 0 : -1;
}

// ---- element access ------------------------------------------------------------------------------
static bool flat_index(const NDArray* a, const size_t* index, size_t* out) {
    size_t off = 0;
    for (size_t i = 0; i < a->lay->ndim; i++) {
        if (index[i] >= a->lay->shape[i]) {
            set_error("IndexError: index %zu out of bounds for axis %zu with extent %zu", index[i], i, a->lay->shape[i]);
            return false;
        }
        off += index[i] * a->lay->stride[i];
    }
    if (view_offset(a) + off >= a->data->size) {
        set_error("IndexError: element offset outside the base buffer");
        return false;
    }
    *out = off;
    return true;
}

f32 array_get_elem(const NDArray* a, const size_t* index) {
    AL_ENTER;
    size_t off;
    f32 v = __builtin_nanf("");
    if (!valid(a, "array_get_elem") || !flat_index(a, index, &off)) return v;
    if (cuda_ok(cudaMemcpyAsync(&v, a->ptr + off, sizeof(float), cudaMemcpyDeviceToHost, g_rt.stream), "D2H"))
        cuda_ok(cudaStreamSynchronize(g_rt.stream), "cudaStreamSynchronize");
    return v;
}

NDArray* array_set_elem_from_scalar(NDArray* a, f32 value, const size_t* index) {
    AL_ENTER;
    size_t off;
    if (!valid_mut(a, "array_set_elem_from_scalar") || !flat_index(a, index, &off)) return nullptr;
    if (!fill_const(a->ptr + off, 1, value)) return nullptr;
    return a;
}

// ---- views ---------------------------------------------------------------------------------------
static bool check_range(const NDArray* a, const size_t* start, const size_t* end, const size_t* step) {
    for (size_t i = 0; i < a->lay->ndim; i++) {
        if (!(start[i] < end[i])) return set_error("ValueError: start index >= end index."), false;
        if (!(end[i] <= a->lay->shape[i])) return set_error("ValueError: end index out of bounds."), false;
        if (!(step[i] > 0)) return set_error("ValueError: non-positive step size."), false;
    }
    return true;
}

// arraylib.c:326-349
NDArray* array_get_view(const NDArray* src, const size_t* start, const size_t* end, const size_t* step) {
    AL_ENTER;
    if (!valid(src, "array_get_view") || !check_range(src, start, end, step)) return nullptr;
    size_t nd = src->lay->ndim, shape[64], stride[64], off = 0;
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    for (size_t i = 0; i < nd; i++) {
        shape[i] = (end[i] - start[i] + step[i] - 1) / step[i];
        stride[i] = src->lay->stride[i] * step[i];
        off += start[i] * src->lay->stride[i];
    }
    AL_REQUIRE(view_offset(src) + off < src->data->size, "ValueError: offset >= size.");
    return new_view(src, shape, stride, nd, off);
}

// arraylib.c:258-269, plus a bounds check the reference lacks: an out-of-range device access
// would poison the CUDA context instead of merely reading garbage.
NDArray* array_as_strided(const NDArray* src, const Layout* lay) {
    AL_ENTER;
    if (!valid(src, "array_as_strided")) return nullptr;
    AL_REQUIRE(lay && lay->ndim > 0, "ValueError: ndim must be positive.");
    size_t reach = 0;
    for (size_t i = 0; i < lay->ndim; i++) {
        AL_REQUIRE(lay->shape[i] > 0, "ValueError: zero extent in as_strided shape");
        reach += (lay->shape[i] - 1) * lay->stride[i];
    }
    AL_REQUIRE(view_offset(src) + reach < src->data->size,
               "ValueError: as_strided view reaches element %zu of a %zu-element buffer",
               view_offset(src) + reach, src->data->size);
    return new_view(src, lay->shape, lay->stride, lay->ndim, 0);
}

// Returns a contiguous alias of src: a view when src already is contiguous, else a gathered copy
// (arraylib.c:423,433,455,479).
static NDArray* contiguous_alias(const NDArray* src) {
    return is_contiguous(src) ? new_view(src, src->lay->shape, src->lay->stride, src->lay->ndim, 0)
                              : array_deep_copy(src);
}

static void relabel(NDArray* a, const size_t* shape, const size_t* stride, size_t ndim) {
    free_layout(a->lay);
    a->lay = make_layout(shape, stride, ndim);
}

// arraylib.c:421-430. Like the reference, the element count is checked against the BASE buffer.
NDArray* array_reshape(const NDArray* src, const size_t* shape, size_t ndim) {
    AL_ENTER;
    if (!valid(src, "array_reshape")) return nullptr;
    AL_REQUIRE(ndim > 0, "ValueError: non-positive ndim");
    AL_REQUIRE(count_of(shape, ndim) == src->data->size,
               "ValueError: cannot reshape a buffer of %zu elements into %zu", src->data->size, count_of(shape, ndim));
    AL_REQUIRE(count_of(shape, ndim) == count_of(src->lay->shape, src->lay->ndim),
               "ValueError: reshape of a view whose element count differs from its base buffer");
    NDArray* dst = contiguous_alias(src);
    if (!dst) return nullptr;
    relabel(dst, shape, nullptr, ndim);
    return dst;
}

// arraylib.c:432-442. For a contiguous source this is the reference's metadata permutation. For a
// non-contiguous source the reference gathers into a contiguous buffer and then permutes the OLD
// strides over it (wrong values); here the gathered copy's own strides are permuted,
// which yields the mathematically correct transpose.
NDArray* array_transpose(const NDArray* src, const size_t* dims) {
    AL_ENTER;
    if (!valid(src, "array_transpose")) return nullptr;
    size_t nd = src->lay->ndim, shape[64], stride[64];
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    bool seen[64] = {false};
    for (size_t i = 0; i < nd; i++) {
        AL_REQUIRE(dims[i] < nd && !seen[dims[i]], "ValueError: transpose axes are not a permutation");
        seen[dims[i]] = true;
    }
    NDArray* dst = contiguous_alias(src);
    if (!dst) return nullptr;
    for (size_t i = 0; i < nd; i++) {
        shape[i] = dst->lay->shape[dims[i]];
        stride[i] = dst->lay->stride[dims[i]];
    }
    relabel(dst, shape, stride, nd);
    return dst;
}

// arraylib.c:444-475
NDArray* array_move_dim(const NDArray* src, const size_t* from, const size_t* to, size_t ndim) {
    AL_ENTER;
    if (!valid(src, "array_move_dim")) return nullptr;
    size_t nd = src->lay->ndim;
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    size_t to_from[64], shape[64], stride[64];
    bool used[64] = {false};
    for (size_t i = 0; i < nd; i++) to_from[i] = SIZE_MAX;
    for (size_t i = 0; i < ndim; i++) {
        AL_REQUIRE(from[i] < nd && to[i] < nd, "ValueError: out of bounds");
        AL_REQUIRE(!used[from[i]] && to_from[to[i]] == SIZE_MAX, "ValueError: repeated axis in move_dim");
        used[from[i]] = true;
        to_from[to[i]] = from[i];
    }
    size_t j = 0;
    for (size_t i = 0; i < nd; i++) {
        if (to_from[i] != SIZE_MAX) continue;
        while (j < nd && used[j]) j++;
        to_from[i] = j++;
    }
    NDArray* dst = contiguous_alias(src);
    if (!dst) return nullptr;
    for (size_t i = 0; i < nd; i++) {
        shape[i] = dst->lay->shape[to_from[i]];
        stride[i] = dst->lay->stride[to_from[i]];
    }
    relabel(dst, shape, stride, nd);
    return dst;
}

// arraylib.c:477-485
NDArray* array_ravel(const NDArray* src) {
    AL_ENTER;
    if (!valid(src, "array_ravel")) return nullptr;
    size_t total = count_of(src->lay->shape, src->lay->ndim);
    NDArray* dst = contiguous_alias(src);
    if (!dst) return nullptr;
    size_t shape1[1] = {total}, stride1[1] = {1};
    relabel(dst, shape1, stride1, 1);
    return dst;
}

}  // extern "C"
