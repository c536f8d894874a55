// api.cu -- the C ABI of libb200serve.so (include/b200serve.h): device lifetime, the shared pinned
// staging arena, per-endpoint streams with staging slots, and the hot path
//     collate (host gather into a pinned slot) -> H2D -> model kernels -> D2H -> scatter.
// Reference interfaces replaced: TritonPreprocessRequest.process marshalling
// (clearml_serving/serving/preprocess_service.py:385-446) and tritonserver's batch execution
// (configured by clearml_serving/engines/triton/triton_helper.py:291-409).
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b2s {

std::atomic<uint64_t> g_launch_count{0};

std::string &tls_error()
{
    static thread_local std::string e;
    return e;
}

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    tls_error() = buf;
    return code;
}

int fail_cuda(cudaError_t e, const char *what)
{
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        // the literal below is what clearml_serving/serving/main.py:117 matches to restart the worker
        return fail(B2S_ERR_OOM, "CUDA out of memory. %s failed (%s)", what, cudaGetErrorString(e));
    }
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
        return fail(B2S_ERR_NOT_INITIALISED, "no usable CUDA device: %s (%s)", cudaGetErrorString(e), what);
    return fail(B2S_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

namespace {

constexpr int kMaxDevices = 16;
constexpr int kMaxIO = 4;
constexpr size_t kZeroCopyOutMax = 64 * 1024;
constexpr size_t kZeroCopyInMax = 16 * 1024;

// ---- shared pinned arena: first-fit free list with coalescing; grows by whole segments on demand
struct Arena {
    std::vector<std::pair<unsigned char *, size_t>> segments;
    std::map<uintptr_t, size_t> free_blocks;  // address -> length
    size_t default_segment = (size_t)64 << 20;
    std::mutex mu;

    int add_segment(size_t bytes)
    {
        void *p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc(pinned arena)");
        segments.emplace_back(static_cast<unsigned char *>(p), bytes);
        free_blocks[reinterpret_cast<uintptr_t>(p)] = bytes;
        return 0;
    }
    int init(size_t bytes)
    {
        std::lock_guard<std::mutex> g(mu);
        if (!segments.empty()) return 0;
        default_segment = bytes;
        return add_segment(bytes);
    }
    void destroy()
    {
        std::lock_guard<std::mutex> g(mu);
        for (auto &sg : segments) cudaFreeHost(sg.first);
        segments.clear();
        free_blocks.clear();
    }
    unsigned char *take(size_t bytes)
    {
        for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it) {
            if (it->second >= bytes) {
                const uintptr_t addr = it->first;
                const size_t len = it->second;
                free_blocks.erase(it);
                if (len > bytes) free_blocks[addr + bytes] = len - bytes;
                return reinterpret_cast<unsigned char *>(addr);
            }
        }
        return nullptr;
    }
    unsigned char *alloc(size_t bytes)
    {
        bytes = (size_t)round_up((int64_t)(bytes ? bytes : 1), 256);
        std::lock_guard<std::mutex> g(mu);
        unsigned char *p = take(bytes);
        if (!p) {  // grow: endpoints with large inputs (images) need more than the initial segment
            const size_t seg = bytes > default_segment ? bytes : default_segment;
            if (add_segment(seg) != 0) return nullptr;
            p = take(bytes);
        }
        return p;
    }
    void release(unsigned char *p, size_t bytes)
    {
        if (!p) return;
        bytes = (size_t)round_up((int64_t)(bytes ? bytes : 1), 256);
        std::lock_guard<std::mutex> g(mu);
        uintptr_t addr = reinterpret_cast<uintptr_t>(p);
        // coalesce only inside one segment
        uintptr_t seg_lo = 0, seg_hi = 0;
        for (auto &sg : segments) {
            const uintptr_t lo = reinterpret_cast<uintptr_t>(sg.first);
            if (addr >= lo && addr < lo + sg.second) { seg_lo = lo; seg_hi = lo + sg.second; }
        }
        auto next = free_blocks.lower_bound(addr);
        if (next != free_blocks.begin()) {
            auto prev = std::prev(next);
            if (prev->first + prev->second == addr && prev->first >= seg_lo) {
                addr = prev->first;
                bytes += prev->second;
                free_blocks.erase(prev);
            }
        }
        if (next != free_blocks.end() && addr + bytes == next->first && next->first < seg_hi) {
            bytes += next->second;
            free_blocks.erase(next);
        }
        free_blocks[addr] = bytes;
    }
};

enum SlotState { SLOT_FREE = 0, SLOT_ACQUIRED = 1, SLOT_INFLIGHT = 2, SLOT_DONE = 3 };

// ---- collate workers --------------------------------------------------------------------------------------------
// Collating a batch of large requests (ResNet: 128 images x 602 KB of fp32 pixels = 77 MB) into the pinned slot with one
// thread runs at one core's memcpy rate (~10 GB/s, 7.7 ms: 2.5x the model's GPU time).  b2s_infer_batch therefore
// hands batches above kParallelGatherMin to a small pool (the caller works too) in pieces of kGatherPiece bytes.
constexpr size_t kParallelGatherMin = 2u << 20, kGatherPiece = 512u << 10;
struct CopyJob { unsigned char *dst; const unsigned char *src; size_t bytes; };

class GatherPool {
public:
    static GatherPool &get()
    {
        static GatherPool *p = new GatherPool();   // never destroyed: the workers may outlive static destructors
        return *p;
    }
    // copies every job; returns when all of them are done
    void run(const std::vector<CopyJob> &jobs)
    {
        struct Batch { std::atomic<size_t> next{0}; std::atomic<size_t> done{0}; std::vector<CopyJob> pieces; };
        auto b = std::make_shared<Batch>();
        for (const CopyJob &j : jobs)
            for (size_t off = 0; off < j.bytes; off += kGatherPiece)
                b->pieces.push_back(CopyJob{j.dst + off, j.src + off, j.bytes - off < kGatherPiece ? j.bytes - off : kGatherPiece});
        const size_t n = b->pieces.size();
        if (n == 0) return;
        auto work = [b, n]() {
            for (;;) {
                const size_t i = b->next.fetch_add(1);
                if (i >= n) return;
                memcpy(b->pieces[i].dst, b->pieces[i].src, b->pieces[i].bytes);
                b->done.fetch_add(1);
            }
        };
        const size_t helpers = n - 1 < workers_.size() ? n - 1 : workers_.size();
        {
            std::lock_guard<std::mutex> l(mu_);
            for (size_t k = 0; k < helpers; ++k) queue_.push_back(work);
        }
        if (helpers == 1) cv_.notify_one(); else if (helpers > 1) cv_.notify_all();
        work();
        while (b->done.load() < n) std::this_thread::yield();
    }

private:
    GatherPool()
    {
        unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw >= 16 ? 7 : (hw >= 4 ? hw / 2 - 1 : 0);
        if (const char *e = getenv("B2S_GATHER_THREADS")) n = (unsigned)atoi(e) > 0 ? (unsigned)atoi(e) - 1 : 0;
        for (unsigned k = 0; k < n; ++k) {
            workers_.emplace_back([this]() {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> l(mu_);
                        cv_.wait(l, [this]() { return !queue_.empty(); });
                        f = std::move(queue_.front());
                        queue_.pop_front();
                    }
                    f();
                }
            });
            workers_.back().detach();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> queue_;
    std::vector<std::thread> workers_;
};

struct Slot {
    int state = SLOT_FREE;
    uint32_t gen = 0;
    cudaEvent_t ev = nullptr;
    unsigned char *h_in[kMaxIO] = {nullptr, nullptr, nullptr, nullptr};
    unsigned char *h_out[kMaxIO] = {nullptr, nullptr, nullptr, nullptr};
    int64_t *h_row_offsets = nullptr;
    // each slot is an independent lane on the device: its own CUDA stream, staging buffers and model
    // scratch, so consecutive batches of one endpoint overlap (copy / kernel / copy-back) on the GPU
    cudaStream_t st = nullptr;
    void *d_in[kMaxIO] = {nullptr, nullptr, nullptr, nullptr};
    void *d_out[kMaxIO] = {nullptr, nullptr, nullptr, nullptr};
    int64_t *d_row_offsets = nullptr;
    void *scratch = nullptr;
    // leading bytes of input i already on their way to the device (b2s_infer_batch copies a large batch group by
    // group while it is still collating the rest); submit_slot copies what is left
    size_t h2d_issued[kMaxIO] = {0, 0, 0, 0};
    // scatter plan of b2s_infer_batch
    bool scatter = false;
    int64_t n_rows = 0;
    std::vector<void *> out_ptr;       // [n_req * n_outputs]
    std::vector<int64_t> out_row0;     // first batch row of each request
    std::vector<int64_t> out_rows;     // rows of each request
};

struct Stream {
    int device = 0;
    Model *model = nullptr;
    cudaStream_t st = nullptr;
    int64_t max_rows = 0, max_row_elems = 0;
    size_t in_bytes[kMaxIO] = {0, 0, 0, 0}, out_bytes[kMaxIO] = {0, 0, 0, 0};
    size_t in_row_bytes[kMaxIO] = {0, 0, 0, 0}, out_row_bytes[kMaxIO] = {0, 0, 0, 0};
    // `st` / `scratch` alias slot 0's (the lane b2s_infer_device, timers and the L2 flush run on)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    bool zero_copy_out = false;
    bool zero_copy_in = false;   // small fixed-width inputs: kernels read the mapped pinned slot directly
    std::vector<Slot> slots;
    int next = 0;
    std::mutex mu;
};

struct Timer {
    Stream *stream = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
};

struct Global {
    std::mutex mu;
    bool device_inited[kMaxDevices] = {false};
    Arena arena;
    std::vector<Model *> models;
    std::vector<Stream *> streams;
    std::vector<Timer *> timers;
    void *flush_buf[kMaxDevices] = {nullptr};
};
Global &G()
{
    static Global g;
    return g;
}
constexpr size_t kFlushBytes = 256u << 20;  // > 126 MB L2

Model *get_model(b2s_model_t h)
{
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    if (h == 0 || h > g.models.size()) return nullptr;
    return g.models[h - 1];
}
Stream *get_stream(b2s_stream_t h)
{
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    if (h == 0 || h > g.streams.size()) return nullptr;
    return g.streams[h - 1];
}
Timer *get_timer(b2s_timer_t h)
{
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    if (h == 0 || h > g.timers.size()) return nullptr;
    return g.timers[h - 1];
}
bool inited(int device)
{
    return device >= 0 && device < kMaxDevices && G().device_inited[device];
}

void free_stream(Stream *s)
{
    Global &g = G();
    cudaSetDevice(s->device);
    for (Slot &sl : s->slots)
        if (sl.st) cudaStreamSynchronize(sl.st);
    for (Slot &sl : s->slots) {
        if (sl.ev) cudaEventDestroy(sl.ev);
        for (int i = 0; i < kMaxIO; ++i) {
            g.arena.release(sl.h_in[i], s->in_bytes[i]);
            g.arena.release(sl.h_out[i], s->out_bytes[i]);
            if (sl.d_in[i]) cudaFree(sl.d_in[i]);
            if (sl.d_out[i]) cudaFree(sl.d_out[i]);
        }
        g.arena.release(reinterpret_cast<unsigned char *>(sl.h_row_offsets), (size_t)(s->max_rows + 1) * 8);
        if (sl.d_row_offsets) cudaFree(sl.d_row_offsets);
        if (sl.scratch) {
            if (s->model) s->model->on_stream_destroy(sl.scratch);
            cudaFree(sl.scratch);
        }
        if (sl.st) cudaStreamDestroy(sl.st);
    }
    delete s;
}

// Enqueue H2D -> kernels -> D2H -> event for a filled slot.
int submit_slot(Model *m, Stream *s, int slot_idx, int64_t n_rows, const int64_t *row_offsets)
{
    Slot &sl = s->slots[slot_idx];
    const b2s_model_info &info = m->info;
    B2S_CUDA(cudaSetDevice(s->device));
    const void *d_in[kMaxIO];
    void *d_out[kMaxIO];
    bool ragged = false;
    for (int i = 0; i < info.n_inputs; ++i) {
        size_t bytes;
        if (info.in_row_elems[i] < 0) {
            if (!row_offsets) return fail(B2S_ERR_INVALID, "variable-length input needs row_offsets");
            bytes = (size_t)row_offsets[n_rows] * dtype_size(info.in_dtype[i]);
            ragged = true;
        } else {
            bytes = (size_t)n_rows * s->in_row_bytes[i];
        }
        if (bytes > s->in_bytes[i]) return fail(B2S_ERR_INVALID, "batch exceeds the stream's staging capacity");
        if (s->zero_copy_in && bytes <= kZeroCopyInMax) {
            d_in[i] = sl.h_in[i];   // UVA: the mapped pinned slot is addressable from the device (one PCIe read, no copy op)
        } else {
            const size_t done = sl.h2d_issued[i] < bytes ? sl.h2d_issued[i] : bytes;
            if (bytes > done)
                B2S_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(sl.d_in[i]) + done, sl.h_in[i] + done, bytes - done,
                                         cudaMemcpyHostToDevice, sl.st));
            d_in[i] = sl.d_in[i];
        }
    }
    if (ragged) {
        if (row_offsets != sl.h_row_offsets) memcpy(sl.h_row_offsets, row_offsets, (size_t)(n_rows + 1) * 8);
        B2S_CUDA(cudaMemcpyAsync(sl.d_row_offsets, sl.h_row_offsets, (size_t)(n_rows + 1) * 8,
                                 cudaMemcpyHostToDevice, sl.st));
    }
    for (int o = 0; o < info.n_outputs; ++o) d_out[o] = s->zero_copy_out ? (void *)sl.h_out[o] : sl.d_out[o];
    LaunchInfo li;
    li.h_row_offsets = ragged ? sl.h_row_offsets : nullptr;
    li.max_rows = s->max_rows;
    li.max_row_elems = s->max_row_elems;
    B2S_TRY(m->launch(sl.st, n_rows, d_in, d_out, ragged ? sl.d_row_offsets : nullptr, sl.scratch, s->scratch_bytes, li));
    if (!s->zero_copy_out) {
        for (int o = 0; o < info.n_outputs; ++o) {
            const size_t bytes = (size_t)n_rows * s->out_row_bytes[o];
            if (bytes) B2S_CUDA(cudaMemcpyAsync(sl.h_out[o], sl.d_out[o], bytes, cudaMemcpyDeviceToHost, sl.st));
        }
    }
    B2S_CUDA(cudaEventRecord(sl.ev, sl.st));
    sl.n_rows = n_rows;
    return 0;
}

int take_slot(Stream *s)
{
    std::lock_guard<std::mutex> l(s->mu);
    const int n = (int)s->slots.size();
    for (int k = 0; k < n; ++k) {
        const int i = (s->next + k) % n;
        if (s->slots[i].state == SLOT_FREE) {
            s->slots[i].state = SLOT_ACQUIRED;
            s->slots[i].gen++;
            for (int k2 = 0; k2 < kMaxIO; ++k2) s->slots[i].h2d_issued[k2] = 0;
            s->next = (i + 1) % n;
            return i;
        }
    }
    return -1;
}

inline b2s_event_t make_event(b2s_stream_t sh, int slot, uint32_t gen)
{
    return ((uint64_t)sh << 32) | ((uint64_t)(slot & 0xffff) << 16) | (uint64_t)(gen & 0xffff);
}

}  // namespace
}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_abi_version(void) { return B2S_ABI_VERSION; }

const char *b2s_last_error(void) { return tls_error().c_str(); }

uint64_t b2s_launch_count(void) { return g_launch_count.load(); }

int b2s_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b2s_init(int device, size_t pinned_arena_bytes)
{
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B2S_ERR_NOT_INITIALISED, "b2s_init: no CUDA device available (%s); libb200serve has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n || device >= kMaxDevices)
        return fail(B2S_ERR_INVALID, "b2s_init: device %d out of range (have %d)", device, n);
    B2S_CUDA(cudaSetDevice(device));
    B2S_CUDA(cudaFree(0));
    B2S_TRY(g.arena.init(pinned_arena_bytes ? pinned_arena_bytes : ((size_t)64 << 20)));
    g.device_inited[device] = true;
    return 0;
}

int b2s_shutdown(void)
{
    Global &g = G();
    std::vector<Stream *> streams;
    std::vector<Model *> models;
    std::vector<Timer *> timers;
    {
        std::lock_guard<std::mutex> l(g.mu);
        streams.swap(g.streams);
        models.swap(g.models);
        timers.swap(g.timers);
    }
    for (Timer *t : timers) {
        if (!t) continue;
        cudaEventDestroy(t->a);
        cudaEventDestroy(t->b);
        delete t;
    }
    for (Stream *s : streams)
        if (s) free_stream(s);
    for (Model *m : models) delete m;
    std::lock_guard<std::mutex> l(g.mu);
    for (int d = 0; d < kMaxDevices; ++d) {
        if (g.flush_buf[d]) {
            cudaSetDevice(d);
            cudaFree(g.flush_buf[d]);
            g.flush_buf[d] = nullptr;
        }
        g.device_inited[d] = false;
    }
    g.arena.destroy();
    return 0;
}

int b2s_model_load(int device, int kind, const void *blob, size_t blob_bytes, const char *cfg_json,
                   b2s_model_t *out_model)
{
    (void)cfg_json;
    if (!out_model || !blob) return fail(B2S_ERR_INVALID, "b2s_model_load: null argument");
    if (!inited(device)) return fail(B2S_ERR_NOT_INITIALISED, "b2s_model_load: device %d not initialised (call b2s_init)", device);
    Model *m = nullptr;
    switch (kind) {
    case B2S_MODEL_FOREST: B2S_TRY(forest_model_create(device, blob, blob_bytes, &m)); break;
    case B2S_MODEL_LINEAR: B2S_TRY(linear_model_create(device, blob, blob_bytes, &m)); break;
    case B2S_MODEL_GRAPH: B2S_TRY(graph_model_create(device, blob, blob_bytes, &m)); break;
    default: return fail(B2S_ERR_INVALID, "b2s_model_load: unknown model kind %d", kind);
    }
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    g.models.push_back(m);
    *out_model = (b2s_model_t)g.models.size();
    return 0;
}

int b2s_model_free(b2s_model_t model)
{
    Global &g = G();
    Model *m = nullptr;
    {
        std::lock_guard<std::mutex> l(g.mu);
        if (model == 0 || model > g.models.size() || !g.models[model - 1])
            return fail(B2S_ERR_INVALID, "b2s_model_free: bad handle");
        for (Stream *s : g.streams)
            if (s && s->model == g.models[model - 1])
                return fail(B2S_ERR_INVALID, "b2s_model_free: a stream still references this model");
        m = g.models[model - 1];
        g.models[model - 1] = nullptr;
    }
    delete m;
    return 0;
}

int b2s_model_get_info(b2s_model_t model, b2s_model_info *out_info)
{
    Model *m = get_model(model);
    if (!m || !out_info) return fail(B2S_ERR_INVALID, "b2s_model_get_info: bad handle");
    *out_info = m->info;
    return 0;
}

int b2s_debug_read(b2s_model_t model, long long *out64)
{
    Model *m = get_model(model);
    if (!m || !out64) return fail(B2S_ERR_INVALID, "b2s_debug_read: bad handle");
    return m->debug_read(out64);
}

int b2s_stream_create(b2s_model_t model, int64_t max_rows, int64_t max_row_elems, int n_slots,
                      b2s_stream_t *out_stream)
{
    Model *m = get_model(model);
    if (!m || !out_stream) return fail(B2S_ERR_INVALID, "b2s_stream_create: bad model handle");
    if (max_rows <= 0) return fail(B2S_ERR_INVALID, "b2s_stream_create: max_rows must be > 0");
    if (n_slots <= 0) n_slots = 4;
    if (n_slots > 64) n_slots = 64;
    Global &g = G();
    const b2s_model_info &info = m->info;
    Stream *s = new Stream();
    s->device = m->device;
    s->model = m;
    s->max_rows = max_rows;
    s->max_row_elems = max_row_elems;
    auto bail = [&](int code) {
        free_stream(s);
        return code;
    };
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return bail(fail_cuda(e, "cudaSetDevice"));
    size_t total_out = 0;
    bool ragged = false;
    for (int i = 0; i < info.n_inputs; ++i) {
        const size_t es = dtype_size(info.in_dtype[i]);
        if (info.in_row_elems[i] < 0) {
            if (max_row_elems <= 0) return bail(fail(B2S_ERR_INVALID, "b2s_stream_create: variable-length model needs max_row_elems"));
            s->in_row_bytes[i] = 0;
            s->in_bytes[i] = (size_t)max_rows * (size_t)max_row_elems * es;
            ragged = true;
        } else {
            s->in_row_bytes[i] = (size_t)info.in_row_elems[i] * es;
            s->in_bytes[i] = (size_t)max_rows * s->in_row_bytes[i];
        }
    }
    for (int o = 0; o < info.n_outputs; ++o) {
        s->out_row_bytes[o] = (size_t)info.out_row_elems[o] * dtype_size(info.out_dtype[o]);
        s->out_bytes[o] = (size_t)max_rows * s->out_row_bytes[o];
        total_out += s->out_bytes[o];
    }
    const char *zc = getenv("B2S_ZEROCOPY_OUT");
    s->zero_copy_out = (total_out <= kZeroCopyOutMax) && !(zc && zc[0] == '0');
    const char *zci = getenv("B2S_ZEROCOPY_IN");
    s->zero_copy_in = !ragged && info.kind != B2S_MODEL_GRAPH && !(zci && zci[0] == '0');
    s->scratch_bytes = m->scratch_bytes(max_rows, max_row_elems);
    s->slots.resize(n_slots);
    for (Slot &sl : s->slots) {
        e = cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking);
        if (e != cudaSuccess) return bail(fail_cuda(e, "cudaStreamCreate"));
        e = cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return bail(fail_cuda(e, "cudaEventCreate"));
        for (int i = 0; i < info.n_inputs; ++i) {
            e = cudaMalloc(&sl.d_in[i], s->in_bytes[i] ? s->in_bytes[i] : 256);
            if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMalloc(slot input)"));
            sl.h_in[i] = g.arena.alloc(s->in_bytes[i]);
            if (!sl.h_in[i]) return bail(fail(B2S_ERR_OOM, "CUDA out of memory. pinned staging arena exhausted (input slot of %zu bytes)", s->in_bytes[i]));
        }
        for (int o = 0; o < info.n_outputs; ++o) {
            e = cudaMalloc(&sl.d_out[o], s->out_bytes[o] ? s->out_bytes[o] : 256);
            if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMalloc(slot output)"));
            sl.h_out[o] = g.arena.alloc(s->out_bytes[o]);
            if (!sl.h_out[o]) return bail(fail(B2S_ERR_OOM, "CUDA out of memory. pinned staging arena exhausted (output slot of %zu bytes)", s->out_bytes[o]));
        }
        sl.h_row_offsets = reinterpret_cast<int64_t *>(g.arena.alloc((size_t)(max_rows + 1) * 8));
        if (!sl.h_row_offsets) return bail(fail(B2S_ERR_OOM, "CUDA out of memory. pinned staging arena exhausted (row offsets)"));
        if (ragged) {
            e = cudaMalloc(reinterpret_cast<void **>(&sl.d_row_offsets), (size_t)(max_rows + 1) * 8);
            if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMalloc(row offsets)"));
        }
        e = cudaMalloc(&sl.scratch, s->scratch_bytes);
        if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMalloc(slot scratch)"));
        e = cudaMemset(sl.scratch, 0, s->scratch_bytes);
        if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMemset(slot scratch)"));
    }
    s->st = s->slots[0].st;
    s->scratch = s->slots[0].scratch;
    std::lock_guard<std::mutex> l(g.mu);
    g.streams.push_back(s);
    *out_stream = (b2s_stream_t)g.streams.size();
    return 0;
}

int b2s_stream_destroy(b2s_stream_t stream)
{
    Global &g = G();
    Stream *s = nullptr;
    {
        std::lock_guard<std::mutex> l(g.mu);
        if (stream == 0 || stream > g.streams.size() || !g.streams[stream - 1])
            return fail(B2S_ERR_INVALID, "b2s_stream_destroy: bad handle");
        s = g.streams[stream - 1];
        g.streams[stream - 1] = nullptr;
        for (Timer *&t : g.timers) {
            if (t && t->stream == s) {
                cudaEventDestroy(t->a);
                cudaEventDestroy(t->b);
                delete t;
                t = nullptr;
            }
        }
    }
    free_stream(s);
    return 0;
}

int b2s_stream_synchronize(b2s_stream_t stream)
{
    Stream *s = get_stream(stream);
    if (!s) return fail(B2S_ERR_INVALID, "b2s_stream_synchronize: bad handle");
    B2S_CUDA(cudaSetDevice(s->device));
    for (Slot &sl : s->slots) B2S_CUDA(cudaStreamSynchronize(sl.st));
    return 0;
}

void *b2s_stream_cuda_handle(b2s_stream_t stream)
{
    Stream *s = get_stream(stream);
    return s ? (void *)s->st : nullptr;
}

int b2s_infer_batch(b2s_model_t model, b2s_stream_t stream, int32_t n_req, const b2s_tensor *in,
                    b2s_tensor *out, b2s_event_t *out_done)
{
    Model *m = get_model(model);
    Stream *s = get_stream(stream);
    if (!m || !s || s->model != m) return fail(B2S_ERR_INVALID, "b2s_infer_batch: bad model/stream handle");
    if (n_req <= 0 || !in || !out || !out_done) return fail(B2S_ERR_INVALID, "b2s_infer_batch: null or empty request list");
    const b2s_model_info &info = m->info;
    bool ragged = false;
    for (int i = 0; i < info.n_inputs; ++i) ragged = ragged || info.in_row_elems[i] < 0;

    // pass 1: validate and count rows (each request carries its own leading batch dim)
    int64_t total_rows = 0, total_elems = 0;
    for (int r = 0; r < n_req; ++r) {
        int64_t rows = -1, row_len = -1;
        for (int i = 0; i < info.n_inputs; ++i) {
            const b2s_tensor &t = in[(size_t)r * info.n_inputs + i];
            if (!t.data && t.ndim > 0) return fail(B2S_ERR_INVALID, "request %d input %d: null data", r, i);
            if (t.dtype != info.in_dtype[i])
                return fail(B2S_ERR_INVALID, "request %d input %d: dtype %d, model expects %d", r, i, t.dtype, info.in_dtype[i]);
            if (t.ndim < 1 || t.ndim > B2S_MAX_DIMS) return fail(B2S_ERR_INVALID, "request %d input %d: bad ndim %d", r, i, t.ndim);
            int64_t elems = 1;
            for (int d = 0; d < t.ndim; ++d) {
                if (t.shape[d] < 0) return fail(B2S_ERR_INVALID, "request %d input %d: negative dim", r, i);
                elems *= t.shape[d];
            }
            int64_t rr;
            if (info.in_row_elems[i] < 0) {  // variable length: [rows, len] (or [len] = one row)
                rr = t.ndim >= 2 ? t.shape[0] : 1;
                const int64_t len = rr > 0 ? elems / rr : 0;
                if (len <= 0) return fail(B2S_ERR_INVALID, "request %d input %d: empty sequence", r, i);
                if (row_len >= 0 && len != row_len) return fail(B2S_ERR_INVALID, "request %d: inputs disagree on sequence length", r);
                row_len = len;
            } else {
                const int64_t re = info.in_row_elems[i];
                if (elems % re != 0)
                    return fail(B2S_ERR_INVALID, "request %d input %d: %lld elements is not a multiple of the model's %lld per row",
                                r, i, (long long)elems, (long long)re);
                rr = elems / re;
            }
            if (rows >= 0 && rr != rows) return fail(B2S_ERR_INVALID, "request %d: inputs disagree on batch rows", r);
            rows = rr;
        }
        total_rows += rows;
        if (ragged) total_elems += rows * row_len;
    }
    if (total_rows > s->max_rows)
        return fail(B2S_ERR_INVALID, "b2s_infer_batch: %lld rows exceed the stream's max_rows %lld",
                    (long long)total_rows, (long long)s->max_rows);
    if (ragged && total_elems > s->max_rows * s->max_row_elems)
        return fail(B2S_ERR_INVALID, "b2s_infer_batch: %lld tokens exceed the stream's capacity %lld",
                    (long long)total_elems, (long long)(s->max_rows * s->max_row_elems));
    const int slot_idx = take_slot(s);
    if (slot_idx < 0) return fail(B2S_ERR_BUSY, "b2s_infer_batch: all %d staging slots are in flight", (int)s->slots.size());
    Slot &sl = s->slots[slot_idx];

    // pass 2: collate -- gather every request's rows into the pinned slot, record the scatter plan
    sl.scatter = true;
    sl.out_ptr.assign((size_t)n_req * info.n_outputs, nullptr);
    sl.out_row0.assign(n_req, 0);
    sl.out_rows.assign(n_req, 0);
    int64_t row = 0, elem = 0;
    if (ragged) sl.h_row_offsets[0] = 0;
    // large fixed-width batches: collate with the worker pool, in a few groups of requests, and start each group's
    // host->device copy while the next group is still being gathered
    size_t fixed_bytes = 0;
    if (!ragged)
        for (int i = 0; i < info.n_inputs; ++i) fixed_bytes += (size_t)total_rows * s->in_row_bytes[i];
    const bool pooled = !ragged && fixed_bytes >= kParallelGatherMin;
    const int n_groups = pooled ? (n_req >= 8 ? 4 : (n_req >= 2 ? 2 : 1)) : 1;
    std::vector<CopyJob> jobs;
    int group_end = pooled ? (n_req + n_groups - 1) / n_groups : n_req;
    auto flush_group = [&](int64_t rows_so_far) -> int {
        if (!pooled) return 0;
        GatherPool::get().run(jobs);
        jobs.clear();
        if (cudaSetDevice(s->device) != cudaSuccess) return fail(B2S_ERR_CUDA, "cudaSetDevice failed");
        for (int i = 0; i < info.n_inputs; ++i) {
            const size_t upto = (size_t)rows_so_far * s->in_row_bytes[i];
            if (upto <= kZeroCopyInMax && s->zero_copy_in) continue;   // submit_slot reads it in place
            if (upto > sl.h2d_issued[i]) {
                B2S_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(sl.d_in[i]) + sl.h2d_issued[i], sl.h_in[i] + sl.h2d_issued[i],
                                         upto - sl.h2d_issued[i], cudaMemcpyHostToDevice, sl.st));
                sl.h2d_issued[i] = upto;
            }
        }
        return 0;
    };
    for (int r = 0; r < n_req; ++r) {
        int64_t rows = 0, row_len = 0;
        for (int i = 0; i < info.n_inputs; ++i) {
            const b2s_tensor &t = in[(size_t)r * info.n_inputs + i];
            int64_t elems = 1;
            for (int d = 0; d < t.ndim; ++d) elems *= t.shape[d];
            if (info.in_row_elems[i] < 0) {
                rows = t.ndim >= 2 ? t.shape[0] : 1;
                row_len = elems / rows;
                memcpy(sl.h_in[i] + (size_t)elem * dtype_size(t.dtype), t.data, (size_t)elems * dtype_size(t.dtype));
            } else {
                rows = elems / info.in_row_elems[i];
                unsigned char *dst = sl.h_in[i] + (size_t)row * s->in_row_bytes[i];
                const size_t nb = (size_t)elems * dtype_size(t.dtype);
                if (pooled) jobs.push_back(CopyJob{dst, static_cast<const unsigned char *>(t.data), nb});
                else memcpy(dst, t.data, nb);
            }
        }
        if (ragged) {
            for (int64_t k = 0; k < rows; ++k) sl.h_row_offsets[row + k + 1] = elem + (k + 1) * row_len;
            elem += rows * row_len;
        }
        sl.out_row0[r] = row;
        sl.out_rows[r] = rows;
        for (int o = 0; o < info.n_outputs; ++o) {
            b2s_tensor &t = out[(size_t)r * info.n_outputs + o];
            if (!t.data && rows > 0) {
                std::lock_guard<std::mutex> l(s->mu);
                sl.state = SLOT_FREE;
                return fail(B2S_ERR_INVALID, "request %d output %d: null destination buffer", r, o);
            }
            sl.out_ptr[(size_t)r * info.n_outputs + o] = t.data;
            t.dtype = info.out_dtype[o];
            if (info.out_row_elems[o] == 1) {
                t.ndim = 1;
                t.shape[0] = rows;
            } else {
                t.ndim = 2;
                t.shape[0] = rows;
                t.shape[1] = info.out_row_elems[o];
            }
        }
        row += rows;
        if (pooled && (r + 1 == group_end || r + 1 == n_req)) {
            const int frc = flush_group(row);
            if (frc != 0) {
                std::lock_guard<std::mutex> l(s->mu);
                sl.state = SLOT_FREE;
                return frc;
            }
            group_end += (n_req + n_groups - 1) / n_groups;
        }
    }
    const int rc = submit_slot(m, s, slot_idx, total_rows, ragged ? sl.h_row_offsets : nullptr);
    std::lock_guard<std::mutex> l(s->mu);
    if (rc != 0) {
        sl.state = SLOT_FREE;
        return rc;
    }
    sl.state = SLOT_INFLIGHT;
    *out_done = make_event(stream, slot_idx, sl.gen);
    return 0;
}

int b2s_slot_acquire(b2s_stream_t stream, int32_t *out_slot, void **in_ptr, void **out_ptr)
{
    Stream *s = get_stream(stream);
    if (!s || !out_slot) return fail(B2S_ERR_INVALID, "b2s_slot_acquire: bad handle");
    const int idx = take_slot(s);
    if (idx < 0) return fail(B2S_ERR_BUSY, "b2s_slot_acquire: all %d staging slots are in flight", (int)s->slots.size());
    Slot &sl = s->slots[idx];
    sl.scatter = false;
    const b2s_model_info &info = s->model->info;
    if (in_ptr) for (int i = 0; i < info.n_inputs; ++i) in_ptr[i] = sl.h_in[i];
    if (out_ptr) for (int o = 0; o < info.n_outputs; ++o) out_ptr[o] = sl.h_out[o];
    *out_slot = idx;
    return 0;
}

int b2s_slot_submit(b2s_model_t model, b2s_stream_t stream, int32_t slot, int64_t n_rows,
                    const int64_t *row_offsets, b2s_event_t *out_done)
{
    Model *m = get_model(model);
    Stream *s = get_stream(stream);
    if (!m || !s || s->model != m || !out_done) return fail(B2S_ERR_INVALID, "b2s_slot_submit: bad handle");
    if (slot < 0 || slot >= (int)s->slots.size()) return fail(B2S_ERR_INVALID, "b2s_slot_submit: bad slot");
    if (n_rows < 0 || n_rows > s->max_rows) return fail(B2S_ERR_INVALID, "b2s_slot_submit: n_rows %lld out of range", (long long)n_rows);
    Slot &sl = s->slots[slot];
    {
        std::lock_guard<std::mutex> l(s->mu);
        if (sl.state != SLOT_ACQUIRED) return fail(B2S_ERR_INVALID, "b2s_slot_submit: slot %d is not acquired", slot);
    }
    const int rc = submit_slot(m, s, slot, n_rows, row_offsets);
    std::lock_guard<std::mutex> l(s->mu);
    if (rc != 0) return rc;
    sl.state = SLOT_INFLIGHT;
    *out_done = make_event(stream, slot, sl.gen);
    return 0;
}

int b2s_slot_collate(b2s_model_t model, b2s_stream_t stream, int32_t slot, int32_t n_req,
                     const void *const *in_ptrs, const int64_t *req_rows, const int64_t *req_row_len,
                     b2s_event_t *out_done)
{
    Model *m = get_model(model);
    Stream *s = get_stream(stream);
    if (!m || !s || s->model != m || !out_done) return fail(B2S_ERR_INVALID, "b2s_slot_collate: bad handle");
    if (slot < 0 || slot >= (int)s->slots.size()) return fail(B2S_ERR_INVALID, "b2s_slot_collate: bad slot");
    if (n_req <= 0 || !in_ptrs || !req_rows) return fail(B2S_ERR_INVALID, "b2s_slot_collate: null or empty request list");
    const b2s_model_info &info = m->info;
    bool ragged = false;
    for (int i = 0; i < info.n_inputs; ++i) ragged = ragged || info.in_row_elems[i] < 0;
    if (ragged && !req_row_len) return fail(B2S_ERR_INVALID, "b2s_slot_collate: variable-length model needs req_row_len");
    Slot &sl = s->slots[slot];
    {
        std::lock_guard<std::mutex> l(s->mu);
        if (sl.state != SLOT_ACQUIRED) return fail(B2S_ERR_INVALID, "b2s_slot_collate: slot %d is not acquired", slot);
    }
    int64_t total_rows = 0, total_elems = 0;
    for (int r = 0; r < n_req; ++r) {
        if (req_rows[r] < 0) return fail(B2S_ERR_INVALID, "b2s_slot_collate: request %d has negative rows", r);
        total_rows += req_rows[r];
        if (ragged) {
            if (req_row_len[r] <= 0 || req_row_len[r] > s->max_row_elems)
                return fail(B2S_ERR_INVALID, "b2s_slot_collate: request %d: sequence length %lld outside (0, %lld]", r,
                            (long long)req_row_len[r], (long long)s->max_row_elems);
            total_elems += req_rows[r] * req_row_len[r];
        }
        for (int i = 0; i < info.n_inputs; ++i)
            if (!in_ptrs[(size_t)r * info.n_inputs + i] && req_rows[r] > 0)
                return fail(B2S_ERR_INVALID, "b2s_slot_collate: request %d input %d: null data", r, i);
    }
    if (total_rows > s->max_rows)
        return fail(B2S_ERR_INVALID, "b2s_slot_collate: %lld rows exceed the stream's max_rows %lld", (long long)total_rows,
                    (long long)s->max_rows);
    if (ragged && total_elems > s->max_rows * s->max_row_elems)
        return fail(B2S_ERR_INVALID, "b2s_slot_collate: %lld tokens exceed the stream's capacity %lld", (long long)total_elems,
                    (long long)(s->max_rows * s->max_row_elems));
    size_t fixed_bytes = 0;
    if (!ragged)
        for (int i = 0; i < info.n_inputs; ++i) fixed_bytes += (size_t)total_rows * s->in_row_bytes[i];
    const bool pooled = !ragged && fixed_bytes >= kParallelGatherMin;
    std::vector<CopyJob> jobs;
    int64_t row = 0, elem = 0;
    if (ragged) sl.h_row_offsets[0] = 0;
    for (int r = 0; r < n_req; ++r) {
        const int64_t rows = req_rows[r];
        for (int i = 0; i < info.n_inputs; ++i) {
            const unsigned char *src = static_cast<const unsigned char *>(in_ptrs[(size_t)r * info.n_inputs + i]);
            if (info.in_row_elems[i] < 0) {
                const size_t es = dtype_size(info.in_dtype[i]);
                memcpy(sl.h_in[i] + (size_t)elem * es, src, (size_t)(rows * req_row_len[r]) * es);
            } else {
                unsigned char *dst = sl.h_in[i] + (size_t)row * s->in_row_bytes[i];
                const size_t nb = (size_t)rows * s->in_row_bytes[i];
                if (pooled) jobs.push_back(CopyJob{dst, src, nb});
                else memcpy(dst, src, nb);
            }
        }
        if (ragged) {
            for (int64_t k = 0; k < rows; ++k) sl.h_row_offsets[row + k + 1] = elem + (k + 1) * req_row_len[r];
            elem += rows * req_row_len[r];
        }
        row += rows;
    }
    if (pooled) GatherPool::get().run(jobs);
    sl.scatter = false;
    const int rc = submit_slot(m, s, slot, total_rows, ragged ? sl.h_row_offsets : nullptr);
    std::lock_guard<std::mutex> l(s->mu);
    if (rc != 0) return rc;
    sl.state = SLOT_INFLIGHT;
    *out_done = make_event(stream, slot, sl.gen);
    return 0;
}

int b2s_slot_release(b2s_stream_t stream, int32_t slot)
{
    Stream *s = get_stream(stream);
    if (!s || slot < 0 || slot >= (int)s->slots.size()) return fail(B2S_ERR_INVALID, "b2s_slot_release: bad handle");
    std::lock_guard<std::mutex> l(s->mu);
    Slot &sl = s->slots[slot];
    if (sl.state == SLOT_INFLIGHT) return fail(B2S_ERR_BUSY, "b2s_slot_release: slot %d still in flight", slot);
    sl.state = SLOT_FREE;
    return 0;
}

static int resolve_event(b2s_event_t ev, Stream **ps, Slot **psl)
{
    Stream *s = get_stream((b2s_stream_t)(ev >> 32));
    const int slot = (int)((ev >> 16) & 0xffff);
    if (!s || slot >= (int)s->slots.size()) return fail(B2S_ERR_INVALID, "bad event handle");
    Slot &sl = s->slots[slot];
    if ((sl.gen & 0xffff) != (uint32_t)(ev & 0xffff)) return fail(B2S_ERR_INVALID, "stale event handle");
    *ps = s;
    *psl = &sl;
    return 0;
}

int b2s_event_query(b2s_event_t ev)
{
    Stream *s;
    Slot *sl;
    B2S_TRY(resolve_event(ev, &s, &sl));
    if (sl->state != SLOT_INFLIGHT) return 0;
    cudaSetDevice(s->device);
    cudaError_t e = cudaEventQuery(sl->ev);
    if (e == cudaSuccess) return 0;
    if (e == cudaErrorNotReady) return B2S_ERR_NOT_READY;
    return fail_cuda(e, "cudaEventQuery");
}

int b2s_event_wait(b2s_event_t ev)
{
    Stream *s;
    Slot *sl;
    B2S_TRY(resolve_event(ev, &s, &sl));
    {
        std::lock_guard<std::mutex> l(s->mu);
        if (sl->state != SLOT_INFLIGHT) return fail(B2S_ERR_INVALID, "b2s_event_wait: batch is not in flight (already waited?)");
    }
    cudaSetDevice(s->device);
    cudaError_t e = cudaEventSynchronize(sl->ev);
    if (e != cudaSuccess) {
        std::lock_guard<std::mutex> l(s->mu);
        sl->state = SLOT_FREE;
        return fail_cuda(e, "cudaEventSynchronize");
    }
    if (sl->scatter) {  // scatter: split the batch output back into the per-request buffers
        const b2s_model_info &info = s->model->info;
        const size_t n_req = sl->out_rows.size();
        for (size_t r = 0; r < n_req; ++r) {
            for (int o = 0; o < info.n_outputs; ++o) {
                const size_t rb = s->out_row_bytes[o];
                if (sl->out_rows[r] > 0)
                    memcpy(sl->out_ptr[r * info.n_outputs + o], sl->h_out[o] + (size_t)sl->out_row0[r] * rb,
                           (size_t)sl->out_rows[r] * rb);
            }
        }
        std::lock_guard<std::mutex> l(s->mu);
        sl->state = SLOT_FREE;
    } else {
        std::lock_guard<std::mutex> l(s->mu);
        sl->state = SLOT_DONE;
    }
    return 0;
}

int b2s_infer_device(b2s_model_t model, b2s_stream_t stream, int64_t n_rows, const void *const *d_in,
                     void *const *d_out, const int64_t *d_row_offsets)
{
    Model *m = get_model(model);
    Stream *s = get_stream(stream);
    if (!m || !s || s->model != m || !d_in || !d_out) return fail(B2S_ERR_INVALID, "b2s_infer_device: bad handle");
    if (n_rows < 0 || n_rows > s->max_rows) return fail(B2S_ERR_INVALID, "b2s_infer_device: n_rows %lld out of range", (long long)n_rows);
    B2S_CUDA(cudaSetDevice(s->device));
    LaunchInfo li;
    li.max_rows = s->max_rows;
    li.max_row_elems = s->max_row_elems;
    std::vector<int64_t> h_off;
    if (d_row_offsets) {  // ragged model: the host needs the offsets too (token count, longest sequence)
        h_off.resize((size_t)n_rows + 1);
        B2S_CUDA(cudaMemcpy(h_off.data(), d_row_offsets, h_off.size() * 8, cudaMemcpyDeviceToHost));
        li.h_row_offsets = h_off.data();
    }
    return m->launch(s->st, n_rows, d_in, d_out, d_row_offsets, s->scratch, s->scratch_bytes, li);
}

int b2s_device_malloc(int device, size_t bytes, void **out_ptr)
{
    if (!inited(device) || !out_ptr) return fail(B2S_ERR_NOT_INITIALISED, "b2s_device_malloc: device not initialised");
    B2S_CUDA(cudaSetDevice(device));
    B2S_CUDA(cudaMalloc(out_ptr, bytes ? bytes : 256));
    return 0;
}

int b2s_device_free(int device, void *ptr)
{
    if (!inited(device)) return fail(B2S_ERR_NOT_INITIALISED, "b2s_device_free: device not initialised");
    B2S_CUDA(cudaSetDevice(device));
    B2S_CUDA(cudaFree(ptr));
    return 0;
}

int b2s_memcpy_h2d(int device, void *dst, const void *src, size_t bytes)
{
    if (!inited(device)) return fail(B2S_ERR_NOT_INITIALISED, "b2s_memcpy_h2d: device not initialised");
    B2S_CUDA(cudaSetDevice(device));
    B2S_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

int b2s_memcpy_d2h(int device, void *dst, const void *src, size_t bytes)
{
    if (!inited(device)) return fail(B2S_ERR_NOT_INITIALISED, "b2s_memcpy_d2h: device not initialised");
    B2S_CUDA(cudaSetDevice(device));
    B2S_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int b2s_flush_l2(int device)
{
    if (!inited(device)) return fail(B2S_ERR_NOT_INITIALISED, "b2s_flush_l2: device not initialised");
    Global &g = G();
    B2S_CUDA(cudaSetDevice(device));
    {
        std::lock_guard<std::mutex> l(g.mu);
        if (!g.flush_buf[device]) B2S_CUDA(cudaMalloc(&g.flush_buf[device], kFlushBytes));
    }
    static std::atomic<int> v{0};
    B2S_CUDA(cudaMemset(g.flush_buf[device], (v++) & 0xff, kFlushBytes));
    B2S_CUDA(cudaDeviceSynchronize());
    return 0;
}

int b2s_stream_flush_l2(b2s_stream_t stream)
{
    Stream *s = get_stream(stream);
    if (!s) return fail(B2S_ERR_INVALID, "b2s_stream_flush_l2: bad handle");
    Global &g = G();
    B2S_CUDA(cudaSetDevice(s->device));
    {
        std::lock_guard<std::mutex> l(g.mu);
        if (!g.flush_buf[s->device]) B2S_CUDA(cudaMalloc(&g.flush_buf[s->device], kFlushBytes));
    }
    static std::atomic<int> v{0};
    B2S_CUDA(cudaMemsetAsync(g.flush_buf[s->device], (v++) & 0xff, kFlushBytes, s->st));
    return 0;
}

int b2s_timer_create(b2s_stream_t stream, b2s_timer_t *out_timer)
{
    Stream *s = get_stream(stream);
    if (!s || !out_timer) return fail(B2S_ERR_INVALID, "b2s_timer_create: bad handle");
    B2S_CUDA(cudaSetDevice(s->device));
    Timer *t = new Timer();
    t->stream = s;
    cudaError_t e = cudaEventCreate(&t->a);
    if (e == cudaSuccess) e = cudaEventCreate(&t->b);
    if (e != cudaSuccess) {
        delete t;
        return fail_cuda(e, "cudaEventCreate(timer)");
    }
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    g.timers.push_back(t);
    *out_timer = (b2s_timer_t)g.timers.size();
    return 0;
}

int b2s_timer_start(b2s_timer_t timer)
{
    Timer *t = get_timer(timer);
    if (!t) return fail(B2S_ERR_INVALID, "b2s_timer_start: bad handle");
    B2S_CUDA(cudaSetDevice(t->stream->device));
    B2S_CUDA(cudaEventRecord(t->a, t->stream->st));
    return 0;
}

int b2s_timer_stop(b2s_timer_t timer)
{
    Timer *t = get_timer(timer);
    if (!t) return fail(B2S_ERR_INVALID, "b2s_timer_stop: bad handle");
    B2S_CUDA(cudaSetDevice(t->stream->device));
    B2S_CUDA(cudaEventRecord(t->b, t->stream->st));
    return 0;
}

int b2s_timer_elapsed_ms(b2s_timer_t timer, float *out_ms)
{
    Timer *t = get_timer(timer);
    if (!t || !out_ms) return fail(B2S_ERR_INVALID, "b2s_timer_elapsed_ms: bad handle");
    B2S_CUDA(cudaSetDevice(t->stream->device));
    B2S_CUDA(cudaEventSynchronize(t->b));
    B2S_CUDA(cudaEventElapsedTime(out_ms, t->a, t->b));
    return 0;
}

int b2s_timer_destroy(b2s_timer_t timer)
{
    Global &g = G();
    std::lock_guard<std::mutex> l(g.mu);
    if (timer == 0 || timer > g.timers.size() || !g.timers[timer - 1]) return fail(B2S_ERR_INVALID, "b2s_timer_destroy: bad handle");
    Timer *t = g.timers[timer - 1];
    cudaEventDestroy(t->a);
    cudaEventDestroy(t->b);
    delete t;
    g.timers[timer - 1] = nullptr;
    return 0;
}

}  // extern "C"
