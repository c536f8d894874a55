// common.cuh -- shared helpers of libb200serve.so (error plumbing, launch accounting, model ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <string>

#include "../../include/b200serve.h"

namespace b2s {

// ---- error plumbing: thread-local message behind b2s_last_error() -----------------------------
std::string &tls_error();
int fail(int code, const char *fmt, ...);
int fail_cuda(cudaError_t e, const char *what);

#define B2S_CUDA(call)                                              \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::b2s::fail_cuda(_e, #call);  \
    } while (0)

#define B2S_TRY(call)             \
    do {                          \
        int _s = (call);          \
        if (_s != 0) return _s;   \
    } while (0)

// every kernel launch of the library goes through this counter (bench.py "gpu_launches")
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

inline size_t dtype_size(int dt)
{
    switch (dt) {
    case B2S_F32: case B2S_I32: case B2S_U32: return 4;
    case B2S_F64: case B2S_I64: case B2S_U64: return 8;
    case B2S_U8: case B2S_I8: case B2S_BOOL: return 1;
    case B2S_F16: return 2;
    default: return 0;
    }
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// what the host knows about a launch (ragged models need the offsets on the host as well)
struct LaunchInfo {
    const int64_t *h_row_offsets = nullptr;  // host copy of d_row_offsets (n_rows + 1) or null
    int64_t max_rows = 0;                    // capacity of the stream the launch runs on
    int64_t max_row_elems = 0;
};

// fused GEMM epilogue description (gemm.cu)
struct GemmEpilogue {
    const void *bias;      // [N] fp32 or null
    const void *residual;  // [M, ldc] fp32 when out_f32 else 16-bit, or null
    void *C;
    int ldc;               // elements
    int act;               // 0 none, 1 GELU(erf), 2 ReLU, 3 tanh
    int out_f32;           // 1: C (and residual) fp32, 0: 16-bit like A/B
    int is_bf16;
    int act_after;         // 1: activation applied AFTER the residual add (ResNet), 0: before (BERT)
    int res_prefetch = 0;  // set by the launchers: epilogue warps request the next tile's residual slab into L2
    int tma_store = 0;     // set by the launchers: output tiles leave through TMA stores (the kernel's C tensor map is valid)
};

// Implicit-GEMM convolution geometry (gemm.cu): when `taps` > 0 the A operand of the GEMM is not a matrix in
// memory but the im2col view of an NHWC activation tensor, fetched tile by tile with im2col-mode TMA.
// GEMM row m = output pixel (n, p, q) in NHW order; GEMM k = (r * KS + s) * Cin + c.
struct ConvGeom {
    int taps = 0;      // KS * KS, 0 = plain GEMM
    int KS = 1;
    int cblocks = 1;   // Cin / 64: k-blocks per filter tap
    int OH = 1, OW = 1;
    int stride = 1, pad = 0;
    // Stem form (s2d > 0): the 7x7 stride-2 convolution over 3 channels rewritten as a 4x4 stride-1 convolution over
    // the 2x2 space-to-depth image z[n, Hz, Wz, 16]; k-block a = filter row a: the 4 taps x 16 channels of one output
    // pixel are 128 contiguous bytes of z, so the A tile is a plain (overlapping-stride) tiled TMA box.  An M tile is
    // `tile_rows` = rows_per_tile * OW output pixels (whole output rows), not 128.
    int s2d = 0;
    int tile_rows = 128;
    int rows_per_tile = 1;
};

// ---- internal model interface: each model kind implements launch() on a stream -----------------
struct Model {
    int device = 0;
    b2s_model_info info{};
    virtual ~Model() {}
    // bytes of per-stream device scratch needed for batches of up to max_rows rows
    virtual size_t scratch_bytes(int64_t max_rows, int64_t max_row_elems) const = 0;
    // enqueue the model's kernels; d_in/d_out are device pointers; `scratch` is zero-initialised
    // once at stream creation and must be left zeroed where the kernels rely on it
    virtual int launch(cudaStream_t st, int64_t n_rows, const void *const *d_in,
                       void *const *d_out, const int64_t *d_row_offsets, void *scratch,
                       size_t scratch_bytes, const LaunchInfo &li) = 0;
    // a stream that used `scratch` is going away: drop anything cached against it
    virtual void on_stream_destroy(void *) {}
    // developer aid: copy 64 int64 of kernel phase stamps (model-specific meaning)
    virtual int debug_read(long long *) { return fail(B2S_ERR_INVALID, "no debug data for this model kind"); }
};

int forest_model_create(int device, const void *blob, size_t bytes, Model **out);
int linear_model_create(int device, const void *blob, size_t bytes, Model **out);
int graph_model_create(int device, const void *blob, size_t bytes, Model **out);

}  // namespace b2s
