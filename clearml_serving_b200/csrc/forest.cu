// forest.cu -- GBDT / random-forest prediction on sm_100a (kernel K3 of SURVEY.md 2.2).
//
// Replaces, bit for bit, the CPU predictors the reference calls:
//   * xgboost.Booster.predict      (clearml_serving/serving/preprocess_service.py:478-483)
//       margin = base_score; margin += leaf_t(x) for t in tree order, all fp32; identity link.
//   * sklearn tree ensembles via `self._model.predict(data)` (preprocess_service.py:459-464)
//       acc = init; acc += lr*value_t(x) in fp64 in tree order; out = acc / divisor.
// Bit-identity forces the per-row sum to be SEQUENTIAL in tree order (fp add is not associative),
// so the design splits the work into
//   phase 1 (parallel over (row, tree) pairs): traverse, write the leaf value to an L2-resident
//           scratch matrix leaf[tree][row];
//   phase 2 (ordered): the last CTA of each 32-row tile streams its column block of the leaf
//           matrix through a cp.async shared-memory ring and one warp adds it in tree order
//           (lane = row).  The 4-cycle FADD chain of n_trees adds is the latency floor.
// For very large batches a second kernel keeps one row per thread and walks all trees
// sequentially (no scratch).
//
// HBM layout (packed by clearml_serving_b200/formats.py, validated here):
//   nodes[]  : 8 B each  {f32 value | u32 meta}; meta = feat | default_left<<fb | left<<(fb+1)
//              children are adjacent (right = left+1, renumbered breadth-first at pack time),
//              left == 0 marks a leaf whose `value` is the fp32 leaf value (fp32 mode) or an
//              index into leaf64[] (fp64 mode, value pre-multiplied by the learning rate with
//              the same single rounding the reference performs).
//   tree_offset[] : u32 first node of each tree.
//   The split test is `x < thr` in fp32; sklearn's `x <= thr64` is converted exactly at pack
//   time (thr32 = nextafter(round_down_f32(thr64), +inf)).
// Algorithmic bytes per launch: sum_t (8*internal_t + 4*leaves_t) + rows*(4F + out).
#include "common.cuh"

#include <string.h>
#include <vector>

namespace b2s {

struct ForestBlobHeader {
    char magic[4];  // "B2SF"
    uint32_t version;
    uint32_t n_trees, n_features, n_nodes, feat_bits;
    uint32_t acc_mode;  // 0: fp32 sequential (xgboost), 1: fp64 sequential (sklearn)
    uint32_t n_leaf64;
    uint32_t reserved0, reserved1;
    double base;     // base_score / init
    double divisor;  // 1.0 unless random-forest averaging
};
static_assert(sizeof(ForestBlobHeader) == 56, "blob header layout");

struct ForestParams {
    const uint2 *nodes;
    const uint32_t *tree_offset;
    const double *leaf64;
    int n_trees, n_features, feat_bits, max_depth;
    double base, divisor;
};

__device__ __forceinline__ uint2 ld_node(const uint2 *p) { return __ldg(p); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// One traversal step for U interleaved trees (ILP across trees; one row per thread).
template <int U, typename XF>
__device__ __forceinline__ void traverse(const ForestParams &p, const uint32_t (&base)[U],
                                         uint2 (&cur)[U], XF xget)
{
    const int fb = p.feat_bits;
    const uint32_t fmask = (1u << fb) - 1u;
    for (int d = 0; d < p.max_depth; ++d) {
        bool any = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t meta = cur[u].y;
            const uint32_t left = meta >> (fb + 1);
            if (left) {
                const float x = xget(meta & fmask);
                const float thr = __uint_as_float(cur[u].x);
                const bool dl = (meta >> fb) & 1u;
                const bool go_left = (x != x) ? dl : (x < thr);
                cur[u] = ld_node(p.nodes + base[u] + left + (go_left ? 0u : 1u));
                any = true;
            }
        }
        if (!__any_sync(0xffffffffu, any)) break;
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel B: (row, tree) pairs + ordered sum in the last CTA of each 32-row tile.
// grid = (ceil(T / (WARPS*U)), ceil(rows / 32)); block = WARPS*32; lane = row within the tile.
// ---------------------------------------------------------------------------------------------
template <bool F64, int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32)
forest_pairs_kernel(ForestParams p, const float *__restrict__ X, int64_t n_rows,
                    void *__restrict__ out, void *__restrict__ leaf_scratch, int ldb,
                    unsigned *__restrict__ counters, int x_in_smem)
{
    using acc_t = typename std::conditional<F64, double, float>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned s_ticket;

    float *xs = reinterpret_cast<float *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = p.n_features, T = p.n_trees;
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int rows_here = (int)min((int64_t)32, n_rows - r0);

    if (x_in_smem) {  // x tile, transposed + padded: xs[f*33 + r]
        const float *src = X + r0 * F;
        for (int i = threadIdx.x; i < 32 * F; i += WARPS * 32) {
            const int r = i / F, f = i - r * F;
            xs[f * 33 + r] = (r < rows_here) ? __ldg(src + i) : 0.0f;
        }
        __syncthreads();
    }
    const bool row_ok = lane < rows_here;
    const float *xrow = X + (r0 + (row_ok ? lane : 0)) * F;

    const int t0 = (blockIdx.x * WARPS + warp) * U;
    uint32_t base[U];
    uint2 cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int t = t0 + u;
        base[u] = (t < T) ? __ldg(p.tree_offset + t) : 0u;
        cur[u] = (t < T) ? ld_node(p.nodes + base[u]) : make_uint2(0u, 0u);
    }
    if (x_in_smem) {
        traverse<U>(p, base, cur, [&](uint32_t f) { return xs[f * 33 + lane]; });
    } else {
        traverse<U>(p, base, cur, [&](uint32_t f) { return __ldg(xrow + f); });
    }
    acc_t *leaf = reinterpret_cast<acc_t *>(leaf_scratch);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int t = t0 + u;
        if (t < T && row_ok) {
            acc_t v;
            if (F64) v = (acc_t)__ldg(p.leaf64 + cur[u].x);
            else v = (acc_t)__uint_as_float(cur[u].x);
            leaf[(int64_t)t * ldb + r0 + lane] = v;
        }
    }

    // ---- hand-off: last CTA of this row tile performs the ordered sum -------------------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&counters[blockIdx.y], 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();

    constexpr int CH = F64 ? 64 : 128;   // trees per ring stage (16 KiB)
    constexpr int NST = 3;
    constexpr int PPR = 32 * (int)sizeof(acc_t) / 16;  // 16-byte pieces per tree row
    acc_t *ring = reinterpret_cast<acc_t *>(smem);
    const acc_t *L = reinterpret_cast<const acc_t *>(leaf_scratch) + r0;
    const int nchunks = (T + CH - 1) / CH;

    auto issue = [&](int c) {
        if (c < nchunks) {
            const int tcount = min(CH, T - c * CH);
            unsigned char *dst = reinterpret_cast<unsigned char *>(ring + (size_t)(c % NST) * CH * 32);
            for (int i = threadIdx.x; i < tcount * PPR; i += WARPS * 32) {
                const int t = i / PPR, piece = i - t * PPR;
                cp_async16(dst + (size_t)t * 32 * sizeof(acc_t) + piece * 16,
                           reinterpret_cast<const unsigned char *>(L + (int64_t)(c * CH + t) * ldb) + piece * 16);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);

    acc_t acc = F64 ? (acc_t)p.base : (acc_t)(float)p.base;
    for (int c = 0; c < nchunks; ++c) {
        issue(c + NST - 1);
        cp_async_wait<NST - 1>();
        __syncthreads();
        if (warp == 0) {
            const int tcount = min(CH, T - c * CH);
            const acc_t *buf = ring + (size_t)(c % NST) * CH * 32 + lane;
            int t = 0;
            for (; t + 16 <= tcount; t += 16) {
                acc_t v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = buf[(t + k) * 32];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc = acc + v[k];
            }
            for (; t < tcount; ++t) acc = acc + buf[t * 32];
        }
        __syncthreads();
    }
    if (warp == 0 && row_ok) {
        if (F64) reinterpret_cast<double *>(out)[r0 + lane] = (double)acc / p.divisor;
        else reinterpret_cast<float *>(out)[r0 + lane] = (float)acc;
    }
    if (threadIdx.x == 0) counters[blockIdx.y] = 0u;  // re-arm for the next launch on this stream
}

// ---------------------------------------------------------------------------------------------
// Kernel A: one row per thread, all trees in order (large batches; no scratch).
// ---------------------------------------------------------------------------------------------
template <bool F64, int BLOCK, int U>
__global__ void __launch_bounds__(BLOCK)
forest_rows_kernel(ForestParams p, const float *__restrict__ X, int64_t n_rows,
                   void *__restrict__ out, int x_in_smem)
{
    using acc_t = typename std::conditional<F64, double, float>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    float *xs = reinterpret_cast<float *>(smem);
    const int F = p.n_features, T = p.n_trees;
    const int64_t r0 = (int64_t)blockIdx.x * BLOCK;
    const int rows_here = (int)min((int64_t)BLOCK, n_rows - r0);
    if (x_in_smem) {  // xs[f*BLOCK + r]: conflict-free reads (bank = thread)
        const float *src = X + r0 * F;
        for (int i = threadIdx.x; i < BLOCK * F; i += BLOCK) {
            const int r = i / F, f = i - r * F;
            xs[f * BLOCK + r] = (r < rows_here) ? __ldg(src + i) : 0.0f;
        }
        __syncthreads();
    }
    const bool row_ok = (int)threadIdx.x < rows_here;
    const float *xrow = X + (r0 + (row_ok ? threadIdx.x : 0)) * F;
    acc_t acc = F64 ? (acc_t)p.base : (acc_t)(float)p.base;
    for (int t0 = 0; t0 < T; t0 += U) {
        uint32_t base[U];
        uint2 cur[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u;
            base[u] = (t < T) ? __ldg(p.tree_offset + t) : 0u;
            cur[u] = (t < T) ? ld_node(p.nodes + base[u]) : make_uint2(0u, 0u);
        }
        if (x_in_smem) {
            traverse<U>(p, base, cur, [&](uint32_t f) { return xs[f * BLOCK + threadIdx.x]; });
        } else {
            traverse<U>(p, base, cur, [&](uint32_t f) { return __ldg(xrow + f); });
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0 + u < T) {
                if (F64) acc = acc + (acc_t)__ldg(p.leaf64 + cur[u].x);
                else acc = acc + (acc_t)__uint_as_float(cur[u].x);
            }
        }
    }
    if (row_ok) {
        if (F64) reinterpret_cast<double *>(out)[r0 + threadIdx.x] = (double)acc / p.divisor;
        else reinterpret_cast<float *>(out)[r0 + threadIdx.x] = (float)acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kPairsWarps = 4;
constexpr int kPairsU = 2;
constexpr int kRowsBlock = 128;
constexpr int kRowsU = 4;
constexpr int64_t kPairsMaxRows = 8192;   // above this the rows kernel is used
constexpr int kRingBytes = 3 * 16384;     // NST * 16 KiB
constexpr size_t kCounterBytes = 1024;    // kPairsMaxRows/32 = 256 tile counters

struct ForestModel : Model {
    ForestParams p{};
    void *d_blob = nullptr;
    bool f64 = false;
    int max_smem_optin = 0;

    ~ForestModel() override
    {
        if (d_blob) { cudaSetDevice(device); cudaFree(d_blob); }
    }

    size_t scratch_bytes(int64_t max_rows, int64_t) const override
    {
        const int64_t rows = max_rows < kPairsMaxRows ? max_rows : kPairsMaxRows;
        const size_t leaf = (size_t)p.n_trees * (size_t)round_up(rows, 32) * (f64 ? 8 : 4);
        return kCounterBytes + (size_t)round_up((int64_t)leaf, 256);
    }

    int launch(cudaStream_t st, int64_t n_rows, const void *const *d_in, void *const *d_out,
               const int64_t *, void *scratch, size_t scratch_sz) override
    {
        if (n_rows <= 0) return 0;
        const float *X = static_cast<const float *>(d_in[0]);
        void *out = d_out[0];
        const int F = p.n_features;
        if (n_rows <= kPairsMaxRows) {
            const int ldb = (int)round_up(n_rows, 32);
            const size_t leaf = (size_t)p.n_trees * ldb * (f64 ? 8 : 4);
            if (kCounterBytes + leaf > scratch_sz)
                return fail(B2S_ERR_INVALID, "forest: batch of %lld rows exceeds the stream's max_rows scratch",
                            (long long)n_rows);
            // scratch = [tile counters (zeroed at stream creation, re-armed by the kernel) | leaf matrix]
            unsigned *counters = static_cast<unsigned *>(scratch);
            void *leaf_scratch = static_cast<unsigned char *>(scratch) + kCounterBytes;
            const size_t xs_bytes = (size_t)F * 33 * sizeof(float);
            const int x_in_smem = xs_bytes <= (size_t)(max_smem_optin - 1024) ? 1 : 0;
            size_t smem = kRingBytes;
            if (x_in_smem && xs_bytes > smem) smem = xs_bytes;
            dim3 grid((p.n_trees + kPairsWarps * kPairsU - 1) / (kPairsWarps * kPairsU),
                      (unsigned)((n_rows + 31) / 32));
            if (f64) {
                forest_pairs_kernel<true, kPairsWarps, kPairsU><<<grid, kPairsWarps * 32, smem, st>>>(
                    p, X, n_rows, out, leaf_scratch, ldb, counters, x_in_smem);
            } else {
                forest_pairs_kernel<false, kPairsWarps, kPairsU><<<grid, kPairsWarps * 32, smem, st>>>(
                    p, X, n_rows, out, leaf_scratch, ldb, counters, x_in_smem);
            }
        } else {
            const size_t xs_bytes = (size_t)F * kRowsBlock * sizeof(float);
            const int x_in_smem = xs_bytes <= (size_t)(max_smem_optin - 1024) ? 1 : 0;
            const size_t smem = x_in_smem ? xs_bytes : 0;
            const unsigned grid = (unsigned)((n_rows + kRowsBlock - 1) / kRowsBlock);
            if (f64) {
                forest_rows_kernel<true, kRowsBlock, kRowsU><<<grid, kRowsBlock, smem, st>>>(p, X, n_rows, out, x_in_smem);
            } else {
                forest_rows_kernel<false, kRowsBlock, kRowsU><<<grid, kRowsBlock, smem, st>>>(p, X, n_rows, out, x_in_smem);
            }
        }
        count_launch();
        B2S_CUDA(cudaGetLastError());
        return 0;
    }
};

}  // namespace

int forest_model_create(int device, const void *blob, size_t bytes, Model **out)
{
    if (bytes < sizeof(ForestBlobHeader)) return fail(B2S_ERR_INVALID, "forest blob too small");
    ForestBlobHeader h;
    memcpy(&h, blob, sizeof(h));
    if (memcmp(h.magic, "B2SF", 4) != 0 || h.version != 1)
        return fail(B2S_ERR_INVALID, "forest blob: bad magic/version");
    if (h.n_trees == 0 || h.n_nodes == 0 || h.n_features == 0 || h.feat_bits == 0 || h.feat_bits > 24)
        return fail(B2S_ERR_INVALID, "forest blob: empty model or bad feat_bits");
    if (h.acc_mode > 1) return fail(B2S_ERR_INVALID, "forest blob: bad acc_mode");
    const size_t off_bytes = (size_t)round_up((int64_t)(h.n_trees + 1) * 4, 8);
    const size_t node_bytes = (size_t)h.n_nodes * 8;
    const size_t leaf_bytes = (size_t)h.n_leaf64 * 8;
    const size_t need = sizeof(h) + off_bytes + node_bytes + leaf_bytes;
    if (bytes < need) return fail(B2S_ERR_INVALID, "forest blob truncated: %zu < %zu", bytes, need);
    const unsigned char *base = static_cast<const unsigned char *>(blob);
    const uint32_t *toff = reinterpret_cast<const uint32_t *>(base + sizeof(h));
    const uint64_t *nodes = reinterpret_cast<const uint64_t *>(base + sizeof(h) + off_bytes);

    // validate: children strictly after their parent and inside the tree => traversal terminates
    const int fb = (int)h.feat_bits;
    const uint32_t fmask = (1u << fb) - 1u;
    int max_depth = 0;
    int64_t algo_fixed = 0;
    std::vector<int> depth;
    if (toff[0] != 0 || toff[h.n_trees] != h.n_nodes) return fail(B2S_ERR_INVALID, "forest blob: bad tree offsets");
    for (uint32_t t = 0; t < h.n_trees; ++t) {
        const uint32_t s = toff[t], e = toff[t + 1];
        if (e <= s || e > h.n_nodes) return fail(B2S_ERR_INVALID, "forest blob: bad offsets at tree %u", t);
        const uint32_t n = e - s;
        depth.assign(n, 0);
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t val = (uint32_t)(nodes[s + i] & 0xffffffffu);
            const uint32_t meta = (uint32_t)(nodes[s + i] >> 32);
            const uint32_t left = meta >> (fb + 1);
            if (left) {
                if (left <= i || left + 1 >= n)
                    return fail(B2S_ERR_INVALID, "forest blob: bad child index in tree %u node %u", t, i);
                if ((meta & fmask) >= h.n_features)
                    return fail(B2S_ERR_INVALID, "forest blob: feature index out of range in tree %u", t);
                depth[left] = depth[left + 1] = depth[i] + 1;
                if (depth[i] + 1 > max_depth) max_depth = depth[i] + 1;
                algo_fixed += 8;
            } else {
                if (h.acc_mode == 1 && val >= h.n_leaf64)
                    return fail(B2S_ERR_INVALID, "forest blob: leaf index out of range in tree %u", t);
                algo_fixed += h.acc_mode == 1 ? 8 : 4;
            }
        }
    }

    ForestModel *m = new ForestModel();
    m->device = device;
    m->f64 = h.acc_mode == 1;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, need - sizeof(h));
    if (e != cudaSuccess) { delete m; return fail_cuda(e, "cudaMalloc(forest)"); }
    e = cudaMemcpy(m->d_blob, base + sizeof(h), need - sizeof(h), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete m; return fail_cuda(e, "cudaMemcpy(forest)"); }
    unsigned char *d = static_cast<unsigned char *>(m->d_blob);
    m->p.tree_offset = reinterpret_cast<const uint32_t *>(d);
    m->p.nodes = reinterpret_cast<const uint2 *>(d + off_bytes);
    m->p.leaf64 = reinterpret_cast<const double *>(d + off_bytes + node_bytes);
    m->p.n_trees = (int)h.n_trees;
    m->p.n_features = (int)h.n_features;
    m->p.feat_bits = fb;
    m->p.max_depth = max_depth;
    m->p.base = h.base;
    m->p.divisor = h.divisor;

    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    m->max_smem_optin = optin;
    const int want = optin < 200 * 1024 ? optin : 200 * 1024;
    cudaFuncSetAttribute(forest_pairs_kernel<false, kPairsWarps, kPairsU>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_pairs_kernel<true, kPairsWarps, kPairsU>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_rows_kernel<false, kRowsBlock, kRowsU>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_rows_kernel<true, kRowsBlock, kRowsU>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    m->max_smem_optin = want;

    b2s_model_info &info = m->info;
    info.kind = B2S_MODEL_FOREST;
    info.n_inputs = 1;
    info.n_outputs = 1;
    info.in_dtype[0] = B2S_F32;
    info.out_dtype[0] = m->f64 ? B2S_F64 : B2S_F32;
    info.in_row_elems[0] = h.n_features;
    info.out_row_elems[0] = 1;
    info.weight_bytes = (int64_t)(need - sizeof(h));
    info.algo_bytes_fixed = algo_fixed;
    info.algo_bytes_per_row = (int64_t)h.n_features * 4 + (m->f64 ? 8 : 4);
    *out = m;
    return 0;
}

}  // namespace b2s
