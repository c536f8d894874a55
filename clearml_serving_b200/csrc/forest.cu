// forest.cu -- GBDT / random-forest prediction on sm_100a (kernel K3 of SURVEY.md 2.2).
//
// Replaces, bit for bit, the CPU predictors the reference calls:
//   * xgboost.Booster.predict      (clearml_serving/serving/preprocess_service.py:478-483)
//       margin = base_score; margin += leaf_t(x) for t in tree order, all fp32; identity link.
//   * sklearn tree ensembles via `self._model.predict(data)` (preprocess_service.py:459-464)
//       acc = init; acc += lr*value_t(x) in fp64 in tree order; out = acc / divisor.
// Bit-identity forces the per-row sum to be SEQUENTIAL in tree order (fp add is not associative),
// so the work is split into
//   phase 1 (parallel over (row, tree) pairs): traverse, produce the leaf value;
//   phase 2 (ordered): one lane per row adds its column of leaf values in tree order.
// Serving-size batches (<= 4096 rows) use one thread-block cluster per 32-row tile: leaf values go
// through distributed shared memory into rank 0's SM, never through L2/HBM (forest_cluster_kernel).
// The 4-cycle FADD chain of n_trees adds is the latency floor (~2 us for 1000 trees).
// For very large batches a second kernel keeps one row per thread and walks all trees
// sequentially (forest_rows_kernel).
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

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace b2s {

struct ForestBlobHeader {
    char magic[4];  // "B2SF"
    uint32_t version;
    uint32_t n_trees, n_features, n_nodes, feat_bits;
    uint32_t acc_mode;  // 0: fp32 sequential (xgboost), 1: fp64 sequential (sklearn)
    uint32_t n_leaf64;
    uint32_t link;   // 0 identity, 1 xgboost's fp32 logistic transform of the margin (binary:logistic / reg:logistic)
    uint32_t reserved1;
    double base;     // base margin / init
    double divisor;  // 1.0 unless random-forest averaging
};
static_assert(sizeof(ForestBlobHeader) == 56, "blob header layout");

struct ForestParams {
    const uint2 *nodes;
    const uint32_t *tree_offset;
    const double *leaf64;
    int n_trees, n_features, feat_bits, max_depth;
    double base, divisor;
    int link;
};

// xgboost's Sigmoid (src/common/math.h): 1 / (expf(min(-x, 88.7f)) + 1 + 1e-16f), fp32.  expf here is CUDA's
// (<= 2 ulp), the CPU library's is libm's: the MARGIN is bit-exact, the probability is within a few ulp.
__device__ __forceinline__ float forest_link_f32(float margin, int link)
{
    if (link == 1) {
        const float x = fminf(-margin, 88.7f);
        return __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(expf(x), 1.0f), 1e-16f));
    }
    return margin;
}

__device__ __forceinline__ uint2 ld_node(const uint2 *p) { return __ldg(p); }

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
// Kernel B (serving batches): one thread-block CLUSTER per 32-row tile.
//   * lane = row of the tile; every warp of every CTA of the cluster traverses U trees at once, so a
//     cluster of C CTAs x WARPS warps covers C*WARPS*U trees per round with all node fetches in flight
//     together (the forest is L2-resident; a depth-d tree costs d+1 dependent L2 round trips);
//   * each warp writes its 32 leaf values (one per row) into the leaf matrix that lives in the
//     shared memory of cluster rank 0 -- a coalesced 128-byte DISTRIBUTED-SHARED-MEMORY store -- so
//     the leaf values never touch L2/HBM and no global fence / atomic hand-off is needed;
//   * after one cluster barrier, warp 0 of rank 0 adds the column of each row in tree order out of
//     its local shared memory (the sequential fp32/fp64 chain bit-identity demands) and writes y.
// grid = C * ceil(rows/32) CTAs, cluster (C,1,1), block = WARPS*32, 1 CTA/SM.
// Forests larger than one chunk (C*WARPS*U*rounds trees) loop over chunks carrying the accumulator.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void *local_smem, uint32_t rank)
{
    uint32_t la = (uint32_t)__cvta_generic_to_shared(local_smem), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(la), "r"(rank));
    return ra;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double v)
{
    asm volatile("st.shared::cluster.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}

template <bool F64, int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32, 1)
forest_cluster_kernel(ForestParams p, const float *__restrict__ X, int64_t n_rows,
                      void *__restrict__ out, int x_in_smem, int rounds, long long *__restrict__ dbg)
{
    // optional phase stamps (B2S_FOREST_TIMING=1): [rank][phase] SM-clock values of thread 0, tile 0
#define B2S_STAMP(k)                                                                    \
    do {                                                                                \
        if (dbg && blockIdx.x < 8 && threadIdx.x == 0) dbg[blockIdx.x * 8 + (k)] = clock64(); \
    } while (0)
    using acc_t = typename std::conditional<F64, double, float>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = p.n_features, T = p.n_trees;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int64_t r0 = (int64_t)(blockIdx.x / C) * 32;
    const int rows_here = (int)min((int64_t)32, n_rows - r0);
    const int per_round = (int)C * WARPS * U;          // trees the cluster covers per round
    const int chunk_trees = per_round * rounds;        // trees held by the leaf matrix at once

    // smem: [leaf matrix: chunk_trees x 32 acc_t (used in rank 0 only)] [x tile: F x 33 floats]
    acc_t *leafbuf = reinterpret_cast<acc_t *>(smem);
    float *xs = reinterpret_cast<float *>(smem + (size_t)chunk_trees * 32 * sizeof(acc_t));
    B2S_STAMP(0);
    if (x_in_smem) {
        const float *src = X + r0 * F;
        for (int i = threadIdx.x; i < 32 * F; i += WARPS * 32) {
            const int r = i / F, f = i - r * F;
            xs[f * 33 + r] = (r < rows_here) ? __ldg(src + i) : 0.0f;
        }
    }
    const bool row_ok = lane < rows_here;
    const float *xrow = X + (r0 + (row_ok ? lane : 0)) * F;
    const uint32_t leaf_remote = map_to_rank(leafbuf, 0);  // rank 0's leaf matrix, cluster address
    __syncthreads();
    B2S_STAMP(1);

    acc_t acc = F64 ? (acc_t)p.base : (acc_t)(float)p.base;
    for (int chunk0 = 0; chunk0 < T; chunk0 += chunk_trees) {
        for (int j = 0; j < rounds; ++j) {
            const int slot0 = ((j * (int)C + (int)rank) * WARPS + warp) * U;  // slot inside the chunk
            const int t0 = chunk0 + slot0;
            if (t0 >= T) break;
            uint32_t base[U];
            uint2 cur[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u;
                base[u] = (t < T) ? __ldg(p.tree_offset + t) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) cur[u] = (t0 + u < T) ? ld_node(p.nodes + base[u]) : make_uint2(0u, 0u);
            if (x_in_smem) {
                traverse<U>(p, base, cur, [&](uint32_t f) { return xs[f * 33 + lane]; });
            } else {
                traverse<U>(p, base, cur, [&](uint32_t f) { return __ldg(xrow + f); });
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (t0 + u < T) {
                    acc_t v;
                    if (F64) v = (acc_t)__ldg(p.leaf64 + cur[u].x);
                    else v = (acc_t)__uint_as_float(cur[u].x);
                    st_cluster(leaf_remote + (uint32_t)(((slot0 + u) * 32 + lane) * sizeof(acc_t)), v);
                }
            }
        }
        B2S_STAMP(2);
        cluster_sync_all();  // every leaf value of this chunk has landed in rank 0's shared memory
        B2S_STAMP(3);
        if (rank == 0 && warp == 0) {
            const int tcount = min(chunk_trees, T - chunk0);
            const acc_t *buf = leafbuf + lane;
            int t = 0;
            for (; t + 16 <= tcount; t += 16) {
                acc_t v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = buf[(t + k) * 32];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc = acc + v[k];
            }
            for (; t < tcount; ++t) acc = acc + buf[t * 32];
            B2S_STAMP(4);
        }
        if (chunk0 + chunk_trees < T) cluster_sync_all();  // leaf matrix is reused by the next chunk
    }
    if (rank == 0 && warp == 0 && row_ok) {
        if (F64) reinterpret_cast<double *>(out)[r0 + lane] = (double)acc / p.divisor;
        else reinterpret_cast<float *>(out)[r0 + lane] = forest_link_f32((float)acc, p.link);
    }
#undef B2S_STAMP
}

// ---------------------------------------------------------------------------------------------
// Kernel C (serving batches, fast path): like kernel B but the forest never pays a dependent L2 round
// trip per tree level.  Measured on B200 (profiles/r01_forest_cluster_phase_timing.txt): a dependent
// L2 access costs ~1000 SM cycles at 1.97 GHz, so 8 of them per tree dominated kernel B.  Here
//   * each CTA of the cluster bulk-copies ITS slice of the forest (trees_per_cta consecutive trees,
//     ~130 KB for 128 depth-6 trees) into shared memory with cp.async.bulk (one L2 latency, then
//     full-rate streaming) and traverses out of shared memory (~40 cycles per level);
//   * a row tile is 16 rows: lanes 0-15 / 16-31 of a warp walk two different trees, which halves the
//     leaf matrix (T x 16) so that it fits next to the tree slice;
//   * leaf values are pushed to rank 0 through distributed shared memory and each rank signals its
//     own mbarrier in rank 0, so the ordered sum starts on rank 0's trees at once and the remaining
//     pushes drain behind the 4-cycle-per-tree add chain instead of in front of it.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSlices = 64;
struct SliceTable {
    uint32_t off[kMaxSlices + 1];  // node offset of the first tree of every trees_per_cta-group (+ end)
};

__device__ __forceinline__ void mbar_init_cta(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_cta(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cta(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster_acq(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// shared::cta -> (remote) shared::cluster bulk copy through the async proxy; completes bytes on an
// mbarrier that lives in the DESTINATION CTA
__device__ __forceinline__ void bulk_s2s_cluster(uint32_t dst_cluster_addr, const void *smem_src, uint32_t bytes,
                                                 uint32_t bar_cluster_addr)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst_cluster_addr), "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes),
                   "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
#define B2S_SPIN_LIMIT (1u << 26)

template <bool F64, int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32, 1)
forest_staged_kernel(ForestParams p, SliceTable st, const float *__restrict__ X, int64_t n_rows,
                     void *__restrict__ out, int slice_cap_bytes, int bulk_piece, long long *__restrict__ dbg)
{
#define B2S_STAMP(k)                                                                    \
    do {                                                                                \
        if (dbg && blockIdx.x < 8 && threadIdx.x == 0) dbg[blockIdx.x * 8 + (k)] = clock64(); \
    } while (0)
    using acc_t = typename std::conditional<F64, double, float>::type;
    constexpr int TPC = WARPS * 2 * U;          // trees per CTA per chunk
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row16 = lane & 15, half = lane >> 4;
    const int F = p.n_features, T = p.n_trees;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int64_t r0 = (int64_t)(blockIdx.x / C) * 16;
    const int rows_here = (int)min((int64_t)16, n_rows - r0);
    const int chunk_trees = (int)C * TPC;

    // smem carve-up
    constexpr int VEC = 16 / (int)sizeof(acc_t);   // values per 128-bit shared-memory load
    const int LD = chunk_trees + VEC;              // leaf-matrix row pitch: +16 B keeps 128-bit row reads conflict-free
    constexpr int SLD = TPC + VEC;                 // pitch of the local staging block
    uint2 *snodes = reinterpret_cast<uint2 *>(smem);
    acc_t *leafbuf = reinterpret_cast<acc_t *>(smem + slice_cap_bytes);                       // [16 rows][LD]: row-major per ROW
    float *xs = reinterpret_cast<float *>(smem + slice_cap_bytes + (size_t)16 * LD * sizeof(acc_t));  // [F][17]
    uint32_t *toff = reinterpret_cast<uint32_t *>(xs + (size_t)F * 17);                       // [TPC + 1]
    uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(toff + TPC + 1) + 7) & ~(uintptr_t)7);
    uint64_t *load_bar = bars;            // tree slice landed (this CTA)
    uint64_t *chunk_bar = bars + 1;       // [C] in rank 0: rank r's leaf values landed
    // this CTA's leaf values [16 rows][SLD], staged locally and pushed to rank 0 with one bulk copy per row
    acc_t *stage_leaf = reinterpret_cast<acc_t *>((reinterpret_cast<uintptr_t>(chunk_bar + 8) + 127) & ~(uintptr_t)127);

    B2S_STAMP(0);
    if (threadIdx.x == 0) {
        mbar_init_cta(load_bar, 1);
        for (uint32_t r = 0; r < C; ++r) mbar_init_cta(&chunk_bar[r], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // split-phase cluster barrier: arrive now, wait only right before the first remote access, so the
    // start-up skew between the CTAs of the cluster hides behind the tree-slice load and the traversal
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");

    const int n_chunks = (T + chunk_trees - 1) / chunk_trees;
    const uint32_t leaf_remote = map_to_rank(leafbuf, 0);
    const uint32_t bar_remote = map_to_rank(&chunk_bar[rank], 0);
    acc_t acc = F64 ? (acc_t)p.base : (acc_t)(float)p.base;

    for (int ch = 0; ch < n_chunks; ++ch) {
        const int slice = ch * (int)C + (int)rank;                 // which trees_per_cta-group this CTA owns
        const int t_lo = slice * TPC;
        const int n_my = max(0, min(TPC, T - t_lo));               // trees of this CTA in this chunk
        const uint32_t node_lo = st.off[min(slice, kMaxSlices)] & ~1u;   // 16-byte aligned bulk source
        const uint32_t node_hi = (st.off[min(slice + 1, kMaxSlices)] + 1u) & ~1u;
        if (warp == 0 && n_my > 0) {
            const uint32_t bytes = (node_hi - node_lo) * 8u;
            if (lane == 0) mbar_expect_tx_cta(load_bar, bytes);
            __syncwarp();
            const unsigned char *src = reinterpret_cast<const unsigned char *>(p.nodes + node_lo);
            for (uint32_t o = (uint32_t)lane * (uint32_t)bulk_piece; o < bytes; o += 32u * (uint32_t)bulk_piece)
                bulk_g2s(smem + o, src + o, min((uint32_t)bulk_piece, bytes - o), load_bar);
        }
        if (rank == 0 && threadIdx.x == 32) {   // arm the per-rank arrival barriers with the bytes each rank pushes
            for (uint32_t r = 1; r < C; ++r) {
                const int n_r = max(0, min(TPC, T - (ch * (int)C + (int)r) * TPC));
                if (n_r > 0) mbar_expect_tx_cta(&chunk_bar[r], (uint32_t)(16 * TPC * sizeof(acc_t)));
                else mbar_arrive_cta(&chunk_bar[r]);
            }
        }
        // overlap with the bulk copy: per-tree offsets and (first chunk) the x tile
        for (int i = threadIdx.x; i <= n_my; i += WARPS * 32) toff[i] = __ldg(p.tree_offset + t_lo + i) - node_lo;
        if (ch == 0) {
            const float *src = X + r0 * F;
            for (int i = threadIdx.x; i < 16 * F; i += WARPS * 32) {
                const int r = i / F, f = i - r * F;
                xs[f * 17 + r] = (r < rows_here) ? __ldg(src + i) : 0.0f;
            }
        }
        __syncthreads();
        if (n_my > 0) {
            uint32_t spins = 0;
            while (!mbar_try_wait_cta(load_bar, (uint32_t)(ch & 1))) {
                if (++spins > B2S_SPIN_LIMIT) __trap();
            }
        }
        B2S_STAMP(1);

        // ---- traverse out of shared memory: this warp's 2*U trees, lane = (row16, half)
        const int tl0 = warp * 2 * U + half;   // local tree index of slot u is tl0 + 2u
        uint32_t base[U];
        uint2 cur[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tl = tl0 + 2 * u;
            base[u] = (tl < n_my) ? toff[tl] : 0u;
            cur[u] = (tl < n_my) ? snodes[base[u]] : make_uint2(0u, 0u);
        }
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
                        const float x = xs[(meta & fmask) * 17 + row16];
                        const float thr = __uint_as_float(cur[u].x);
                        const bool dl = (meta >> fb) & 1u;
                        const bool go_left = (x != x) ? dl : (x < thr);
                        cur[u] = snodes[base[u] + left + (go_left ? 0u : 1u)];
                        any = true;
                    }
                }
                if (!__any_sync(0xffffffffu, any)) break;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tl = tl0 + 2 * u;
            if (tl < n_my) {
                acc_t v;
                if (F64) v = (acc_t)__ldg(p.leaf64 + cur[u].x);
                else v = (acc_t)__uint_as_float(cur[u].x);
                if (rank == 0) leafbuf[row16 * LD + tl] = v;          // rank 0: straight into the leaf matrix
                else stage_leaf[row16 * SLD + tl] = v;                // others: local staging
            }
        }
        // publish: rank 0 arrives locally; every other rank pushes its block with one bulk DSMEM copy that
        // completes bytes on its barrier in rank 0 (generic-proxy writes -> async-proxy read needs the fence)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        if (ch == 0) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");  // peers started, barriers initialised
        __syncthreads();
        if (rank == 0) {
            if (threadIdx.x == 0) mbar_arrive_cta(&chunk_bar[0]);
        } else if (warp == 0 && lane < 16 && n_my > 0) {   // one 16-byte-aligned row segment per lane
            bulk_s2s_cluster(leaf_remote + (uint32_t)(((size_t)lane * LD + (size_t)rank * TPC) * sizeof(acc_t)),
                             stage_leaf + (size_t)lane * SLD, (uint32_t)(TPC * sizeof(acc_t)), bar_remote);
        }
        B2S_STAMP(2);

        // ---- ordered sum in rank 0: rank by rank as their leaf values land
        if (rank == 0 && warp == 0) {
            for (uint32_t r = 0; r < C; ++r) {
                uint32_t spins = 0;
                while (!mbar_try_wait_cluster_acq(&chunk_bar[r], (uint32_t)(ch & 1))) {
                    if (++spins > B2S_SPIN_LIMIT) __trap();
                }
                if (r == 0) B2S_STAMP(3);
                if (r == 1) B2S_STAMP(5);        // rank 1's block landed
                if (r == C - 1) B2S_STAMP(6);    // last rank's block landed
                const int t_first = ch * chunk_trees + (int)r * TPC;
                const int tcount = max(0, min(TPC, T - t_first));
                const acc_t *bp = leafbuf + (size_t)row16 * LD + (size_t)r * TPC;   // this lane's row, rank r's trees
                if (tcount == TPC) {
                    // straight-line, fully unrolled: 128-bit loads of block b+1 are issued among the dependent
                    // adds of block b, so the chain runs at the 4-cycle FADD/DADD issue distance
                    constexpr int BLK = 8;                 // vectors per block
                    constexpr int NBLK = TPC / (VEC * BLK);
                    typedef typename std::conditional<F64, double2, float4>::type vec_t;
                    const vec_t *vp = reinterpret_cast<const vec_t *>(bp);
                    vec_t cur[BLK], nxt[BLK];
#pragma unroll
                    for (int k = 0; k < BLK; ++k) cur[k] = vp[k];
#pragma unroll
                    for (int blk = 0; blk < NBLK; ++blk) {
                        if (blk + 1 < NBLK) {
#pragma unroll
                            for (int k = 0; k < BLK; ++k) nxt[k] = vp[(blk + 1) * BLK + k];
                        }
#pragma unroll
                        for (int k = 0; k < BLK; ++k) {
                            if (F64) {
                                const double2 d = *reinterpret_cast<const double2 *>(&cur[k]);
                                acc = acc + (acc_t)d.x;
                                acc = acc + (acc_t)d.y;
                            } else {
                                const float4 f4 = *reinterpret_cast<const float4 *>(&cur[k]);
                                acc = acc + (acc_t)f4.x;
                                acc = acc + (acc_t)f4.y;
                                acc = acc + (acc_t)f4.z;
                                acc = acc + (acc_t)f4.w;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < BLK; ++k) cur[k] = nxt[k];
                    }
                } else {
                    for (int t = 0; t < tcount; ++t) acc = acc + bp[t];
                }
            }
            B2S_STAMP(4);
        }
        if (ch + 1 < n_chunks) cluster_sync_all();   // leaf matrix / tree slice are reused by the next chunk
    }
    if (rank == 0 && warp == 0 && half == 0 && row16 < rows_here) {
        if (F64) reinterpret_cast<double *>(out)[r0 + row16] = (double)acc / p.divisor;
        else reinterpret_cast<float *>(out)[r0 + row16] = forest_link_f32((float)acc, p.link);
    }
#undef B2S_STAMP
}


// ---------------------------------------------------------------------------------------------
// Kernel D (serving batches, default): "wide" clusters over COMPACT, lane-interleaved tree blocks.
// What bounded kernel C (profiles/r01_forest_staged_phase_timing.txt, r01_ncu_forest_staged.txt):
//   (1) 8-byte leaf slots: 1.37x the algorithmic DRAM bytes and 130 KB per CTA to stage;
//   (2) every leaf value of a 16-row tile funnelled into ONE SM through distributed shared memory, whose
//       ingest rate is ~20 B/clk (57 KB -> ~2900 cycles of the 4400-cycle traverse+publish phase);
//   (3) a per-source-rank barrier round trip (~230 cycles x 8) inside the ordered sum;
//   (4) random 8-byte node reads out of shared memory: ~4-way bank conflicts on every level.
// Here
//   * the model is re-packed at load time into one image per cluster rank, made of BLOCKS of 32 consecutive trees:
//     internal nodes 8 B {f32 threshold | feat, default_left, left ref, right ref}, leaves 4 B (fp32 mode) / 8 B
//     (fp64 mode, the pre-scaled double itself: no leaf64[] indirection) -- the algorithmic bytes of SURVEY.md 8(d) for
//     complete trees; inside a block, node j of tree t sits at [j][t] (lane-interleaved, trees padded to the block's
//     largest): a warp walks 32 trees, lane = tree, so every level's node read hits 32 distinct banks whatever
//     path each lane took;
//   * a cluster of up to 16 CTAs (non-portable size) shares a tile of R rows: rank r stages trees
//     [r*TPC, (r+1)*TPC) (48.6 KB for 64 depth-6 trees) with cp.async.bulk, one mbarrier per block, so the warps
//     of block 0 start walking while block 1 is still in flight; one (tree, row) pair per thread;
//   * ALL-TO-ALL publish: row `i` of the tile is OWNED by rank i / rows_per_rank; a warp holds 32 consecutive trees
//     of one row, so it sends one coalesced 128-byte st.shared::cluster to the owner and arrives (release.cluster) on
//     the owner's mbarrier: every SM ingests only rows_per_rank * T values (4 KB at R = 16, C = 16), and nobody
//     waits for more than its own rows -- no cluster-wide barrier after start-up;
//   * each rank then runs the bit-exactness-mandated sequential chain of ITS rows out of local shared memory
//     (128-bit loads one block ahead of the dependent adds, 64 adds per branch).
// grid = C * ceil(rows / R) CTAs, cluster (C,1,1), 1 CTA per SM.  Forests the compact encoding or a single image
// per rank cannot hold fall back to kernels C / B.
// ---------------------------------------------------------------------------------------------
constexpr int kWideMaxC = 16;
constexpr int kWideMaxBlocks = 8;         // 32-tree blocks per rank image (TPC <= 256)
struct WideParams {
    const unsigned char *images;          // all rank images back to back (16-byte aligned each)
    uint32_t image_off[kWideMaxC + 1];    // byte offset of rank r's image, [C] = end
    int n_trees, n_features, feat_bits, child_bits, tpc, max_depth, link;
    int total_blocks;                     // 32-tree blocks of the whole forest
    int null_mode;                        // developer aid (B2S_FOREST_WIDE_NULL=1): every thread returns at once -> launch floor
    uint32_t sentinel;                    // fp32 bit pattern no leaf of the model carries (fp64: high and low word)
    double base, divisor;
    uint32_t blk_end[kWideMaxC][kWideMaxBlocks];   // byte offset of the end of block b inside rank r's image
};
// rank image: [hdr 16 B: trees here | blocks here | - | -][block table: kWideMaxBlocks x {u32 inode byte offset, u32 leaf
// byte offset, u32 end byte offset, u32 -}][block 0: inodes [ni0][32] uint2, leaves [nl0][32]][block 1 ...]
constexpr int kWideHdrBytes = 16 + kWideMaxBlocks * 16;

template <bool F64>
__global__ void __launch_bounds__(1024, 1)
forest_wide_kernel(const __grid_constant__ WideParams p, const float *__restrict__ X, int64_t n_rows, void *__restrict__ out, int R_log2,
                   int rpr_log2, int image_cap, int bulk_piece, long long *__restrict__ dbg)
{
#define B2S_STAMP(k)                                                                    \
    do {                                                                                \
        if (dbg && blockIdx.x < 8 && threadIdx.x == 0) dbg[blockIdx.x * 8 + (k)] = clock64(); \
    } while (0)
    if (p.null_mode) return;
    using acc_t = typename std::conditional<F64, double, float>::type;
    using bits_t = typename std::conditional<F64, unsigned long long, uint32_t>::type;
    constexpr int VEC = 16 / (int)sizeof(acc_t);
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int F = p.n_features, T = p.n_trees, TPC = p.tpc;
    const int R = 1 << R_log2, rpr = 1 << rpr_log2;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int64_t r0 = (int64_t)(blockIdx.x / C) << R_log2;
    const int rows_here = (int)min((int64_t)R, n_rows - r0);
    const int LD = (int)C * TPC + VEC;     // leaf-matrix row pitch (+16 B: 128-bit row reads of different rows hit different banks)
    const int XP = F | 1;                  // x tile pitch
    acc_t *leafbuf = reinterpret_cast<acc_t *>(smem + image_cap);                               // [rpr][LD], rows this rank owns
    float *xs = reinterpret_cast<float *>(smem + image_cap + (size_t)rpr * LD * sizeof(acc_t));  // [R][XP]
    uint64_t *load_bar = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(xs + (size_t)R * XP) + 15) & ~(uintptr_t)15);  // [kWideMaxBlocks]

    B2S_STAMP(0);
    const uint32_t img_lo = p.image_off[rank], img_bytes = p.image_off[rank + 1] - img_lo;
    const int n_my = max(0, min(T - (int)rank * TPC, TPC));
    const int n_tb = (n_my + 31) >> 5;     // 32-tree blocks of this rank
    const bits_t sentinel = F64 ? (((unsigned long long)p.sentinel << 32) | p.sentinel) : (bits_t)p.sentinel;
    const bool use_bulk = bulk_piece > 0;
    if (!use_bulk) {
        // plain cooperative copy: every thread keeps up to four 16-byte loads in flight (one L2 / DRAM round trip for the
        // whole 49 KB image; the bulk-copy engine showed ~2500 cycles of fixed latency for a copy this small)
        const uint4 *src = reinterpret_cast<const uint4 *>(p.images + img_lo);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        const int n16 = (int)(img_bytes >> 4), stride = (int)blockDim.x;
        for (int i = threadIdx.x; i < n16; i += 4 * stride) {
            uint4 v0 = __ldg(src + i), v1, v2, v3;
            const bool h1 = i + stride < n16, h2 = i + 2 * stride < n16, h3 = i + 3 * stride < n16;
            if (h1) v1 = __ldg(src + i + stride);
            if (h2) v2 = __ldg(src + i + 2 * stride);
            if (h3) v3 = __ldg(src + i + 3 * stride);
            dst[i] = v0;
            if (h1) dst[i + stride] = v1;
            if (h2) dst[i + 2 * stride] = v2;
            if (h3) dst[i + 3 * stride] = v3;
        }
    } else if (threadIdx.x == 0) {
        for (int b = 0; b < n_tb; ++b) mbar_init_cta(&load_bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // one bulk copy per piece, all issued by this thread: block b's bytes are [end of block b-1, end of block b)
        const unsigned char *src = p.images + img_lo;
        uint32_t lo = 0;
        for (int b = 0; b < n_tb; ++b) {
            const uint32_t hi = p.blk_end[rank][b];
            mbar_expect_tx_cta(&load_bar[b], hi - lo);
            for (uint32_t o = lo; o < hi; o += (uint32_t)bulk_piece)
                bulk_g2s(smem + o, src + o, min((uint32_t)bulk_piece, hi - o), &load_bar[b]);
            lo = hi;
        }
    }
    {   // the rows this rank owns start as "nothing has landed": every slot carries the sentinel pattern
        bits_t *lb = reinterpret_cast<bits_t *>(leafbuf);
        for (int i = threadIdx.x; i < rpr * LD; i += blockDim.x) lb[i] = sentinel;
    }
    // split-phase cluster barrier: arrive (sentinels written: release) now, wait right before the first remote store
    __syncwarp();   // .aligned: lane 0 of warp 0 is back from its bulk-copy loop
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    {   // x tile, overlapped with the bulk copies
        const float *src = X + r0 * F;
        const int n = rows_here * F;
        for (int i = threadIdx.x; i < R * F; i += blockDim.x) {
            const int r = i / F, f = i - r * F;
            xs[r * XP + f] = (i < n) ? __ldg(src + i) : 0.0f;
        }
    }
    __syncthreads();

    // ---- traverse out of shared memory: task q = (block of 32 trees, row), lane = tree inside the block
    const int fb = p.feat_bits, cb = p.child_bits;
    const uint32_t fmask = (1u << fb) - 1u, cmask = (1u << cb) - 1u, leafbit = 1u << (cb - 1);
    const uint32_t leaf_local = (uint32_t)__cvta_generic_to_shared(leafbuf);
    const int n_tasks = n_tb << R_log2;
    bool waited = false;
    int tb_ready = -1;
    for (int q = warp; q < n_tasks; q += n_warps) {
        const int row = q & (R - 1), tb = q >> R_log2;
        const int tl = tb * 32 + lane;
        if (use_bulk && tb != tb_ready) {
            uint32_t spins = 0;
            if (tb_ready < 0 && tb > 0) {     // the block table rides with block 0
                while (!mbar_try_wait_cta(&load_bar[0], 0u)) {
                    if (++spins > B2S_SPIN_LIMIT) __trap();
                }
            }
            while (!mbar_try_wait_cta(&load_bar[tb], 0u)) {
                if (++spins > B2S_SPIN_LIMIT) __trap();
            }
            tb_ready = tb;
        }
        if (q == 0) B2S_STAMP(1);
        acc_t v = (acc_t)0;
        if (tl < n_my) {
            const uint32_t *tab = reinterpret_cast<const uint32_t *>(smem + 16) + tb * 4;
            const uint2 *inodes = reinterpret_cast<const uint2 *>(smem + tab[0]) + lane;
            const acc_t *leaves = reinterpret_cast<const acc_t *>(smem + tab[1]) + lane;
            const float *xr = xs + row * XP;
            uint2 nd = inodes[0];
            for (int it = 0; it <= p.max_depth; ++it) {
                const float x = xr[nd.y & fmask];
                const float thr = __uint_as_float(nd.x);
                const bool dl = (nd.y >> fb) & 1u;
                const bool go_left = (x != x) ? dl : (x < thr);
                const uint32_t ref = (nd.y >> (go_left ? fb + 1 : fb + 1 + cb)) & cmask;
                if (ref & leafbit) {
                    v = leaves[(ref & (leafbit - 1u)) * 32u];
                    break;
                }
                nd = inodes[ref * 32u];
            }
        }
        __syncwarp();
        if (!waited) {   // peers have started and filled their leaf matrices with sentinels before the first remote store
            if (q == 0) B2S_STAMP(5);
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
            waited = true;
            if (q == 0) B2S_STAMP(6);
        }
        const int owner = row >> rpr_log2, lr = row & (rpr - 1);
        uint32_t remote_leaf;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote_leaf) : "r"(leaf_local), "r"((uint32_t)owner));
        if (tl < n_my)
            st_cluster(remote_leaf + (uint32_t)(((size_t)lr * LD + (size_t)rank * TPC + tl) * sizeof(acc_t)), v);
    }
    if (!waited) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    B2S_STAMP(2);

    // ---- ordered sum: lane i of warp 0 adds the column of row rank*rpr + i in tree order.  A value IS its own
    // arrival flag: a slot still holding the sentinel has not landed yet (the leaf values of a model never carry that
    // bit pattern: checked when the images are built).  The whole warp scans the rows this rank owns with 128-bit
    // volatile loads until nothing is missing -- normally one pass of ~100 cycles -- instead of a release + mbarrier
    // round trip per publishing warp (~1400 cycles measured between the last remote store and the first add).
    if (warp == 0) {
        {
            const uint32_t base_addr = (uint32_t)__cvta_generic_to_shared(leafbuf);
            const int vec_per_row = (T + VEC - 1) / VEC;           // the pad slots of the last vector are skipped below
            uint32_t spins = 0;
            for (;;) {
                bool ok = true;
                for (int r = 0; r < rpr; ++r) {
                    for (int i = lane; i < vec_per_row; i += 32) {
                        const uint32_t a = base_addr + (uint32_t)(((size_t)r * LD) * sizeof(acc_t)) + (uint32_t)i * 16u;
                        uint32_t w0, w1, w2, w3;
                        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(a) : "memory");
                        const int t0 = i * VEC;
                        if (F64) {
                            ok = ok && !(w0 == p.sentinel && w1 == p.sentinel);
                            if (t0 + 1 < T) ok = ok && !(w2 == p.sentinel && w3 == p.sentinel);
                        } else {
                            ok = ok && w0 != p.sentinel;
                            if (t0 + 1 < T) ok = ok && w1 != p.sentinel;
                            if (t0 + 2 < T) ok = ok && w2 != p.sentinel;
                            if (t0 + 3 < T) ok = ok && w3 != p.sentinel;
                        }
                    }
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if (++spins > (B2S_SPIN_LIMIT >> 6)) __trap();
            }
        }
        B2S_STAMP(3);
        const int row = ((int)rank << rpr_log2) + lane;
        if (lane < rpr && row < rows_here) {
            acc_t acc = F64 ? (acc_t)p.base : (acc_t)(float)p.base;
            const acc_t *bp = leafbuf + (size_t)lane * LD;
            typedef typename std::conditional<F64, double2, float4>::type vec_t;
            constexpr int BLK = 4;                    // vectors per block (16 fp32 / 8 fp64 values): 64 cycles of dependent adds cover the 29-cycle LDS
            const vec_t *vp = reinterpret_cast<const vec_t *>(bp);
            const int n_blk = T / (VEC * BLK);
            vec_t cur[BLK], nxt[BLK];
            if (n_blk > 0) {
#pragma unroll
                for (int k = 0; k < BLK; ++k) cur[k] = vp[k];
            }
#pragma unroll 4
            for (int blk = 0; blk < n_blk; ++blk) {
                if (blk + 1 < n_blk) {                // the next block's loads issue among this block's dependent adds
#pragma unroll
                    for (int k = 0; k < BLK; ++k) nxt[k] = vp[(blk + 1) * BLK + k];
                }
#pragma unroll
                for (int k = 0; k < BLK; ++k) {
                    if (F64) {
                        const double2 d2 = *reinterpret_cast<const double2 *>(&cur[k]);
                        acc = acc + (acc_t)d2.x;
                        acc = acc + (acc_t)d2.y;
                    } else {
                        const float4 f4 = *reinterpret_cast<const float4 *>(&cur[k]);
                        acc = acc + (acc_t)f4.x;
                        acc = acc + (acc_t)f4.y;
                        acc = acc + (acc_t)f4.z;
                        acc = acc + (acc_t)f4.w;
                    }
                }
#pragma unroll
                for (int k = 0; k < BLK; ++k) cur[k] = nxt[k];
            }
            for (int t = n_blk * VEC * BLK; t < T; ++t) acc = acc + bp[t];
            if (F64) reinterpret_cast<double *>(out)[r0 + row] = (double)acc / p.divisor;
            else reinterpret_cast<float *>(out)[r0 + row] = forest_link_f32((float)acc, p.link);
        }
        B2S_STAMP(4);
    }
#undef B2S_STAMP
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
        else reinterpret_cast<float *>(out)[r0 + threadIdx.x] = forest_link_f32((float)acc, p.link);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kStWarps = 16;               // staged kernel: 512 threads per CTA
constexpr int kStU32 = 4, kStU64 = 2;      // => 128 / 64 trees per CTA per chunk
constexpr int kClWarps = 32;               // 1024 threads per CTA, one CTA per SM
constexpr int kClU32 = 4, kClU64 = 2;      // trees in flight per warp (fp32 / fp64 leaf matrix)
constexpr int kRowsBlock = 128;
constexpr int kRowsU = 4;
constexpr int64_t kClusterMaxRows = 4096;  // above this the one-row-per-thread kernel is used
constexpr size_t kLeafBudget = 160 * 1024; // bytes of shared memory for the leaf matrix

struct ForestModel : Model {
    ForestParams p{};
    void *d_blob = nullptr;
    bool f64 = false;
    int max_smem_optin = 0;
    long long *dbg_stamps = nullptr;  // device buffer [8 CTAs][8 phases], only with B2S_FOREST_TIMING=1
    // wide-cluster kernel over compact per-rank images (default serving path)
    bool wide_ok = false;
    WideParams wp{};
    void *d_images = nullptr;
    int wide_C = 1, wide_cap = 0;
    size_t wide_fixed_smem = 0;   // image + barrier; the leaf matrix and the x tile depend on the row tile
    // staged (shared-memory resident) fast path
    bool staged_ok = false;
    SliceTable slices{};
    int staged_C = 1, staged_slice_cap = 0;
    size_t staged_smem = 0;

    ~ForestModel() override
    {
        if (d_blob) { cudaSetDevice(device); cudaFree(d_blob); }
        if (d_images) { cudaSetDevice(device); cudaFree(d_images); }
        if (dbg_stamps) cudaFree(dbg_stamps);
    }

    int debug_read(long long *out64) override
    {
        if (!dbg_stamps) return fail(B2S_ERR_INVALID, "set B2S_FOREST_TIMING=1 before loading the model");
        B2S_CUDA(cudaSetDevice(device));
        B2S_CUDA(cudaDeviceSynchronize());
        B2S_CUDA(cudaMemcpy(out64, dbg_stamps, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
        return 0;
    }

    size_t scratch_bytes(int64_t, int64_t) const override { return 256; }

    template <bool F64, int U>
    int launch_cluster(cudaStream_t st, const float *X, int64_t n_rows, void *out)
    {
        using acc_t = typename std::conditional<F64, double, float>::type;
        const int F = p.n_features, T = p.n_trees;
        int C = 1;
        while (C < 8 && C * kClWarps * U < T) C *= 2;
        const int per_round = C * kClWarps * U;
        const int max_rounds = (int)(kLeafBudget / ((size_t)per_round * 32 * sizeof(acc_t)));
        int rounds = (T + per_round - 1) / per_round;
        if (rounds > max_rounds) rounds = max_rounds;
        if (rounds < 1) rounds = 1;
        const size_t leaf_bytes = (size_t)per_round * rounds * 32 * sizeof(acc_t);
        const size_t xs_bytes = (size_t)F * 33 * sizeof(float);
        const int x_in_smem = leaf_bytes + xs_bytes + 1024 <= (size_t)max_smem_optin ? 1 : 0;
        const size_t smem = leaf_bytes + (x_in_smem ? xs_bytes : 0);
        const unsigned tiles = (unsigned)((n_rows + 31) / 32);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(tiles * C, 1, 1);
        cfg.blockDim = dim3(kClWarps * 32, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        B2S_CUDA(cudaLaunchKernelEx(&cfg, forest_cluster_kernel<F64, kClWarps, U>, p, X, n_rows, out, x_in_smem, rounds,
                                    dbg_stamps));
        return 0;
    }


    // row tile of the wide kernel for a batch: 16 rows per cluster at serving sizes, 32 for big batches
    int wide_rows_per_tile(int64_t n_rows) const
    {
        static const int forced = []() { const char *e = getenv("B2S_FOREST_WIDE_R"); const int v = e ? atoi(e) : 0; return (v == 16 || v == 32) ? v : 0; }();
        if (forced) return forced;
        return n_rows > 1024 ? 32 : 16;
    }
    size_t wide_smem(int R) const
    {
        const size_t esz = f64 ? 8 : 4, vec = 16 / esz;
        const int rpr = (R + wide_C - 1) / wide_C;
        return (size_t)wide_cap + (size_t)rpr * ((size_t)wide_C * wp.tpc + vec) * esz + (size_t)R * (size_t)(p.n_features | 1) * 4 + 32 + kWideMaxBlocks * 8;
    }

    template <bool F64>
    int launch_wide(cudaStream_t st, const float *X, int64_t n_rows, void *out)
    {
        const int R = wide_rows_per_tile(n_rows);
        const int rpr = (R + wide_C - 1) / wide_C;          // R in {16, 32}, C a power of two: rpr is one too
        const unsigned tiles = (unsigned)((n_rows + R - 1) / R);
        int threads = ((wp.tpc + 31) / 32) * R * 32;     // one (tree, row) pair per thread when it fits
        if (threads > 1024) threads = 1024;
        static const int thread_cap = []() { const char *e = getenv("B2S_FOREST_WIDE_THREADS"); return e ? atoi(e) : 1024; }();
        if (thread_cap >= 64 && threads > thread_cap) threads = thread_cap;
        if (threads < 64) threads = 64;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(tiles * wide_C, 1, 1);
        cfg.blockDim = dim3(threads, 1, 1);
        cfg.dynamicSmemBytes = wide_smem(R);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = wide_C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        static const int bulk_piece = []() {
            const char *e = getenv("B2S_FOREST_BULK_PIECE");
            int v = e ? atoi(e) : 0;      // 0: plain cooperative loads (default); >= 1024: cp.async.bulk pieces of that size
            return (v >= 1024 && v % 16 == 0) ? v : 0;
        }();
        int R_log2 = 0, rpr_log2 = 0;
        while ((1 << R_log2) < R) ++R_log2;
        while ((1 << rpr_log2) < rpr) ++rpr_log2;
        B2S_CUDA(cudaLaunchKernelEx(&cfg, forest_wide_kernel<F64>, wp, X, n_rows, out, R_log2, rpr_log2, wide_cap, bulk_piece, dbg_stamps));
        return 0;
    }

    // Re-pack the v1 node array into one compact image per cluster rank (see kernel D) and upload them.
    int build_wide(const uint32_t *toff, const uint64_t *nodes, const double *leaf64, const ForestBlobHeader &h, int max_depth)
    {
        const char *off = getenv("B2S_FOREST_WIDE");
        if (off && off[0] == '0') return 0;
        const int fb = (int)h.feat_bits;
        const uint32_t T = h.n_trees;
        // per-tree internal / leaf counts -> index width of the child references
        uint32_t max_cnt = 1;
        for (uint32_t t = 0; t < T; ++t) {
            uint32_t ni = 0, nl = 0;
            for (uint32_t i = toff[t]; i < toff[t + 1]; ++i) (((uint32_t)(nodes[i] >> 32) >> (fb + 1)) ? ni : nl) += 1;
            if (ni > max_cnt) max_cnt = ni;
            if (nl > max_cnt) max_cnt = nl;
        }
        int idx_bits = 1;
        while ((1u << idx_bits) < max_cnt) ++idx_bits;
        const int cb = idx_bits + 1;                       // + leaf flag
        if (fb + 1 + 2 * cb > 32) return 0;                // trees too large for the compact node: kernels C / B serve the model
        // cluster size: the largest the device schedules for this kernel, no larger than the forest can use
        int c_max = 8;
        {
            const char *e = getenv("B2S_FOREST_WIDE_C");
            int want = e ? atoi(e) : 16;
            if (want > 8) {
                cudaError_t ce = f64 ? cudaFuncSetAttribute(forest_wide_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)
                                     : cudaFuncSetAttribute(forest_wide_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                if (ce == cudaSuccess) c_max = 16; else cudaGetLastError();
            } else if (want >= 1) {
                c_max = want;
            }
        }
        int C = 1;
        while (C < c_max && C * 32 < (int)T) C *= 2;
        for (;; C /= 2) {
            const int tpc = (int)round_up(((int64_t)T + C - 1) / C, 32);
            // images: per rank, blocks of 32 trees, node j of tree t at [j][t]
            std::vector<std::vector<unsigned char>> img(C);
            size_t cap = 0;
            const size_t lsz = f64 ? 8 : 4;
            int total_blocks = 0;
            bool too_many_blocks = tpc / 32 > kWideMaxBlocks;
            for (int r = 0; r < C && !too_many_blocks; ++r) {
                const uint32_t t_lo = (uint32_t)r * (uint32_t)tpc;
                if (t_lo >= T) continue;                   // empty rank: no image, no bytes
                const uint32_t t_hi = t_lo + (uint32_t)tpc < T ? t_lo + (uint32_t)tpc : T;
                const int n_blocks = (int)((t_hi - t_lo + 31) / 32);
                total_blocks += n_blocks;
                std::vector<unsigned char> &im = img[r];
                im.assign(kWideHdrBytes, 0);
                uint32_t hdr[4] = {t_hi - t_lo, (uint32_t)n_blocks, 0u, 0u};
                memcpy(im.data(), hdr, 16);
                for (int bk = 0; bk < n_blocks; ++bk) {
                    const uint32_t b_lo = t_lo + (uint32_t)bk * 32u, b_hi = b_lo + 32u < t_hi ? b_lo + 32u : t_hi;
                    // per tree: compact internal / leaf arrays
                    std::vector<std::vector<uint64_t>> tin(32);
                    std::vector<std::vector<unsigned char>> tlf(32);
                    size_t ni_max = 1, nl_max = 1;
                    for (uint32_t t = b_lo; t < b_hi; ++t) {
                        const uint32_t s = toff[t], n = toff[t + 1] - s;
                        std::vector<uint32_t> loc(n);
                        uint32_t ni = 0, nl = 0;
                        for (uint32_t i = 0; i < n; ++i) loc[i] = ((uint32_t)(nodes[s + i] >> 32) >> (fb + 1)) ? ni++ : nl++;
                        auto ref_of = [&](uint32_t i) -> uint32_t {
                            const bool internal = ((uint32_t)(nodes[s + i] >> 32) >> (fb + 1)) != 0;
                            return internal ? loc[i] : (loc[i] | (1u << (cb - 1)));
                        };
                        std::vector<uint64_t> &in = tin[t - b_lo];
                        std::vector<unsigned char> &lf = tlf[t - b_lo];
                        if (ni == 0)   // a tree that is a single leaf: one dummy split whose both sides are leaf 0
                            in.push_back(((uint64_t)((1u << (cb - 1)) << (fb + 1) | (1u << (cb - 1)) << (fb + 1 + cb))) << 32);
                        for (uint32_t i = 0; i < n; ++i) {
                            const uint32_t val = (uint32_t)(nodes[s + i] & 0xffffffffu), meta = (uint32_t)(nodes[s + i] >> 32);
                            const uint32_t left = meta >> (fb + 1);
                            if (left) {
                                const uint32_t m2 = (meta & ((1u << (fb + 1)) - 1u)) | ref_of(left) << (fb + 1) | ref_of(left + 1) << (fb + 1 + cb);
                                in.push_back((uint64_t)val | ((uint64_t)m2 << 32));
                            } else if (f64) {
                                const double dv = leaf64[val];
                                const unsigned char *bb = reinterpret_cast<const unsigned char *>(&dv);
                                lf.insert(lf.end(), bb, bb + 8);
                            } else {
                                const unsigned char *bb = reinterpret_cast<const unsigned char *>(&val);
                                lf.insert(lf.end(), bb, bb + 4);
                            }
                        }
                        if (in.size() > ni_max) ni_max = in.size();
                        if (lf.size() / lsz > nl_max) nl_max = lf.size() / lsz;
                    }
                    const size_t off_in = (size_t)round_up((int64_t)im.size(), 16);
                    const size_t off_lf = off_in + ni_max * 32 * 8;
                    const size_t end = (size_t)round_up((int64_t)(off_lf + nl_max * 32 * lsz), 16);
                    im.resize(end, 0);
                    for (uint32_t t = 0; t < b_hi - b_lo; ++t) {
                        for (size_t j = 0; j < tin[t].size(); ++j) memcpy(im.data() + off_in + (j * 32 + t) * 8, &tin[t][j], 8);
                        for (size_t j = 0; j < tlf[t].size() / lsz; ++j) memcpy(im.data() + off_lf + (j * 32 + t) * lsz, tlf[t].data() + j * lsz, lsz);
                    }
                    const uint32_t ent[4] = {(uint32_t)off_in, (uint32_t)off_lf, (uint32_t)end, 0u};
                    memcpy(im.data() + 16 + (size_t)bk * 16, ent, 16);
                    wp.blk_end[r][bk] = (uint32_t)end;
                }
                if (im.size() > cap) cap = im.size();
            }
            if (too_many_blocks) return 0;   // more than kWideMaxBlocks blocks per rank: kernels C / B serve the model
            wp.total_blocks = total_blocks;
            {   // a bit pattern (a quiet NaN with a payload) that no leaf of THIS model carries marks "not landed yet"
                uint32_t sent = 0x7fc5a5a5u;
                for (int tries = 0; tries < 64; ++tries, sent += 0x1003u) {
                    bool clash = false;
                    for (uint32_t i = 0; i < h.n_nodes && !clash; ++i) {
                        const uint32_t val = (uint32_t)(nodes[i] & 0xffffffffu), meta = (uint32_t)(nodes[i] >> 32);
                        if (meta >> (fb + 1)) continue;
                        if (!f64) clash = val == sent;
                        else {
                            unsigned long long bits;
                            memcpy(&bits, &leaf64[val], 8);
                            clash = bits == (((unsigned long long)sent << 32) | sent);
                        }
                    }
                    if (!clash) break;
                }
                wp.sentinel = sent;
            }
            {
                const char *nm = getenv("B2S_FOREST_WIDE_NULL");
                wp.null_mode = (nm && nm[0] == '1') ? 1 : 0;
            }
            wide_C = C;
            wide_cap = (int)round_up((int64_t)cap, 128);
            wp.tpc = tpc;
            const bool fits = wide_smem(32) <= (size_t)max_smem_optin;
            bool schedulable = fits;
            if (fits) {   // can the device co-schedule a cluster of C CTAs with this much shared memory?
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(C, 1, 1);
                cfg.blockDim = dim3(1024, 1, 1);
                cfg.dynamicSmemBytes = wide_smem(32);
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = C;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                int n_clusters = 0;
                cudaError_t ce = f64 ? cudaOccupancyMaxActiveClusters(&n_clusters, forest_wide_kernel<true>, &cfg)
                                     : cudaOccupancyMaxActiveClusters(&n_clusters, forest_wide_kernel<false>, &cfg);
                if (ce != cudaSuccess) { cudaGetLastError(); n_clusters = 0; }
                schedulable = n_clusters >= 1;
            }
            if (schedulable) {
                size_t total = 0;
                for (int r = 0; r < C; ++r) { wp.image_off[r] = (uint32_t)total; total += img[r].size(); }
                for (int r = C; r <= kWideMaxC; ++r) wp.image_off[r] = (uint32_t)total;
                B2S_CUDA(cudaMalloc(&d_images, total + 256));
                B2S_CUDA(cudaMemset(d_images, 0, total + 256));
                for (int r = 0; r < C; ++r)
                    if (!img[r].empty())
                        B2S_CUDA(cudaMemcpy(static_cast<unsigned char *>(d_images) + wp.image_off[r], img[r].data(), img[r].size(), cudaMemcpyHostToDevice));
                wp.images = static_cast<const unsigned char *>(d_images);
                wp.n_trees = (int)T;
                wp.n_features = (int)h.n_features;
                wp.feat_bits = fb;
                wp.child_bits = cb;
                wp.max_depth = max_depth;
                wp.link = (int)h.link;
                wp.base = h.base;
                wp.divisor = h.divisor;
                wide_ok = true;
                return 0;
            }
            if (C == 1) return 0;   // not even a single CTA can hold the forest: kernels C / B / A serve the model
            // halving C doubles the image: only worth retrying when the cluster could not be scheduled
            if (!fits) return 0;
        }
    }

    template <bool F64, int U>
    int launch_staged(cudaStream_t st, const float *X, int64_t n_rows, void *out)
    {
        const unsigned tiles = (unsigned)((n_rows + 15) / 16);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(tiles * staged_C, 1, 1);
        cfg.blockDim = dim3(kStWarps * 32, 1, 1);
        cfg.dynamicSmemBytes = staged_smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = staged_C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        static const int bulk_piece = []() {
            const char *e = getenv("B2S_FOREST_BULK_PIECE");
            int v = e ? atoi(e) : 8192;
            return (v >= 1024 && v % 16 == 0) ? v : 8192;
        }();
        B2S_CUDA(cudaLaunchKernelEx(&cfg, forest_staged_kernel<F64, kStWarps, U>, p, slices, X, n_rows, out,
                                    staged_slice_cap, bulk_piece, dbg_stamps));
        return 0;
    }

    int launch(cudaStream_t st, int64_t n_rows, const void *const *d_in, void *const *d_out,
               const int64_t *, void *, size_t, const LaunchInfo &) override
    {
        if (n_rows <= 0) return 0;
        const float *X = static_cast<const float *>(d_in[0]);
        void *out = d_out[0];
        const int F = p.n_features;
        if (n_rows <= kClusterMaxRows && wide_ok) {
            if (f64) B2S_TRY((launch_wide<true>(st, X, n_rows, out)));
            else B2S_TRY((launch_wide<false>(st, X, n_rows, out)));
        } else if (n_rows <= kClusterMaxRows && staged_ok) {
            if (f64) B2S_TRY((launch_staged<true, kStU64>(st, X, n_rows, out)));
            else B2S_TRY((launch_staged<false, kStU32>(st, X, n_rows, out)));
        } else if (n_rows <= kClusterMaxRows) {
            if (f64) B2S_TRY((launch_cluster<true, kClU64>(st, X, n_rows, out)));
            else B2S_TRY((launch_cluster<false, kClU32>(st, X, n_rows, out)));
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
            B2S_CUDA(cudaGetLastError());
        }
        count_launch();
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
    if (h.link > 1 || (h.link != 0 && h.acc_mode != 0)) return fail(B2S_ERR_INVALID, "forest blob: bad link");
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
    // device layout: [tree_offset | nodes (+16 B pad for 16-byte bulk copies) | leaf64], 256-byte aligned sections
    const size_t d_off_nodes = (size_t)round_up((int64_t)off_bytes, 256);
    const size_t d_off_leaf = (size_t)round_up((int64_t)(d_off_nodes + node_bytes + 16), 256);
    const size_t d_total = d_off_leaf + leaf_bytes + 256;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, d_total);
    if (e != cudaSuccess) { delete m; return fail_cuda(e, "cudaMalloc(forest)"); }
    unsigned char *d = static_cast<unsigned char *>(m->d_blob);
    e = cudaMemset(d, 0, d_total);
    if (e == cudaSuccess) e = cudaMemcpy(d, toff, (size_t)(h.n_trees + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + d_off_nodes, nodes, node_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && leaf_bytes)
        e = cudaMemcpy(d + d_off_leaf, base + sizeof(h) + off_bytes + node_bytes, leaf_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete m; return fail_cuda(e, "cudaMemcpy(forest)"); }
    m->p.tree_offset = reinterpret_cast<const uint32_t *>(d);
    m->p.nodes = reinterpret_cast<const uint2 *>(d + d_off_nodes);
    m->p.leaf64 = reinterpret_cast<const double *>(d + d_off_leaf);
    m->p.n_trees = (int)h.n_trees;
    m->p.n_features = (int)h.n_features;
    m->p.feat_bits = fb;
    m->p.max_depth = max_depth;
    m->p.base = h.base;
    m->p.divisor = h.divisor;
    m->p.link = (int)h.link;

    const char *dbg_env = getenv("B2S_FOREST_TIMING");
    if (dbg_env && dbg_env[0] == '1') {
        cudaMalloc(reinterpret_cast<void **>(&m->dbg_stamps), 64 * sizeof(long long));
        cudaMemset(m->dbg_stamps, 0, 64 * sizeof(long long));
    }
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    m->max_smem_optin = optin;
    const int want = optin;  // 227 KB on sm_100: the staged kernel needs ~203 KB for the configs[1] forest
    cudaFuncSetAttribute(forest_cluster_kernel<false, kClWarps, kClU32>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_cluster_kernel<true, kClWarps, kClU64>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_rows_kernel<false, kRowsBlock, kRowsU>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_rows_kernel<true, kRowsBlock, kRowsU>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    m->max_smem_optin = want;
    cudaFuncSetAttribute(forest_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    {
        const int wrc = m->build_wide(toff, nodes, reinterpret_cast<const double *>(base + sizeof(h) + off_bytes + node_bytes), h, max_depth);
        if (wrc != 0) { delete m; return wrc; }
    }
    cudaFuncSetAttribute(forest_staged_kernel<false, kStWarps, kStU32>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    cudaFuncSetAttribute(forest_staged_kernel<true, kStWarps, kStU64>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    {   // can every CTA keep its slice of the forest in shared memory?
        const int tpc = kStWarps * 2 * (m->f64 ? kStU64 : kStU32);
        const int n_slices = ((int)h.n_trees + tpc - 1) / tpc;
        const char *no_staged = getenv("B2S_FOREST_NO_STAGED");
        if (n_slices <= kMaxSlices && !(no_staged && no_staged[0] == '1')) {
            uint32_t max_nodes = 0;
            for (int k = 0; k <= kMaxSlices; ++k) {
                const uint32_t t = (uint32_t)k * (uint32_t)tpc;
                m->slices.off[k] = toff[t < h.n_trees ? t : h.n_trees];
            }
            for (int k = 0; k < n_slices; ++k) {
                const uint32_t lo = m->slices.off[k] & ~1u, hi = (m->slices.off[k + 1] + 1u) & ~1u;
                if (hi - lo > max_nodes) max_nodes = hi - lo;
            }
            int C = 1;
            while (C < 8 && C * tpc < (int)h.n_trees) C *= 2;
            const size_t cap = (size_t)round_up((int64_t)max_nodes * 8 + 16, 128);
            const size_t esz = m->f64 ? 8 : 4, vec = 16 / esz;
            const size_t leaf = (size_t)16 * ((size_t)C * tpc + vec) * esz;
            const size_t smem = cap + leaf + (size_t)h.n_features * 17 * 4 + (size_t)(tpc + 1) * 4 + 16 + 9 * 8 + 256 +
                                (size_t)16 * ((size_t)tpc + vec) * esz;   // + local leaf staging block
            if (smem <= (size_t)want) {
                m->staged_ok = true;
                m->staged_C = C;
                m->staged_slice_cap = (int)cap;
                m->staged_smem = smem;
            }
        }
    }

    b2s_model_info &info = m->info;
    info.kind = B2S_MODEL_FOREST;
    info.n_inputs = 1;
    info.n_outputs = 1;
    info.in_dtype[0] = B2S_F32;
    info.out_dtype[0] = m->f64 ? B2S_F64 : B2S_F32;
    info.in_row_elems[0] = h.n_features;
    info.out_row_elems[0] = 1;
    info.weight_bytes = (int64_t)d_total;
    info.algo_bytes_fixed = algo_fixed;
    info.algo_bytes_per_row = (int64_t)h.n_features * 4 + (m->f64 ? 8 : 4);
    *out = m;
    return 0;
}

}  // namespace b2s
