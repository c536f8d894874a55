// llm_attention.cu -- causal grouped-query attention over the KV cache for the decoder-only LLM endpoint
// (BASELINE.json configs[4]; the reference delegates it to vLLM's paged attention,
// clearml_serving/serving/preprocess_service.py:1097-1348).  head_dim = 128, bf16 in, fp32 softmax/accumulate.
//
// KV cache layout (one layer): PAGED.  K and V each [page][kv_head][64][128] bf16; a sequence (KV slot) owns the pages its
// row of the page table names: token `pos` of slot `s` lives in page page_table[s * pages_per_seq + pos / 64] at row
// pos % 64.  A page is exactly one 64-key block of these kernels, so a block of one (sequence, kv head) is still one
// contiguous [64, 128] matrix streamed with full 256-byte rows -- the indirection costs one 4-byte load per block.
//
//  * prefill: one CTA = (64-query tile, q head, sequence), 4 warps x 16 query rows; K/V blocks of 64 keys are
//    double-buffered through shared memory with cp.async; S = QK^T and O += PV on mma.sync m16n8k16 (bf16) with
//    the online softmax in registers; only key blocks at or below the diagonal are visited.
//    FLOPs: 4 * 128 * S^2 / 2 per (sequence, q head).  ~1 % of the prefill FLOPs of Llama-3-8B at S=512, so
//    it stays on the legacy tensor path; the tcgen05 budget is in the GEMMs.
//  * decode: one CTA = (sequence, kv head); the G query heads of the group are the rows of one 16-row MMA tile,
//    keys stream through a 3-stage cp.async ring, RoPE + KV append are fused into the prologue (see below).
//    A first CUDA-core version (one lane per 4 dims, shuffle reductions) was issue-bound at ~25 us per layer
//    for 32 x 8 x 576 cached keys; HBM-bound target: 512 bytes per cached token per kv head.
#include "common.cuh"

#include <cuda_bf16.h>

#include <mutex>

namespace b2s {

constexpr int LA_D = 128;
constexpr int LA_BQ = 64;
constexpr int LA_BK = 64;
constexpr int LA_LD = LA_D + 8;   // padded smem row: 272 B stride, conflict-free ldmatrix

__device__ __forceinline__ void la_ldmatrix_x4(uint32_t (&r)[4], const void *smem_ptr)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void la_ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_ptr)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void la_mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t la_pack_bf16(float lo, float hi)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
// [64 rows x 128] bf16 tile -> padded smem, 16-byte cp.async, rows >= valid zero-filled
__device__ __forceinline__ void la_load_tile_async(__nv_bfloat16 *dst, const __nv_bfloat16 *src, int64_t ld_src,
                                                   int valid_rows, int tid)
{
    for (int i = tid; i < 64 * 16; i += 128) {
        const int r = i >> 4, c = (i & 15) * 8;
        const bool ok = r < valid_rows;
        const __nv_bfloat16 *g = src + (int64_t)(ok ? r : 0) * ld_src + c;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + r * LA_LD + c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(g), "r"(ok ? 16 : 0) : "memory");
    }
}
__device__ __forceinline__ void la_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void la_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int LA_PREFILL_SMEM = (LA_BQ + 4 * LA_BK) * LA_LD * 2;   // Q + 2 x (K, V)

__global__ void __launch_bounds__(128)
llm_attn_prefill_kernel(const __nv_bfloat16 *__restrict__ qkv, int ld_qkv, const __nv_bfloat16 *__restrict__ kc,
                        const __nv_bfloat16 *__restrict__ vc, const int32_t *__restrict__ cu_seqlens,
                        const int32_t *__restrict__ slots, const int32_t *__restrict__ page_table, int pages_per_seq,
                        __nv_bfloat16 *__restrict__ out, int ld_out, int group, int kvh_r, float scale_log2e)
{
    extern __shared__ __align__(16) unsigned char la_smem[];
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(la_smem);
    __nv_bfloat16 *Ks2 = Qs + LA_BQ * LA_LD;            // [2][64 * LD]
    __nv_bfloat16 *Vs2 = Ks2 + 2 * LA_BK * LA_LD;

    const int b = blockIdx.z, h = blockIdx.y;
    const int qt = gridDim.x - 1 - blockIdx.x;          // longest (diagonal-most) tiles first
    const int s0 = __ldg(cu_seqlens + b);
    const int S = __ldg(cu_seqlens + b + 1) - s0;
    const int q0 = qt * LA_BQ;
    if (q0 >= S) return;
    const int slot = __ldg(slots + b), kvh = h / group;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int32_t *pages = page_table + (int64_t)slot * pages_per_seq;
    const int k_end = min(S, q0 + LA_BQ);                // keys [0, k_end) can be visible to this tile

    la_load_tile_async(Qs, qkv + (int64_t)(s0 + q0) * ld_qkv + h * LA_D, ld_qkv, min(LA_BQ, S - q0), tid);
    auto issue_block = [&](int k0, int buf) {
        const int kv_valid = min(LA_BK, k_end - k0);
        const int64_t blk = ((int64_t)__ldg(pages + (k0 >> 6)) * kvh_r + kvh) * (LA_BK * LA_D);   // one page = one key block
        la_load_tile_async(Ks2 + buf * LA_BK * LA_LD, kc + blk, LA_D, kv_valid, tid);
        la_load_tile_async(Vs2 + buf * LA_BK * LA_LD, vc + blk, LA_D, kv_valid, tid);
        la_commit();
    };
    issue_block(0, 0);   // group 0 = Q + first K/V block

    uint32_t qa[8][4];
    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;   // query positions of this thread's two rows

    int buf = 0;
    for (int k0 = 0; k0 < k_end; k0 += LA_BK, buf ^= 1) {
        const bool more = k0 + LA_BK < k_end;
        if (more) issue_block(k0 + LA_BK, buf ^ 1);
        if (more) la_wait<1>(); else la_wait<0>();
        __syncthreads();
        if (k0 == 0) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                la_ldmatrix_x4(qa[kk], Qs + (warp * 16 + (lane & 15)) * LA_LD + kk * 16 + (lane >> 4) * 8);
        }
        const __nv_bfloat16 *Ks = Ks2 + buf * LA_BK * LA_LD, *Vs = Vs2 + buf * LA_BK * LA_LD;

        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t kb[4];
                la_ldmatrix_x4(kb, Ks + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LA_LD + kk * 16 + ((lane >> 3) & 1) * 8);
                la_mma_bf16(s[2 * np], qa[kk], kb[0], kb[1]);
                la_mma_bf16(s[2 * np + 1], qa[kk], kb[2], kb[3]);
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int key = k0 + n * 8 + 2 * t;
            s[n][0] = key <= r_lo ? s[n][0] * scale_log2e : -INFINITY;
            s[n][1] = key + 1 <= r_lo ? s[n][1] * scale_log2e : -INFINITY;
            s[n][2] = key <= r_hi ? s[n][2] * scale_log2e : -INFINITY;
            s[n][3] = key + 1 <= r_hi ? s[n][3] * scale_log2e : -INFINITY;
            mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
        }
        float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = (m_new == -INFINITY) ? 1.f : exp2f(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        const float m0 = (m_run[0] == -INFINITY) ? 0.f : m_run[0];
        const float m1 = (m_run[1] == -INFINITY) ? 0.f : m_run[1];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = exp2f(s[n][0] - m0);
            s[n][1] = exp2f(s[n][1] - m0);
            s[n][2] = exp2f(s[n][2] - m1);
            s[n][3] = exp2f(s[n][3] - m1);
            rs[0] += s[n][0] + s[n][1];
            rs[1] += s[n][2] + s[n][3];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
            l_run[r] = l_run[r] * corr[r] + rs[r];
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            o[n][0] *= corr[0]; o[n][1] *= corr[0];
            o[n][2] *= corr[1]; o[n][3] *= corr[1];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // 16 keys per k-step
            uint32_t pa[4];
            pa[0] = la_pack_bf16(s[2 * j][0], s[2 * j][1]);
            pa[1] = la_pack_bf16(s[2 * j][2], s[2 * j][3]);
            pa[2] = la_pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
            pa[3] = la_pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
            for (int np = 0; np < 8; ++np) {   // pairs of 8-wide d tiles
                uint32_t vb[4];
                la_ldmatrix_x4_trans(vb, Vs + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LA_LD + np * 16 + (lane >> 4) * 8);
                la_mma_bf16(o[2 * np], pa, vb[0], vb[1]);
                la_mma_bf16(o[2 * np + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();
    }

    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const int col = h * LA_D + n * 8 + 2 * t;
        if (r_lo < S) *reinterpret_cast<uint32_t *>(out + (int64_t)(s0 + r_lo) * ld_out + col) = la_pack_bf16(o[n][0] * inv0, o[n][1] * inv0);
        if (r_hi < S) *reinterpret_cast<uint32_t *>(out + (int64_t)(s0 + r_hi) * ld_out + col) = la_pack_bf16(o[n][2] * inv1, o[n][3] * inv1);
    }
}

int llm_attn_prefill(cudaStream_t st, const void *qkv, int ld_qkv, const void *kc, const void *vc, const int32_t *cu_seqlens,
                     const int32_t *slots, const int32_t *page_table, int pages_per_seq, void *out, int ld_out, int n_seq,
                     int max_seqlen, int hq_r, int kvh_r, float scale)
{
    if (n_seq <= 0 || max_seqlen <= 0) return 0;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, []() {
        attr_err = cudaFuncSetAttribute(llm_attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LA_PREFILL_SMEM);
    });
    if (attr_err != cudaSuccess) return fail_cuda(attr_err, "cudaFuncSetAttribute(llm prefill attention)");
    dim3 grid((max_seqlen + LA_BQ - 1) / LA_BQ, hq_r, n_seq);
    llm_attn_prefill_kernel<<<grid, 128, LA_PREFILL_SMEM, st>>>(
        static_cast<const __nv_bfloat16 *>(qkv), ld_qkv, static_cast<const __nv_bfloat16 *>(kc),
        static_cast<const __nv_bfloat16 *>(vc), cu_seqlens, slots, page_table, pages_per_seq, static_cast<__nv_bfloat16 *>(out),
        ld_out, hq_r / kvh_r, kvh_r, scale * 1.4426950408889634f);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ decode
// One CTA = (sequence, kv head), 4 warps.  Fused prologue: the q heads of the group and the new token's k / v
// are taken straight from the fp32 accumulator of the QKV projection (cleared here), RoPE is applied, k / v are
// appended to the cache.  The G query heads form the rows of ONE 16-row MMA tile (rows >= G are zero); cached
// keys stream through a 3-stage cp.async ring in blocks of 64, warp w owning keys 16w..16w+15 of every block
// (S = QK^T: 16 mma.sync, O += PV: 16 mma.sync per warp and block), and the four per-warp online-softmax
// states are merged through shared memory at the end.
constexpr int LDM_STAGES = 3;
constexpr int LDM_TILE = LA_BK * LA_LD;                                      // elements of one K or V tile
constexpr int LDM_SMEM = (16 * LA_LD + LDM_STAGES * 2 * LDM_TILE) * 2;       // Q + ring of (K, V)

__global__ void __launch_bounds__(128)
llm_attn_decode_kernel(float *__restrict__ ws_qkv, __nv_bfloat16 *__restrict__ kc, __nv_bfloat16 *__restrict__ vc,
                       const int32_t *__restrict__ ctx_len, const int32_t *__restrict__ slots,
                       const int32_t *__restrict__ page_table, int pages_per_seq,
                       const float *__restrict__ rope_cos, const float *__restrict__ rope_sin,
                       __nv_bfloat16 *__restrict__ out, int ld_out, int hq_r, int kvh_r, int max_ctx, float scale_log2e)
{
    extern __shared__ __align__(16) unsigned char la_smem[];
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(la_smem);
    __nv_bfloat16 *ring = Qs + 16 * LA_LD;

    asm volatile("griddepcontrol.wait;\n" ::: "memory");                 // QKV projection complete
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");    // the O projection may prefetch its weights
    const int b = blockIdx.x, kvh = blockIdx.y;
    const int G = hq_r / kvh_r;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int QKV = (hq_r + 2 * kvh_r) * LA_D;
    int pos = __ldg(ctx_len + b);
    pos = pos < max_ctx ? pos : max_ctx - 1;
    const int n_ctx = pos + 1;
    const int slot = __ldg(slots + b);
    const int32_t *pages = page_table + (int64_t)slot * pages_per_seq;
    // the row appended by this step: page of `pos`, row pos % 64
    const int64_t app = (((int64_t)__ldg(pages + (pos >> 6)) * kvh_r + kvh) * LA_BK + (pos & 63)) * LA_D;
    __nv_bfloat16 *Kapp = kc + app, *Vapp = vc + app;

    const int n_blocks = (n_ctx + LA_BK - 1) / LA_BK;
    auto issue_block = [&](int blk) {
        if (blk < n_blocks) {
            const int k0 = blk * LA_BK;
            __nv_bfloat16 *dst = ring + (blk % LDM_STAGES) * 2 * LDM_TILE;
            const int64_t off = ((int64_t)__ldg(pages + blk) * kvh_r + kvh) * (LA_BK * LA_D);   // one page = one key block
            la_load_tile_async(dst, kc + off, LA_D, min(LA_BK, n_ctx - k0), tid);
            la_load_tile_async(dst + LDM_TILE, vc + off, LA_D, min(LA_BK, n_ctx - k0), tid);
        }
        la_commit();   // (possibly empty) group: keeps the wait_group arithmetic uniform
    };
    // Blocks 0 and 1 hold only OLD cache rows unless the row appended below falls into them: request them before
    // the RoPE prologue so its three dependent round trips overlap the first HBM fetches.
    const bool early = pos >= 2 * LA_BK;
    if (early) { issue_block(0); issue_block(1); }

    // ---- prologue: RoPE on q (-> Qs) and k (-> cache), v -> cache, accumulator cleared
    float *row = ws_qkv + (int64_t)b * QKV;
    for (int idx = tid; idx < 16 * 64; idx += 128) {
        const int r = idx >> 6, i = idx & 63;
        __nv_bfloat16 o1 = __float2bfloat16_rn(0.f), o2 = o1;
        if (r < G) {
            float *src = row + (kvh * G + r) * LA_D;
            const float x1 = src[i], x2 = src[i + 64];
            src[i] = 0.f;
            src[i + 64] = 0.f;
            const float c = __ldg(rope_cos + (int64_t)pos * 64 + i), sv = __ldg(rope_sin + (int64_t)pos * 64 + i);
            o1 = __float2bfloat16_rn(x1 * c - x2 * sv);
            o2 = __float2bfloat16_rn(x2 * c + x1 * sv);
        }
        Qs[r * LA_LD + i] = o1;
        Qs[r * LA_LD + i + 64] = o2;
    }
    if (tid < 64) {
        float *src = row + (hq_r + kvh) * LA_D;
        const float x1 = src[tid], x2 = src[tid + 64];
        src[tid] = 0.f;
        src[tid + 64] = 0.f;
        const float c = __ldg(rope_cos + (int64_t)pos * 64 + tid), sv = __ldg(rope_sin + (int64_t)pos * 64 + tid);
        Kapp[tid] = __float2bfloat16_rn(x1 * c - x2 * sv);
        Kapp[tid + 64] = __float2bfloat16_rn(x2 * c + x1 * sv);
    } else {
        const int i = tid - 64;
        float *src = row + (hq_r + kvh_r + kvh) * LA_D;
        const float x1 = src[i], x2 = src[i + 64];
        src[i] = 0.f;
        src[i + 64] = 0.f;
        Vapp[i] = __float2bfloat16_rn(x1);
        Vapp[i + 64] = __float2bfloat16_rn(x2);
    }
    __syncthreads();   // Qs complete; the appended K/V row is ordered before this CTA's tile loads

    if (!early) { issue_block(0); issue_block(1); }

    uint32_t qa[8][4];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
        la_ldmatrix_x4(qa[kk], Qs + (lane & 15) * LA_LD + kk * 16 + (lane >> 4) * 8);
    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int blk = 0; blk < n_blocks; ++blk) {
        issue_block(blk + 2);
        la_wait<2>();
        __syncthreads();
        const __nv_bfloat16 *Ks = ring + (blk % LDM_STAGES) * 2 * LDM_TILE, *Vs = Ks + LDM_TILE;
        const int key0 = blk * LA_BK + warp * 16;   // this warp's 16 keys
        if (key0 < n_ctx) {
            float s[2][4];
            s[0][0] = s[0][1] = s[0][2] = s[0][3] = s[1][0] = s[1][1] = s[1][2] = s[1][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                uint32_t kb[4];
                la_ldmatrix_x4(kb, Ks + (warp * 16 + (lane & 7) + (lane >> 4) * 8) * LA_LD + kk * 16 + ((lane >> 3) & 1) * 8);
                la_mma_bf16(s[0], qa[kk], kb[0], kb[1]);
                la_mma_bf16(s[1], qa[kk], kb[2], kb[3]);
            }
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int key = key0 + n * 8 + 2 * t;
                s[n][0] = key < n_ctx ? s[n][0] * scale_log2e : -INFINITY;
                s[n][1] = key + 1 < n_ctx ? s[n][1] * scale_log2e : -INFINITY;
                s[n][2] = key < n_ctx ? s[n][2] * scale_log2e : -INFINITY;
                s[n][3] = key + 1 < n_ctx ? s[n][3] * scale_log2e : -INFINITY;
                mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
            }
            float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
                const float m_new = fmaxf(m_run[r], mx[r]);   // finite: key0 < n_ctx
                corr[r] = exp2f(m_run[r] - m_new);
                m_run[r] = m_new;
            }
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                s[n][0] = exp2f(s[n][0] - m_run[0]);
                s[n][1] = exp2f(s[n][1] - m_run[0]);
                s[n][2] = exp2f(s[n][2] - m_run[1]);
                s[n][3] = exp2f(s[n][3] - m_run[1]);
                rs[0] += s[n][0] + s[n][1];
                rs[1] += s[n][2] + s[n][3];
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
                rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
                l_run[r] = l_run[r] * corr[r] + rs[r];
            }
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                o[n][0] *= corr[0]; o[n][1] *= corr[0];
                o[n][2] *= corr[1]; o[n][3] *= corr[1];
            }
            uint32_t pa[4];
            pa[0] = la_pack_bf16(s[0][0], s[0][1]);
            pa[1] = la_pack_bf16(s[0][2], s[0][3]);
            pa[2] = la_pack_bf16(s[1][0], s[1][1]);
            pa[3] = la_pack_bf16(s[1][2], s[1][3]);
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                uint32_t vb[4];
                la_ldmatrix_x4_trans(vb, Vs + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LA_LD + np * 16 + (lane >> 4) * 8);
                la_mma_bf16(o[2 * np], pa, vb[0], vb[1]);
                la_mma_bf16(o[2 * np + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();   // all warps done with this stage before it is refilled
    }

    // ---- merge the four per-warp states (only rows g < G are real)
    float *sm_m = reinterpret_cast<float *>(ring);        // [4][8]
    float *sm_l = sm_m + 32;                              // [4][8]
    float *sm_o = sm_l + 32;                              // [4][8][128]
    if (t == 0) { sm_m[warp * 8 + g] = m_run[0]; sm_l[warp * 8 + g] = l_run[0]; }
#pragma unroll
    for (int n = 0; n < 16; ++n)
        *reinterpret_cast<float2 *>(sm_o + ((warp * 8 + g) * LA_D + n * 8 + 2 * t)) = make_float2(o[n][0], o[n][1]);
    __syncthreads();
    for (int idx = tid; idx < G * LA_D; idx += 128) {
        const int r = idx >> 7, d = idx & 127;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, sm_m[w * 8 + r]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float mw = sm_m[w * 8 + r];
            const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
            num = fmaf(wgt, sm_o[(w * 8 + r) * LA_D + d], num);
            den = fmaf(wgt, sm_l[w * 8 + r], den);
        }
        out[(int64_t)b * ld_out + (kvh * G + r) * LA_D + d] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    }
}

// decode attention fused with RoPE + KV append: reads (and clears) the fp32 QKV accumulator [32, (hq+2hkv)*128]
int llm_attn_decode(cudaStream_t st, float *ws_qkv, void *kc, void *vc, const int32_t *ctx_len, const int32_t *slots,
                    const int32_t *page_table, int pages_per_seq, const float *rope_cos, const float *rope_sin, void *out,
                    int ld_out, int n_seq, int hq_r, int kvh_r, int max_ctx, float scale)
{
    if (n_seq <= 0) return 0;
    const int G = hq_r / kvh_r;
    if (G < 1 || G > 8 || G * kvh_r != hq_r) return fail(B2S_ERR_INVALID, "llm attention: query group size %d not supported (1..8)", G);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, []() {
        attr_err = cudaFuncSetAttribute(llm_attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LDM_SMEM);
    });
    if (attr_err != cudaSuccess) return fail_cuda(attr_err, "cudaFuncSetAttribute(llm decode attention)");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_seq, (unsigned)kvh_r);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = LDM_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2S_CUDA(cudaLaunchKernelEx(&cfg, llm_attn_decode_kernel, ws_qkv, static_cast<__nv_bfloat16 *>(kc), static_cast<__nv_bfloat16 *>(vc),
                                ctx_len, slots, page_table, pages_per_seq, rope_cos, rope_sin, static_cast<__nv_bfloat16 *>(out), ld_out,
                                hq_r, kvh_r, max_ctx, scale * 1.4426950408889634f));
    count_launch();
    return 0;
}

}  // namespace b2s
