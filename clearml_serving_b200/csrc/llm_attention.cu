// llm_attention.cu -- causal grouped-query attention over the KV cache for the decoder-only LLM endpoint
// (BASELINE.json configs[4]; the reference delegates it to vLLM's paged attention,
// clearml_serving/serving/preprocess_service.py:1097-1348).  head_dim = 128, bf16 in, fp32 softmax/accumulate.
//
// KV cache layout (one layer): K and V each [slot][kv_head][max_ctx][128] bf16 -- one (sequence, kv head) is a
// contiguous [ctx, 128] matrix, so both kernels stream it with full 256-byte rows.
//
//  * prefill: one CTA = (64-query tile, q head, sequence), 4 warps x 16 query rows; K/V blocks of 64 keys are
//    double-buffered through shared memory with cp.async; S = QK^T and O += PV on mma.sync m16n8k16 (bf16) with
//    the online softmax in registers; only key blocks at or below the diagonal are visited.
//    FLOPs: 4 * 128 * S^2 / 2 per (sequence, q head).  ~1 % of the prefill FLOPs of Llama-3-8B at S=512, so
//    it stays on the legacy tensor path; the tcgen05 budget is in the GEMMs.
//  * decode: one CTA = (sequence, kv head), 8 warps; warp w walks key batches w, w+8, ... straight from global
//    memory (each lane owns 4 of the 128 dims: one coalesced 256-byte row per load), scores of a batch of
//    32/G keys x G q-heads are reduced with a 31-shuffle transpose-reduction, and the 8 partial softmax states
//    are merged through shared memory.  HBM-bound: 512 bytes per cached token per kv head.
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
                        const int32_t *__restrict__ slots, __nv_bfloat16 *__restrict__ out, int ld_out, int group,
                        int kvh_r, int max_ctx, float scale_log2e)
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
    const __nv_bfloat16 *Kg = kc + ((int64_t)slot * kvh_r + kvh) * max_ctx * LA_D;
    const __nv_bfloat16 *Vg = vc + ((int64_t)slot * kvh_r + kvh) * max_ctx * LA_D;
    const int k_end = min(S, q0 + LA_BQ);                // keys [0, k_end) can be visible to this tile

    la_load_tile_async(Qs, qkv + (int64_t)(s0 + q0) * ld_qkv + h * LA_D, ld_qkv, min(LA_BQ, S - q0), tid);
    auto issue_block = [&](int k0, int buf) {
        const int kv_valid = min(LA_BK, k_end - k0);
        la_load_tile_async(Ks2 + buf * LA_BK * LA_LD, Kg + (int64_t)k0 * LA_D, LA_D, kv_valid, tid);
        la_load_tile_async(Vs2 + buf * LA_BK * LA_LD, Vg + (int64_t)k0 * LA_D, LA_D, kv_valid, tid);
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
                     const int32_t *slots, void *out, int ld_out, int n_seq, int max_seqlen, int hq_r, int kvh_r,
                     int max_ctx, float scale)
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
        static_cast<const __nv_bfloat16 *>(vc), cu_seqlens, slots, static_cast<__nv_bfloat16 *>(out), ld_out, hq_r / kvh_r,
        kvh_r, max_ctx, scale * 1.4426950408889634f);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ decode
constexpr int LD_WARPS = 8;

template <int G>
__global__ void __launch_bounds__(LD_WARPS * 32)
llm_attn_decode_kernel(const __nv_bfloat16 *__restrict__ q, int ld_q, const __nv_bfloat16 *__restrict__ kc,
                       const __nv_bfloat16 *__restrict__ vc, const int32_t *__restrict__ ctx_len,
                       const int32_t *__restrict__ slots, __nv_bfloat16 *__restrict__ out, int ld_out, int kvh_r,
                       int max_ctx, float scale_log2e)
{
    constexpr int KB = 32 / G;   // keys per batch: KB * G (key, head) pairs = one per lane after the reduction
    __shared__ float sm_m[LD_WARPS][G], sm_l[LD_WARPS][G];
    __shared__ __align__(16) float sm_o[LD_WARPS][G][LA_D];

    const int b = blockIdx.x, kvh = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_ctx = min(__ldg(ctx_len + b) + 1, max_ctx);   // cached tokens + the one appended this step
    const int slot = __ldg(slots + b);
    const __nv_bfloat16 *Kg = kc + ((int64_t)slot * kvh_r + kvh) * max_ctx * LA_D + lane * 4;
    const __nv_bfloat16 *Vg = vc + ((int64_t)slot * kvh_r + kvh) * max_ctx * LA_D + lane * 4;

    float qr[G][4];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const uint2 u = *reinterpret_cast<const uint2 *>(q + (int64_t)b * ld_q + (kvh * G + gi) * LA_D + lane * 4);
        const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&u);
        const float2 a = __bfloat1622float2(p[0]), c = __bfloat1622float2(p[1]);
        qr[gi][0] = a.x * scale_log2e; qr[gi][1] = a.y * scale_log2e;
        qr[gi][2] = c.x * scale_log2e; qr[gi][3] = c.y * scale_log2e;
    }
    float o[G][4];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) o[gi][0] = o[gi][1] = o[gi][2] = o[gi][3] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;   // softmax state of head (lane % G), replicated over the lanes sharing it
    const int my_key = lane / G;

    for (int k0 = warp * KB; k0 < n_ctx; k0 += LD_WARPS * KB) {
        float part[32];
        uint2 vraw[KB];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const int key = min(k0 + kb, n_ctx - 1);
            const uint2 u = __ldg(reinterpret_cast<const uint2 *>(Kg + (int64_t)key * LA_D));
            vraw[kb] = __ldg(reinterpret_cast<const uint2 *>(Vg + (int64_t)key * LA_D));
            const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&u);
            const float2 a = __bfloat1622float2(p[0]), c = __bfloat1622float2(p[1]);
#pragma unroll
            for (int gi = 0; gi < G; ++gi)
                part[kb * G + gi] = fmaf(qr[gi][0], a.x, fmaf(qr[gi][1], a.y, fmaf(qr[gi][2], c.x, qr[gi][3] * c.y)));
        }
        // transpose-reduce: 32 partials on 32 lanes -> lane j holds the full dot product of pair j
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
            const bool upper = (lane & w) != 0;
#pragma unroll
            for (int i = 0; i < w; ++i) {
                const float send = upper ? part[i] : part[i + w];
                const float keep = upper ? part[i + w] : part[i];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
            }
        }
        float sc = (k0 + my_key < n_ctx) ? part[0] : -INFINITY;
        // per-head max / sum over the KB keys of the batch (lanes with equal lane % G)
        float mx = sc;
#pragma unroll
        for (int off = G; off < 32; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float m_new = fmaxf(m_run, mx);            // finite: key k0 of the batch is always valid
        const float corr = exp2f(m_run - m_new);         // 0 on the first batch (m_run = -inf)
        const float p = exp2f(sc - m_new);
        float ps = p;
#pragma unroll
        for (int off = G; off < 32; off <<= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
        l_run = l_run * corr + ps;
        m_run = m_new;
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            const float cg = __shfl_sync(0xffffffffu, corr, gi);
            o[gi][0] *= cg; o[gi][1] *= cg; o[gi][2] *= cg; o[gi][3] *= cg;
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const __nv_bfloat162 *pv = reinterpret_cast<const __nv_bfloat162 *>(&vraw[kb]);
            const float2 a = __bfloat1622float2(pv[0]), c = __bfloat1622float2(pv[1]);
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                const float pg = __shfl_sync(0xffffffffu, p, kb * G + gi);
                o[gi][0] = fmaf(pg, a.x, o[gi][0]); o[gi][1] = fmaf(pg, a.y, o[gi][1]);
                o[gi][2] = fmaf(pg, c.x, o[gi][2]); o[gi][3] = fmaf(pg, c.y, o[gi][3]);
            }
        }
    }
    if (lane < G) { sm_m[warp][lane] = m_run; sm_l[warp][lane] = l_run; }
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
        *reinterpret_cast<float4 *>(&sm_o[warp][gi][lane * 4]) = make_float4(o[gi][0], o[gi][1], o[gi][2], o[gi][3]);
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * LA_D; idx += LD_WARPS * 32) {
        const int gi = idx / LA_D, d = idx - gi * LA_D;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < LD_WARPS; ++w) M = fmaxf(M, sm_m[w][gi]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < LD_WARPS; ++w) {
            const float mw = sm_m[w][gi];
            const float wgt = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
            num = fmaf(wgt, sm_o[w][gi][d], num);
            den = fmaf(wgt, sm_l[w][gi], den);
        }
        out[(int64_t)b * ld_out + (kvh * G + gi) * LA_D + d] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    }
}

int llm_attn_decode(cudaStream_t st, const void *q, int ld_q, const void *kc, const void *vc, const int32_t *ctx_len,
                    const int32_t *slots, void *out, int ld_out, int n_seq, int hq_r, int kvh_r, int max_ctx, float scale)
{
    if (n_seq <= 0) return 0;
    const int G = hq_r / kvh_r;
    dim3 grid(n_seq, kvh_r);
    const float sl = scale * 1.4426950408889634f;
#define B2S_LAUNCH_DEC(GG)                                                                                             \
    llm_attn_decode_kernel<GG><<<grid, LD_WARPS * 32, 0, st>>>(static_cast<const __nv_bfloat16 *>(q), ld_q,            \
        static_cast<const __nv_bfloat16 *>(kc), static_cast<const __nv_bfloat16 *>(vc), ctx_len, slots,               \
        static_cast<__nv_bfloat16 *>(out), ld_out, kvh_r, max_ctx, sl)
    switch (G) {
    case 1: B2S_LAUNCH_DEC(1); break;
    case 2: B2S_LAUNCH_DEC(2); break;
    case 4: B2S_LAUNCH_DEC(4); break;
    case 8: B2S_LAUNCH_DEC(8); break;
    default: return fail(B2S_ERR_INVALID, "llm attention: query group size %d not supported (1, 2, 4, 8)", G);
    }
#undef B2S_LAUNCH_DEC
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2s
