// attention.cu -- fused variable-length self-attention for encoder models (kernel K8, BERT form):
//     ctx[t, h*64:(h+1)*64] = softmax(Q_h K_h^T / sqrt(64) + key_mask) V_h      (non-causal)
// over PACKED tokens: qkv is [T, 3*H] fp16 (Q | K | V per token, H = heads*64), sequences are
// delimited by cu_seqlens, so a request's result can never depend on its batch-mates (SURVEY.md 5.9
// rule 4: exact masking).  This is the softmax/attention block tritonserver's backends run as
// three separate cuBLAS/cuDNN calls for the reference's transformer endpoint (examples/huggingface).
//
// One CTA = (64-query tile, head, sequence), 4 warps x 16 query rows.  K/V are streamed through
// shared memory in 64-key blocks; S = QK^T and O += PV run on the tensor cores (mma.sync m16n8k16,
// fp16 in / fp32 accumulate) with an online (flash-style) softmax in registers, so the S x S score
// matrix never exists in memory.  Attention is ~5 % of BERT-base FLOPs at S=256, the
// tcgen05 budget is spent on the GEMMs; the roofline of this kernel is the tensor pipe at HMMA rate.
// Algorithmic FLOPs: 4 * S^2 * 64 per (sequence, head).
#include "common.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

namespace b2s {

constexpr int ATT_D = 64;        // head dim
constexpr int ATT_BQ = 64;       // queries per CTA (4 warps x 16 rows)
constexpr int ATT_BK = 64;       // keys per smem block
constexpr int ATT_LD = ATT_D + 8;  // padded smem row (halfs): 144 B stride, conflict-free ldmatrix

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem_ptr)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_ptr)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// exp2 of a finite-or-minus-infinity argument: the raw MUFU form (exp2f adds a range fix-up per element)
__device__ __forceinline__ float ex2_raw(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi)
{
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

// copy a [64 rows x 64 halfs] tile (row stride ld_src elements) into padded smem with cp.async (16-byte pieces, zero-fill for
// rows >= valid): the copy of key block k+1 overlaps the MMAs of block k
__device__ __forceinline__ void load_tile_async(__half *dst, const __half *src, int64_t ld_src, int valid_rows, int tid)
{
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i >> 3, c = (i & 7) * 8;
        const bool ok = r < valid_rows;
        const __half *g = src + (int64_t)(ok ? r : 0) * ld_src + c;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + r * ATT_LD + c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(g), "r"(ok ? 16 : 0) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// 5 CTAs (20 warps) per SM: 96 registers, 37 KB of shared memory each.  ncu of the 4-CTA form showed the tensor pipe 26 %
// active with 22 % of the warp slots occupied: the kernel is bound by latency at low occupancy, not by the HMMA rate.
__global__ void __launch_bounds__(128, 5)
attention_varlen_kernel(const __half *__restrict__ qkv, const int64_t *__restrict__ cu_seqlens,
                        const int32_t *__restrict__ key_mask, __half *__restrict__ out, int heads, float scale_log2e, int skip_upto)
{
    __shared__ __align__(16) __half Ks2[2][ATT_BK * ATT_LD];   // double-buffered key / value blocks
    __shared__ __align__(16) __half Vs2[2][ATT_BK * ATT_LD];
    __shared__ float mask_bias2[2][ATT_BK];

    const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
    const int64_t s0 = __ldg(cu_seqlens + b);
    const int S = (int)(__ldg(cu_seqlens + b + 1) - s0);
    const int q0 = qt * ATT_BQ;
    if (q0 >= S || S <= skip_upto) return;     // sequences of <= skip_upto tokens belong to the tcgen05 form (attention_tc.cu)
    const int H = heads * ATT_D;
    const int64_t ld = 3 * (int64_t)H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    // Q fragments (A operand of m16n8k16) for this warp's 16 rows, 4 k-steps over d, straight from global memory: no
    // shared-memory tile, no barrier, and the loads fly together with the first K/V block's cp.async
    uint32_t qa[4][4];
    {
        const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
        const __half *p_lo = qkv + (s0 + r_lo) * ld + h * ATT_D + 2 * t, *p_hi = p_lo + 8 * ld;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            qa[kk][0] = r_lo < S ? __ldg(reinterpret_cast<const uint32_t *>(p_lo + kk * 16)) : 0u;
            qa[kk][1] = r_hi < S ? __ldg(reinterpret_cast<const uint32_t *>(p_hi + kk * 16)) : 0u;
            qa[kk][2] = r_lo < S ? __ldg(reinterpret_cast<const uint32_t *>(p_lo + kk * 16 + 8)) : 0u;
            qa[kk][3] = r_hi < S ? __ldg(reinterpret_cast<const uint32_t *>(p_hi + kk * 16 + 8)) : 0u;
        }
    }

    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    auto issue_block = [&](int k0, int buf) {
        const int kv_valid = min(ATT_BK, S - k0);
        load_tile_async(Ks2[buf], qkv + (s0 + k0) * ld + H + h * ATT_D, ld, kv_valid, tid);
        load_tile_async(Vs2[buf], qkv + (s0 + k0) * ld + 2 * H + h * ATT_D, ld, kv_valid, tid);
        if (tid < ATT_BK) {
            float mb = 0.f;
            if (tid >= kv_valid) mb = -INFINITY;
            else if (key_mask && __ldg(key_mask + s0 + k0 + tid) == 0) mb = -INFINITY;
            mask_bias2[buf][tid] = mb;
        }
        cp_async_commit_group();
    };
    issue_block(0, 0);
    int buf = 0;
    for (int k0 = 0; k0 < S; k0 += ATT_BK, buf ^= 1) {
        const bool more = k0 + ATT_BK < S;
        if (more) issue_block(k0 + ATT_BK, buf ^ 1);   // prefetch the next block into the other buffer
        if (more) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
        __syncthreads();
        const __half *Ks = Ks2[buf], *Vs = Vs2[buf];
        const float *mask_bias = mask_bias2[buf];

        // S = Q K^T for 16 rows x 64 keys
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
                uint32_t kb[4];
                ldmatrix_x4(kb, Ks + (np * 16 + (lane & 7) + (lane >> 4) * 8) * ATT_LD + kk * 16 + ((lane >> 3) & 1) * 8);
                mma_16816(s[2 * np], qa[kk], kb[0], kb[1]);
                mma_16816(s[2 * np + 1], qa[kk], kb[2], kb[3]);
            }
        }
        // scale (log2 domain), mask, online softmax
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float b0 = mask_bias[n * 8 + 2 * t], b1 = mask_bias[n * 8 + 2 * t + 1];
            s[n][0] = s[n][0] * scale_log2e + b0;
            s[n][1] = s[n][1] * scale_log2e + b1;
            s[n][2] = s[n][2] * scale_log2e + b0;
            s[n][3] = s[n][3] * scale_log2e + b1;
            mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
        }
        float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = (m_new == -INFINITY) ? 1.f : ex2_raw(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        const float m0 = (m_run[0] == -INFINITY) ? 0.f : m_run[0];
        const float m1 = (m_run[1] == -INFINITY) ? 0.f : m_run[1];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = ex2_raw(s[n][0] - m0);
            s[n][1] = ex2_raw(s[n][1] - m0);
            s[n][2] = ex2_raw(s[n][2] - m1);
            s[n][3] = ex2_raw(s[n][3] - m1);
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
        for (int n = 0; n < 8; ++n) {
            o[n][0] *= corr[0]; o[n][1] *= corr[0];
            o[n][2] *= corr[1]; o[n][3] *= corr[1];
        }
        // O += P V : P (16 x 64 keys) from the score registers, V^T fragments via ldmatrix.trans
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 16 keys per k-step
            uint32_t pa[4];
            pa[0] = pack_half2(s[2 * j][0], s[2 * j][1]);
            pa[1] = pack_half2(s[2 * j][2], s[2 * j][3]);
            pa[2] = pack_half2(s[2 * j + 1][0], s[2 * j + 1][1]);
            pa[3] = pack_half2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-wide d tiles
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, Vs + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LD + np * 16 + (lane >> 4) * 8);
                mma_16816(o[2 * np], pa, vb[0], vb[1]);
                mma_16816(o[2 * np + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();   // everyone is done with this buffer before the next prefetch overwrites it
    }

    // normalise and store: rows g and g+8 of this warp's 16-row slab
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int col = h * ATT_D + n * 8 + 2 * t;
        if (r_lo < S) *reinterpret_cast<uint32_t *>(out + (s0 + r_lo) * H + col) = pack_half2(o[n][0] * inv0, o[n][1] * inv0);
        if (r_hi < S) *reinterpret_cast<uint32_t *>(out + (s0 + r_hi) * H + col) = pack_half2(o[n][2] * inv1, o[n][3] * inv1);
    }
}

int attention_varlen_tc(cudaStream_t st, const void *qkv, const int64_t *cu_seqlens, const int32_t *key_mask, void *out, int n_seq,
                        int max_seqlen, int64_t total_tokens, int heads);

// total_tokens: rows of the packed qkv matrix (cu_seqlens[n_seq]), < 0 when the caller does not know it on the host
int attention_varlen(cudaStream_t st, const void *qkv, const int64_t *cu_seqlens, const int32_t *key_mask, void *out,
                     int n_seq, int max_seqlen, int heads, int head_dim, int64_t total_tokens)
{
    if (n_seq <= 0 || max_seqlen <= 0) return 0;
    if (head_dim != ATT_D) return fail(B2S_ERR_INVALID, "attention: head_dim %d not supported (64 only)", head_dim);
    // tcgen05 form (attention_tc.cu) whenever a sequence's scores fit tensor memory; B2S_ATTN_TC=0 selects the mma.sync form
    static const bool tc_on = []() { const char *e = getenv("B2S_ATTN_TC"); return !(e && e[0] == '0'); }();
    // WHICH form a sequence takes depends on its own length only -- never on its batch-mates (SURVEY.md 5.9 rule 4)
    int skip_upto = 0;
    if (tc_on && total_tokens > 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0) {
        B2S_TRY(attention_varlen_tc(st, qkv, cu_seqlens, key_mask, out, n_seq, max_seqlen, total_tokens, heads));
        if (max_seqlen <= 384) return 0;
        skip_upto = 384;
    }
    dim3 grid((max_seqlen + ATT_BQ - 1) / ATT_BQ, heads, n_seq);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)head_dim);
    attention_varlen_kernel<<<grid, 128, 0, st>>>(static_cast<const __half *>(qkv), cu_seqlens, key_mask,
                                                  static_cast<__half *>(out), heads, scale_log2e, skip_upto);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2s

extern "C" B2S_API int b2s_op_attention(int device, void *cuda_stream, const void *qkv, const int64_t *cu_seqlens,
                                         const int32_t *key_mask, void *out, int n_seq, int max_seqlen, int heads,
                                         int head_dim, int64_t total_tokens)
{
    using namespace b2s;
    B2S_CUDA(cudaSetDevice(device));
    int64_t total = total_tokens;     // the tensor map of the tcgen05 form needs the row count on the host
    if (total <= 0 && n_seq > 0 && cu_seqlens) {
        B2S_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
        B2S_CUDA(cudaMemcpy(&total, cu_seqlens + n_seq, 8, cudaMemcpyDeviceToHost));
    }
    return attention_varlen(static_cast<cudaStream_t>(cuda_stream), qkv, cu_seqlens, key_mask, out, n_seq, max_seqlen,
                            heads, head_dim, total);
}
