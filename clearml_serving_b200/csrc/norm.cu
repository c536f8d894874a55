// norm.cu -- bandwidth-bound transformer glue kernels (K7 LayerNorm, K10 embedding gather, CLS
// gather) for B2S_MODEL_GRAPH models.  One warp per row, 128-bit coalesced accesses, fp32 math,
// warp-shuffle reductions, the row cached in registers between the statistics and the write pass
// (each element is read from HBM once and written once).
//
// Numerics follow torch.nn.LayerNorm (the op tritonserver's libtorch backend executes for the
// reference's transformer endpoint, examples/huggingface): mean, then biased variance of the
// centred values, rsqrt(var + eps), affine.  The residual stream is kept in fp32 (out32) next to
// the fp16 copy that feeds the next tensor-core GEMM (out16), so rounding does not accumulate
// across the 24 LayerNorms of BERT-base (parity bar: 1e-3 relative on the logits).
// Algorithmic bytes per row: LayerNorm 4H (fp32 in) + 2H + 4H (out); embedding 3*2H (gathers) + 6H.
#include "common.cuh"

#include <cuda_fp16.h>

namespace b2s {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// normalise NCH*4 values per lane held in registers and write fp16 + fp32 rows
template <int NCH>
__device__ __forceinline__ void ln_finish(float (&v)[NCH][4], int lane, int H, const float *__restrict__ gamma,
                                          const float *__restrict__ beta, float eps, __half *__restrict__ o16,
                                          float *__restrict__ o32)
{
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        if ((c * 32 + lane) * 4 < H) s += (v[c][0] + v[c][1]) + (v[c][2] + v[c][3]);
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        if ((c * 32 + lane) * 4 < H) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = v[c][j] - mean;
                q += d * d;
            }
        }
    const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < H) {
            const float4 g = *reinterpret_cast<const float4 *>(gamma + col);
            const float4 b = *reinterpret_cast<const float4 *>(beta + col);
            float4 y;
            y.x = (v[c][0] - mean) * rstd * g.x + b.x;
            y.y = (v[c][1] - mean) * rstd * g.y + b.y;
            y.z = (v[c][2] - mean) * rstd * g.z + b.z;
            y.w = (v[c][3] - mean) * rstd * g.w + b.w;
            if (o32) *reinterpret_cast<float4 *>(o32 + col) = y;
            if (o16) {
                __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t *>(&h0);
                u.y = *reinterpret_cast<uint32_t *>(&h1);
                *reinterpret_cast<uint2 *>(o16 + col) = u;
            }
        }
    }
}

// LayerNorm over the last dim of an fp32 [rows, H] matrix (H % 4 == 0, H <= NCH*128).
template <int NCH>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float *__restrict__ in, int64_t rows, int H, const float *__restrict__ gamma,
                 const float *__restrict__ beta, float eps, __half *__restrict__ out16, float *__restrict__ out32)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float *x = in + row * H;
    float v[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < H) {
            const float4 t = *reinterpret_cast<const float4 *>(x + col);
            v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
        } else {
            v[c][0] = v[c][1] = v[c][2] = v[c][3] = 0.f;
        }
    }
    ln_finish<NCH>(v, lane, H, gamma, beta, eps, out16 ? out16 + row * H : nullptr, out32 ? out32 + row * H : nullptr);
}

// BERT embeddings: word[id] + position[pos in sequence] + token_type[tt], then LayerNorm.
// Tokens are packed (ragged): cu_seqlens[b] is the first token of sequence b.
template <int NCH>
__global__ void __launch_bounds__(256)
embed_layernorm_kernel(const int32_t *__restrict__ ids, const int32_t *__restrict__ types,
                       const int64_t *__restrict__ cu_seqlens, int n_seq, int64_t n_tokens, int H,
                       const __half *__restrict__ word, const __half *__restrict__ pos, const __half *__restrict__ type,
                       int vocab, int max_pos, int n_types, const float *__restrict__ gamma,
                       const float *__restrict__ beta, float eps, __half *__restrict__ out16, float *__restrict__ out32)
{
    const int lane = threadIdx.x & 31;
    const int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= n_tokens) return;
    // position inside the sequence: binary search over cu_seqlens (n_seq <= a few hundred)
    int lo = 0, hi = n_seq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cu_seqlens + mid) <= tok) lo = mid; else hi = mid;
    }
    int p = (int)(tok - __ldg(cu_seqlens + lo));
    p = min(p, max_pos - 1);
    int id = __ldg(ids + tok);
    id = min(max(id, 0), vocab - 1);
    int tt = types ? __ldg(types + tok) : 0;
    tt = min(max(tt, 0), n_types - 1);
    const __half *w = word + (int64_t)id * H, *pe = pos + (int64_t)p * H, *te = type + (int64_t)tt * H;
    float v[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < H) {
            const uint2 a = *reinterpret_cast<const uint2 *>(w + col);
            const uint2 b = *reinterpret_cast<const uint2 *>(pe + col);
            const uint2 d = *reinterpret_cast<const uint2 *>(te + col);
            const float2 a0 = __half22float2(*reinterpret_cast<const __half2 *>(&a.x));
            const float2 a1 = __half22float2(*reinterpret_cast<const __half2 *>(&a.y));
            const float2 b0 = __half22float2(*reinterpret_cast<const __half2 *>(&b.x));
            const float2 b1 = __half22float2(*reinterpret_cast<const __half2 *>(&b.y));
            const float2 d0 = __half22float2(*reinterpret_cast<const __half2 *>(&d.x));
            const float2 d1 = __half22float2(*reinterpret_cast<const __half2 *>(&d.y));
            v[c][0] = (a0.x + d0.x) + b0.x;   // torch order: (word + token_type) + position
            v[c][1] = (a0.y + d0.y) + b0.y;
            v[c][2] = (a1.x + d1.x) + b1.x;
            v[c][3] = (a1.y + d1.y) + b1.y;
        } else {
            v[c][0] = v[c][1] = v[c][2] = v[c][3] = 0.f;
        }
    }
    ln_finish<NCH>(v, lane, H, gamma, beta, eps, out16 ? out16 + tok * H : nullptr, out32 ? out32 + tok * H : nullptr);
}

// out[b, :] = in[cu_seqlens[b], :]   (first token of every sequence: BERT pooler input)
__global__ void __launch_bounds__(256)
gather_first_kernel(const __half *__restrict__ in, const int64_t *__restrict__ cu_seqlens, int n_seq, int H,
                    __half *__restrict__ out)
{
    const int b = blockIdx.x;
    if (b >= n_seq) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(in + __ldg(cu_seqlens + b) * H);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (int64_t)b * H);
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
}

int layernorm(cudaStream_t st, const float *in, int64_t rows, int H, const float *gamma, const float *beta, float eps,
              void *out16, float *out32)
{
    if (rows <= 0) return 0;
    if (H % 4 != 0 || H > 4096) return fail(B2S_ERR_INVALID, "layernorm: H=%d must be a multiple of 4 and <= 4096", H);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    __half *o16 = static_cast<__half *>(out16);
    if (H <= 1024) layernorm_kernel<8><<<grid, 256, 0, st>>>(in, rows, H, gamma, beta, eps, o16, out32);
    else layernorm_kernel<32><<<grid, 256, 0, st>>>(in, rows, H, gamma, beta, eps, o16, out32);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

int embed_layernorm(cudaStream_t st, const int32_t *ids, const int32_t *types, const int64_t *cu_seqlens, int n_seq,
                    int64_t n_tokens, int H, const void *word, const void *pos, const void *type, int vocab, int max_pos,
                    int n_types, const float *gamma, const float *beta, float eps, void *out16, float *out32)
{
    if (n_tokens <= 0) return 0;
    if (H % 4 != 0 || H > 4096) return fail(B2S_ERR_INVALID, "embed_layernorm: H=%d must be a multiple of 4 and <= 4096", H);
    const unsigned grid = (unsigned)((n_tokens + 7) / 8);
    const __half *w = static_cast<const __half *>(word), *p = static_cast<const __half *>(pos),
                 *t = static_cast<const __half *>(type);
    __half *o16 = static_cast<__half *>(out16);
    if (H <= 1024)
        embed_layernorm_kernel<8><<<grid, 256, 0, st>>>(ids, types, cu_seqlens, n_seq, n_tokens, H, w, p, t, vocab, max_pos,
                                                        n_types, gamma, beta, eps, o16, out32);
    else
        embed_layernorm_kernel<32><<<grid, 256, 0, st>>>(ids, types, cu_seqlens, n_seq, n_tokens, H, w, p, t, vocab, max_pos,
                                                         n_types, gamma, beta, eps, o16, out32);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

int gather_first(cudaStream_t st, const void *in, const int64_t *cu_seqlens, int n_seq, int H, void *out)
{
    if (n_seq <= 0) return 0;
    if (H % 8 != 0) return fail(B2S_ERR_INVALID, "gather_first: H must be a multiple of 8");
    gather_first_kernel<<<n_seq, 128, 0, st>>>(static_cast<const __half *>(in), cu_seqlens, n_seq, H,
                                               static_cast<__half *>(out));
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2s

extern "C" B2S_API int b2s_op_layernorm(int device, void *cuda_stream, const float *in, int64_t rows, int H,
                                         const float *gamma, const float *beta, float eps, void *out16, float *out32)
{
    using namespace b2s;
    B2S_CUDA(cudaSetDevice(device));
    return layernorm(static_cast<cudaStream_t>(cuda_stream), in, rows, H, gamma, beta, eps, out16, out32);
}

extern "C" B2S_API int b2s_op_embed_layernorm(int device, void *cuda_stream, const int32_t *ids, const int32_t *types,
                                               const int64_t *cu_seqlens, int n_seq, int64_t n_tokens, int H,
                                               const void *word, const void *pos, const void *type, int vocab,
                                               int max_pos, int n_types, const float *gamma, const float *beta,
                                               float eps, void *out16, float *out32)
{
    using namespace b2s;
    B2S_CUDA(cudaSetDevice(device));
    return embed_layernorm(static_cast<cudaStream_t>(cuda_stream), ids, types, cu_seqlens, n_seq, n_tokens, H, word, pos,
                           type, vocab, max_pos, n_types, gamma, beta, eps, out16, out32);
}
