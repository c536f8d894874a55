// conv.cu -- data-movement kernels around the convolutions of a ResNet-class network (BASELINE.json configs[2]); the
// convolutions themselves run on the tcgen05 GEMM kernels of gemm.cu (K6 of SURVEY.md 2.2):
//   * activations are NHWC fp16 with channels padded to a multiple of 8 (16-byte pixels), so a 1x1 stride-1
//     convolution IS `gemm_tn` on the activation matrix [N*H*W, C] with the weight [Cout, C];
//   * 3x3 / strided convolutions are implicit GEMMs: im2col-mode TMA gathers the A tiles from the NHWC tensor
//     (gemm.cu: make_tmap_im2col_nhwc), BatchNorm folded into weight/bias, bias + ReLU (+ residual) in the epilogue;
//   * the 7x7 stride-2 stem over the request pixels becomes a 4x4 stride-1 convolution over the 2x2 space-to-depth
//     image written by nchw_to_s2d_kernel here (gemm.cu: make_tmap_stem_s2d);
//   * im2col_nhwc_kernel materialises patch matrices for the comparison form only (pack_resnet(implicit_conv=False),
//     convolutions the TMA forms do not cover: input channels not a multiple of 64 outside the stem);
//   * max-pool / global average pool are plain coalesced NHWC kernels.
// This is what tritonserver's libtorch backend does with cuDNN for the reference's pytorch endpoint
// (clearml_serving/engines/triton/triton_helper.py:166-168,382-383; examples/pytorch).
#include "common.cuh"

#include <cuda_fp16.h>

namespace b2s {

// in: NCHW fp32 (in_dtype 0) or uint8 (in_dtype 4); out: NHWC fp16 with Cp >= C channels (zero padded)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const void *__restrict__ in, int in_dtype, int64_t n_pix_total, int C, int HW, int Cp,
                    __half *__restrict__ out)
{
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // n*HW + hw
    if (pix >= n_pix_total) return;
    const int64_t n = pix / HW, hw = pix - n * HW;
    __half *o = out + pix * Cp;
    for (int c0 = 0; c0 < Cp; c0 += 8) {
        __half v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            float f = 0.f;
            if (c < C) {
                const int64_t idx = (n * C + c) * HW + hw;
                f = in_dtype == B2S_U8 ? (float)static_cast<const uint8_t *>(in)[idx] : static_cast<const float *>(in)[idx];
            }
            v[j] = __float2half_rn(f);
        }
        *reinterpret_cast<uint4 *>(o + c0) = *reinterpret_cast<const uint4 *>(v);
    }
}

// Stem operand (see make_tmap_stem_s2d in gemm.cu): the 7x7 stride-2 pad-3 convolution over C <= 4 channels becomes a
// 4x4 stride-1 convolution over z[n, Hz, Wz, 16], z[n, hz, wz, (dy*2 + dx)*4 + c] = x[n, c, 2hz + dy - 3, 2wz + dx - 3]
// (0 outside the image and for c >= C).  in: NCHW fp32 / uint8 request pixels; one thread writes one 32-byte z pixel.
__global__ void __launch_bounds__(256)
nchw_to_s2d_kernel(const void *__restrict__ in, int in_dtype, int64_t n_zpix_total, int C, int H, int W, int Hz, int Wz,
                   __half *__restrict__ out)
{
    const int64_t zp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (n*Hz + hz)*Wz + wz
    if (zp >= n_zpix_total) return;
    const int wz = (int)(zp % Wz);
    const int64_t t = zp / Wz;
    const int hz = (int)(t % Hz);
    const int64_t n = t / Hz;
    __align__(16) __half v[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int ih = 2 * hz + dy - 3;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int iw = 2 * wz + dx - 3;
            const bool inside = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float f = 0.f;
                if (inside && c < C) {
                    const int64_t idx = ((n * C + c) * H + ih) * (int64_t)W + iw;
                    f = in_dtype == B2S_U8 ? (float)static_cast<const uint8_t *>(in)[idx] : static_cast<const float *>(in)[idx];
                }
                v[(dy * 2 + dx) * 4 + c] = __float2half_rn(f);
            }
        }
    }
    uint4 *o = reinterpret_cast<uint4 *>(out + zp * 16);
    o[0] = *reinterpret_cast<const uint4 *>(v);
    o[1] = *reinterpret_cast<const uint4 *>(v + 8);
}

// out[row, (kh*KW + kw)*C + c] = in[n, oh*s - pad + kh, ow*s - pad + kw, c]  (0 outside), row = (n, oh, ow);
// columns K..Kp-1 are zero.  One thread moves one 16-byte (8-channel) chunk.
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const __half *__restrict__ in, int64_t n_rows, int H, int W, int C, int KH, int KW, int stride, int pad,
                   int OH, int OW, int Kp, __half *__restrict__ out)
{
    const int chunks = Kp >> 3;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * chunks) return;
    const int64_t row = idx / chunks;
    const int chunk = (int)(idx - row * chunks);
    const int k = chunk << 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k < KH * KW * C) {
        const int tap = k / C, c = k - tap * C;
        const int kh = tap / KW, kw = tap - kh * KW;
        const int64_t n = row / ((int64_t)OH * OW);
        const int rem = (int)(row - n * (int64_t)OH * OW);
        const int oh = rem / OW, ow = rem - oh * OW;
        const int ih = oh * stride - pad + kh, iw = ow * stride - pad + kw;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W)
            v = *reinterpret_cast<const uint4 *>(in + ((n * H + ih) * (int64_t)W + iw) * C + c);
    }
    *reinterpret_cast<uint4 *>(out + row * Kp + k) = v;
}

// 3x3 stride-2 pad-1 max pooling, NHWC fp16, 8 channels per thread
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __half *__restrict__ in, int64_t n_out_pix, int H, int W, int C, int OH, int OW,
                    __half *__restrict__ out)
{
    const int cg = C >> 3;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out_pix * cg) return;
    const int64_t pix = idx / cg;
    const int c = (int)(idx - pix * cg) << 3;
    const int64_t n = pix / ((int64_t)OH * OW);
    const int rem = (int)(pix - n * (int64_t)OH * OW);
    const int oh = rem / OW, ow = rem - oh * OW;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * 2 - 1 + kh;
        if (ih < 0 || ih >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = ow * 2 - 1 + kw;
            if (iw < 0 || iw >= W) continue;
            const uint4 u = *reinterpret_cast<const uint4 *>(in + ((n * H + ih) * (int64_t)W + iw) * C + c);
            const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                m[2 * j] = fmaxf(m[2 * j], f.x);
                m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y);
            }
        }
    }
    __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __float2half_rn(m[j]);
    *reinterpret_cast<uint4 *>(out + pix * C + c) = *reinterpret_cast<const uint4 *>(v);
}

// global average pool: [N, HW, C] fp16 -> [N, C] fp16 (fp32 accumulation in pixel order)
__global__ void __launch_bounds__(256)
avgpool_kernel(const __half *__restrict__ in, int64_t n_img, int HW, int C, __half *__restrict__ out)
{
    const int cg = C >> 3;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_img * cg) return;
    const int64_t n = idx / cg;
    const int c = (int)(idx - n * cg) << 3;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < HW; ++p) {
        const uint4 u = *reinterpret_cast<const uint4 *>(in + (n * HW + p) * (int64_t)C + c);
        const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            s[2 * j] += f.x;
            s[2 * j + 1] += f.y;
        }
    }
    __half v[8];
    const float inv = 1.0f / (float)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __float2half_rn(s[j] * inv);
    *reinterpret_cast<uint4 *>(out + n * C + c) = *reinterpret_cast<const uint4 *>(v);
}

static unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

int nchw_to_nhwc(cudaStream_t st, const void *in, int in_dtype, int64_t n_img, int C, int H, int W, int Cp, void *out)
{
    if (n_img <= 0) return 0;
    if (Cp % 8 != 0 || Cp < C) return fail(B2S_ERR_INVALID, "nchw_to_nhwc: bad channel padding");
    if (in_dtype != B2S_F32 && in_dtype != B2S_U8) return fail(B2S_ERR_INVALID, "nchw_to_nhwc: input must be float32 or uint8");
    const int64_t pix = n_img * H * W;
    nchw_to_nhwc_kernel<<<blocks_for(pix), 256, 0, st>>>(in, in_dtype, pix, C, H * W, Cp, static_cast<__half *>(out));
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

int nchw_to_s2d(cudaStream_t st, const void *in, int in_dtype, int64_t n_img, int C, int H, int W, int Hz, int Wz, void *out)
{
    if (n_img <= 0) return 0;
    if (C < 1 || C > 4 || 2 * Hz < H + 6 || 2 * Wz < W + 6) return fail(B2S_ERR_INVALID, "stem: bad space-to-depth geometry");
    if (in_dtype != B2S_F32 && in_dtype != B2S_U8) return fail(B2S_ERR_INVALID, "stem: input must be float32 or uint8");
    const int64_t zpix = n_img * Hz * Wz;
    nchw_to_s2d_kernel<<<blocks_for(zpix), 256, 0, st>>>(in, in_dtype, zpix, C, H, W, Hz, Wz, static_cast<__half *>(out));
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

int im2col_nhwc(cudaStream_t st, const void *in, int64_t n_img, int H, int W, int C, int KH, int KW, int stride, int pad,
                int OH, int OW, int Kp, void *out)
{
    if (n_img <= 0) return 0;
    if (C % 8 != 0 || Kp % 8 != 0 || Kp < KH * KW * C) return fail(B2S_ERR_INVALID, "im2col: channels / K padding must be multiples of 8");
    const int64_t rows = n_img * OH * OW;
    im2col_nhwc_kernel<<<blocks_for(rows * (Kp / 8)), 256, 0, st>>>(static_cast<const __half *>(in), rows, H, W, C, KH, KW,
                                                                  stride, pad, OH, OW, Kp, static_cast<__half *>(out));
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

int maxpool3x3s2(cudaStream_t st, const void *in, int64_t n_img, int H, int W, int C, int OH, int OW, void *out)
{
    if (n_img <= 0) return 0;
    if (C % 8 != 0) return fail(B2S_ERR_INVALID, "maxpool: C must be a multiple of 8");
    const int64_t pix = n_img * OH * OW;
    maxpool3x3s2_kernel<<<blocks_for(pix * (C / 8)), 256, 0, st>>>(static_cast<const __half *>(in), pix, H, W, C, OH, OW,
                                                                  static_cast<__half *>(out));
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

int avgpool(cudaStream_t st, const void *in, int64_t n_img, int HW, int C, void *out)
{
    if (n_img <= 0) return 0;
    if (C % 8 != 0) return fail(B2S_ERR_INVALID, "avgpool: C must be a multiple of 8");
    avgpool_kernel<<<blocks_for(n_img * (C / 8)), 256, 0, st>>>(static_cast<const __half *>(in), n_img, HW, C,
                                                               static_cast<__half *>(out));
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2s
