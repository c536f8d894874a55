// attention_tc.cu -- variable-length self-attention for encoder models on the 5th-generation tensor cores
// (kernel K8, tcgen05 form; attention.cu keeps the mma.sync form for sequences longer than 384 tokens).
//     ctx[t, h*64:(h+1)*64] = softmax(Q_h K_h^T / sqrt(64) + key_mask) V_h      (non-causal, packed tokens)
// This is the attention block tritonserver's backends run as cuBLAS / cuDNN calls for the reference's transformer
// endpoint (examples/huggingface; model placed by engines/triton/triton_helper.py:159-186).
//
// PERSISTENT kernel, one CTA per SM, 10 warps, walking the work items (sequence, head, 128-query tile) round-robin:
//   warp 8 (one lane)  TMA producer: Q tile + every key / value block of the sequence (<= 3 blocks of 128 keys; box
//                      128 tokens x 64 halfs of the packed [T, 3H] qkv matrix, 128-byte swizzle) into a 2-stage ring.
//   warp 9 (one lane)  MMA issuer:  S = Q K^T     tcgen05.mma 128 x N x 16 (N = keys of the block rounded up to 16),
//                                                 accumulators in TENSOR MEMORY, 128 columns per key block;
//                                   O = P V       tcgen05.mma 128 x 64 x 16 with A = P read FROM TENSOR MEMORY (written
//                                                 there by the softmax threads: the score matrix never touches shared
//                                                 memory) and B = V consumed as an MN-MAJOR operand exactly as TMA laid it
//                                                 down (no transpose).  S of item i+1 is issued before P of item i is
//                                                 awaited (two S buffers), so loads and MMAs hide behind the softmax.
//   warps 0-3 / 4-7    two softmax warpgroups, one per score buffer (even / odd items), so one group's exp2 stream
//                      covers the other's tensor-memory load latency; inside a group
//                      thread i owns query row i = tensor-memory lane i (thread-per-row softmax, no shuffles): with
//                      <= 384 keys the whole score row sits in tensor memory, so the softmax is EXACT two-pass (row
//                      maximum, then exp2 / sum) and the output accumulator is never rescaled; P (fp16 pairs) overwrites
//                      the score columns already consumed; O lands in the dead upper score columns; epilogue
//                      O / l -> fp16 -> one 128-byte row store.
// Keys outside the sequence (rows of the box that belong to the next sequence, or zero-filled past the last token) and
// masked keys get probability exactly 0, so a request's result cannot depend on its batch-mates.
// Algorithmic FLOPs: 4 * S^2 * 64 per (sequence, head); bound: tensor pipe (in practice the exp2 / issue rate of the
// softmax warps, see profiles/).
#include "common.cuh"
#include "sm100.cuh"

#include <cuda_fp16.h>

#include <mutex>

namespace b2s {

using namespace sm100;

int make_tmap_2d_kmajor(CUtensorMap *out, const void *base, int64_t rows, int64_t K, int64_t ld_elems, int box_rows,
                        int is_bf16);

constexpr int AT_BM = 128;                  // queries per work item
constexpr int AT_BN = 128;                  // keys per block
constexpr int AT_D = 64;                    // head dim
constexpr int AT_MAX_KB = 3;                // <= 384 keys: S fits tensor memory
constexpr int AT_TILE = AT_BM * AT_D * 2;   // 16 KB: one [128 x 64] fp16 tile
constexpr int AT_STAGES = 2;
constexpr int AT_THREADS = 320;              // 2 softmax warpgroups (warps 0-3, 4-7) + TMA producer (8) + MMA issuer (9)

// instruction descriptor, kind::f16, fp32 accumulate, fp16 operands; A K-major; B K-major (b_mn = 0) or MN-major (1)
__device__ __forceinline__ uint32_t at_idesc(int m, int n, int b_mn)
{
    return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// MN-major operand tile stored as [K rows][64 x 16-bit] with the 128-byte swizzle (what TMA writes for a box of 64
// contiguous elements): 8-row groups 1024 B apart (stride byte offset), one 64-element atom along MN (LBO unused)
__device__ __forceinline__ uint64_t at_desc_mnmajor(const void *smem_tile)
{
    const uint32_t addr = smem_u32(smem_tile);
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (P) is read from tensor memory (lane = row, 2 halfs per column)
__device__ __forceinline__ void at_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// warp-collective: 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void at_tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void at_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ float at_ex2(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t at_pack(float lo, float hi)
{
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void at_named_barrier_128(int id) { asm volatile("bar.sync %0, 128;\n" ::"r"(id) : "memory"); }

struct AtItem { int b, h, q0, S; int64_t s0; };

// the n-th valid work item of this CTA (every role walks the same list): items are (sequence, head, query tile) in that
// nesting order, tiles whose first query lies beyond the sequence -- and sequences longer than the tcgen05 form handles --
// are skipped
struct AtWalker {
    const int64_t *cu;
    int n_seq, heads, max_qt;
    int64_t next, total;
    int stride;
    __device__ AtWalker(const int64_t *cu_, int n_seq_, int heads_, int max_qt_)
        : cu(cu_), n_seq(n_seq_), heads(heads_), max_qt(max_qt_), next(blockIdx.x), total((int64_t)n_seq_ * heads_ * max_qt_), stride(gridDim.x) {}
    __device__ bool pop(AtItem &it)
    {
        while (next < total) {
            const int64_t i = next;
            next += stride;
            const int qt = (int)(i % max_qt);
            const int64_t r = i / max_qt;
            const int h = (int)(r % heads), b = (int)(r / heads);
            const int64_t s0 = __ldg(cu + b);
            const int S = (int)(__ldg(cu + b + 1) - s0);
            if (S > AT_MAX_KB * AT_BN || qt * AT_BM >= S) continue;
            it.b = b; it.h = h; it.q0 = qt * AT_BM; it.S = S; it.s0 = s0;
            return true;
        }
        return false;
    }
};

template <int NSB>   // score buffers in tensor memory: 2 when a sequence has <= 256 keys, else 1
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const int64_t *__restrict__ cu_seqlens,
                    const int32_t *__restrict__ key_mask, __half *__restrict__ out, int n_seq, int heads, int max_qt, int nkb_max,
                    float scale_log2e)
{
    extern __shared__ unsigned char at_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = heads * AT_D;
    const int stage_bytes = (1 + 2 * nkb_max) * AT_TILE;         // Q | K blocks | V blocks
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + AT_STAGES * stage_bytes);
    uint64_t *bar_full = bars, *bar_empty = bars + 2, *bar_s = bars + 4, *bar_p = bars + 6, *bar_o = bars + 8, *bar_sfree = bars + 10;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
    uint32_t *kbits = tmem_slot + 2;                             // [2 groups][2][12]: validity word of every 32-key chunk (masked path)
    const int sbuf_cols = nkb_max * AT_BN;                       // tensor-memory columns of one score buffer

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
            mbar_init(&bar_s[i], 1);
            mbar_init(&bar_p[i], 128);
            mbar_init(&bar_o[i], 1);
            mbar_init(&bar_sfree[i], 128);
        }
        fence_barrier_init();
        prefetch_tensormap(&tm_qkv);
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 8) {
        // =========================================================================== TMA producer
        if (lane == 0) {
            AtWalker w(cu_seqlens, n_seq, heads, max_qt);
            AtItem it;
            for (uint32_t n = 0; w.pop(it); ++n) {
                const int st = n & 1;
                const int nkb = (it.S + AT_BN - 1) / AT_BN;
                mbar_wait(&bar_empty[st], ((n >> 1) & 1) ^ 1);          // the MMAs that read this stage have retired
                unsigned char *Qs = smem + st * stage_bytes, *Ks = Qs + AT_TILE, *Vs = Ks + nkb_max * AT_TILE;
                mbar_arrive_expect_tx(&bar_full[st], (uint32_t)((1 + 2 * nkb) * AT_TILE));
                tma_load_2d(Qs, &tm_qkv, &bar_full[st], it.h * AT_D, (int)(it.s0 + it.q0));
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_2d(Ks + kb * AT_TILE, &tm_qkv, &bar_full[st], H + it.h * AT_D, (int)(it.s0 + kb * AT_BN));
                    tma_load_2d(Vs + kb * AT_TILE, &tm_qkv, &bar_full[st], 2 * H + it.h * AT_D, (int)(it.s0 + kb * AT_BN));
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =========================================================================== MMA issuer
        if (lane == 0) {
            AtWalker w(cu_seqlens, n_seq, heads, max_qt);
            AtItem cur, nxt;
            bool have = w.pop(cur);
            auto issue_s = [&](const AtItem &it, uint32_t n) {
                const int st = n & 1, sb = (NSB == 2) ? (int)(n & 1) : 0;
                const uint32_t use = (NSB == 2) ? (n >> 1) : n;
                mbar_wait(&bar_full[st], (n >> 1) & 1);
                mbar_wait(&bar_sfree[sb], (use & 1) ^ 1);                // softmax threads are done with this score buffer
                tc_fence_after();
                unsigned char *Qs = smem + st * stage_bytes, *Ks = Qs + AT_TILE;
                const uint64_t qd = make_sw128_kmajor_desc(Qs);
                const int nkb = (it.S + AT_BN - 1) / AT_BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int keys = min(AT_BN, it.S - kb * AT_BN);
                    const uint32_t idesc = at_idesc(AT_BM, (keys + 15) & ~15, 0);
                    const uint64_t kd = make_sw128_kmajor_desc(Ks + kb * AT_TILE);
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k)
                        umma_f16(tmem_base + (uint32_t)(sb * sbuf_cols + kb * AT_BN), desc_advance(qd, k * 32), desc_advance(kd, k * 32), idesc,
                                 k > 0 ? 1u : 0u);
                }
                umma_commit(&bar_s[sb]);
            };
            auto issue_pv = [&](const AtItem &it, uint32_t n) {
                const int st = n & 1, sb = (NSB == 2) ? (int)(n & 1) : 0;
                const uint32_t use = (NSB == 2) ? (n >> 1) : n;
                mbar_wait(&bar_p[sb], use & 1);                          // P of this item sits in tensor memory
                tc_fence_after();
                unsigned char *Vs = smem + st * stage_bytes + (1 + nkb_max) * AT_TILE;
                const int nkb = (it.S + AT_BN - 1) / AT_BN;
                const uint32_t sbase = tmem_base + (uint32_t)(sb * sbuf_cols);
                const uint32_t o_col = sbase + (uint32_t)(nkb * AT_BN - AT_D);      // dead upper score columns
                const uint32_t idesc = at_idesc(AT_BM, AT_D, 1);
                const int ksteps = (it.S + 15) >> 4;                     // 16 keys per MMA; P is 0 beyond the sequence
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t vd = at_desc_mnmajor(Vs + (k >> 3) * AT_TILE);
                    at_umma_ts(o_col, sbase + (uint32_t)(k * 8), desc_advance(vd, (k & 7) * 2048), idesc, k > 0 ? 1u : 0u);
                }
                umma_commit(&bar_o[sb]);
                umma_commit(&bar_empty[st]);                             // ... and the stage may be refilled
            };
            uint32_t n = 0;
            if (have) issue_s(cur, 0);
            while (have) {
                const bool more = w.pop(nxt);
                if (NSB == 2 && more) issue_s(nxt, n + 1);               // overlap: next item's scores before this item's P
                issue_pv(cur, n);
                if (NSB == 1 && more) issue_s(nxt, n + 1);
                cur = nxt;
                have = more;
                ++n;
            }
        }
        __syncwarp();
    } else {
        // =========================================================================== softmax + epilogue (warps 0-7)
        // group g = warp / 4 serves score buffer g (NSB == 2: items with n % 2 == g); with one buffer only group 0 works
        const int grp = warp >> 2;
        if (NSB == 2 || grp == 0) {
            const int row = tid & 127;                                       // query row = tensor-memory lane
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            AtWalker w(cu_seqlens, n_seq, heads, max_qt);
            AtItem it;
            for (uint32_t n = 0; w.pop(it); ++n) {
                if (NSB == 2 && (int)(n & 1u) != grp) continue;
                const int sb = (NSB == 2) ? grp : 0;
                const uint32_t use = (NSB == 2) ? (n >> 1) : n;
                const int nkb = (it.S + AT_BN - 1) / AT_BN;
                const int n_chunks = (it.S + 31) >> 5;                       // 32-key chunks that hold at least one key of the sequence
                const uint32_t sbase = tmem_base + (uint32_t)(sb * sbuf_cols) + lane_base;
                uint32_t *kb_words = kbits + (grp * 2 + (int)(use & 1u)) * 12;
                if (key_mask != nullptr) {                                   // validity words from the mask: warp w builds chunks w, w+4, w+8
                    for (int c = warp & 3; c < n_chunks; c += 4) {
                        const int key = c * 32 + lane;
                        const bool ok = key < it.S && __ldg(key_mask + it.s0 + key) != 0;
                        const uint32_t word = __ballot_sync(0xffffffffu, ok);
                        if (lane == 0) kb_words[c] = word;
                    }
                    at_named_barrier_128(1 + grp);
                }
                auto chunk_word = [&](int c) -> uint32_t {
                    if (key_mask != nullptr) return kb_words[c];
                    const int left = it.S - c * 32;
                    return left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                };
                mbar_wait(&bar_s[sb], use & 1);
                tc_fence_after();
                uint32_t va[32], vb[32];
                // ---- pass 1: row maximum over the keys that take part (the load of chunk c+1 flies under the work on chunk c)
                float m = -INFINITY;
                auto max_chunk = [&](const uint32_t (&v)[32], uint32_t word) {
                    if (word == 0xffffffffu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if ((word >> j) & 1u) m = fmaxf(m, __uint_as_float(v[j]));
                    }
                };
                tmem_ld_32x32(sbase, va);
                for (int c = 0; c < n_chunks; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < n_chunks) tmem_ld_32x32(sbase + (uint32_t)((c + 1) * 32), vb);
                    max_chunk(va, chunk_word(c));
                    if (c + 1 < n_chunks) {
                        tmem_ld_wait();
                        if (c + 2 < n_chunks) tmem_ld_32x32(sbase + (uint32_t)((c + 2) * 32), va);
                        max_chunk(vb, chunk_word(c + 1));
                    }
                }
                const bool any = m != -INFINITY;
                const float mscaled = any ? m * scale_log2e : 0.f;           // scale > 0: max commutes with the scaling
                // ---- pass 2: p = exp2(s * scale - m), row sum, P (fp16 pairs) over the score columns already consumed
                float l = 0.f;
                auto exp_chunk = [&](const uint32_t (&v)[32], uint32_t word, int c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float e0 = at_ex2(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -mscaled));
                        float e1 = at_ex2(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -mscaled));
                        if (word != 0xffffffffu) {
                            e0 = ((word >> (2 * j)) & 1u) ? e0 : 0.f;
                            e1 = ((word >> (2 * j + 1)) & 1u) ? e1 : 0.f;
                        }
                        l += e0 + e1;
                        pk[j] = at_pack(e0, e1);
                    }
                    at_tmem_st_32x16(sbase + (uint32_t)(c * 16), pk);
                };
                tmem_ld_32x32(sbase, va);
                for (int c = 0; c < n_chunks; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < n_chunks) tmem_ld_32x32(sbase + (uint32_t)((c + 1) * 32), vb);
                    exp_chunk(va, chunk_word(c), c);
                    if (c + 1 < n_chunks) {
                        tmem_ld_wait();
                        if (c + 2 < n_chunks) tmem_ld_32x32(sbase + (uint32_t)((c + 2) * 32), va);
                        exp_chunk(vb, chunk_word(c + 1), c + 1);
                    }
                }
                at_tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bar_p[sb]);

                // ---- epilogue: O / l -> fp16 -> one 128-byte row
                mbar_wait(&bar_o[sb], use & 1);
                tc_fence_after();
                const float inv = l > 0.f ? 1.f / l : 0.f;
                const bool store = it.q0 + row < it.S;
                __half *dst = out + (it.s0 + it.q0 + row) * (int64_t)H + it.h * AT_D;
                const uint32_t o_col = sbase + (uint32_t)(nkb * AT_BN - AT_D);
                tmem_ld_32x32(o_col, va);
                tmem_ld_32x32(o_col + 32u, vb);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bar_sfree[sb]);                                 // O is in registers: the score buffer may take the next item
                if (store) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o4;
                        o4.x = at_pack(__uint_as_float(va[8 * q]) * inv, __uint_as_float(va[8 * q + 1]) * inv);
                        o4.y = at_pack(__uint_as_float(va[8 * q + 2]) * inv, __uint_as_float(va[8 * q + 3]) * inv);
                        o4.z = at_pack(__uint_as_float(va[8 * q + 4]) * inv, __uint_as_float(va[8 * q + 5]) * inv);
                        o4.w = at_pack(__uint_as_float(va[8 * q + 6]) * inv, __uint_as_float(va[8 * q + 7]) * inv);
                        *reinterpret_cast<uint4 *>(dst + q * 8) = o4;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o4;
                        o4.x = at_pack(__uint_as_float(vb[8 * q]) * inv, __uint_as_float(vb[8 * q + 1]) * inv);
                        o4.y = at_pack(__uint_as_float(vb[8 * q + 2]) * inv, __uint_as_float(vb[8 * q + 3]) * inv);
                        o4.z = at_pack(__uint_as_float(vb[8 * q + 4]) * inv, __uint_as_float(vb[8 * q + 5]) * inv);
                        o4.w = at_pack(__uint_as_float(vb[8 * q + 6]) * inv, __uint_as_float(vb[8 * q + 7]) * inv);
                        *reinterpret_cast<uint4 *>(dst + 32 + q * 8) = o4;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static size_t at_smem_bytes(int nkb_max) { return (size_t)AT_STAGES * (1 + 2 * nkb_max) * AT_TILE + 384 + 1024; }

// total_tokens: rows of the packed qkv matrix (the tensor map clips / zero-fills beyond it).  Sequences longer than 384
// tokens are skipped here (the caller runs the mma.sync form for them).
int attention_varlen_tc(cudaStream_t st, const void *qkv, const int64_t *cu_seqlens, const int32_t *key_mask, void *out, int n_seq,
                        int max_seqlen, int64_t total_tokens, int heads)
{
    if (n_seq <= 0 || max_seqlen <= 0 || total_tokens <= 0) return 0;
    const int cap = max_seqlen < AT_MAX_KB * AT_BN ? max_seqlen : AT_MAX_KB * AT_BN;
    const int H = heads * AT_D;
    CUtensorMap tm;
    B2S_TRY(make_tmap_2d_kmajor(&tm, qkv, total_tokens, 3 * H, 3 * H, AT_BM, 0));
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int n_sm = 148;
    std::call_once(once, []() {
        attr_err = cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem_bytes(AT_MAX_KB));
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem_bytes(2));
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n_sm = v;
    });
    if (attr_err != cudaSuccess) return fail_cuda(attr_err, "cudaFuncSetAttribute(attention_tc_kernel)");
    const int nkb_max = (cap + AT_BN - 1) / AT_BN;
    const int max_qt = (cap + AT_BM - 1) / AT_BM;
    const int64_t items = (int64_t)n_seq * heads * max_qt;
    const int grid = (int)(items < n_sm ? items : n_sm);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)AT_D);
    if (nkb_max <= 2)
        attention_tc_kernel<2><<<grid, AT_THREADS, at_smem_bytes(nkb_max), st>>>(tm, cu_seqlens, key_mask, static_cast<__half *>(out), n_seq, heads,
                                                                                  max_qt, nkb_max, scale_log2e);
    else
        attention_tc_kernel<1><<<grid, AT_THREADS, at_smem_bytes(nkb_max), st>>>(tm, cu_seqlens, key_mask, static_cast<__half *>(out), n_seq, heads,
                                                                                  max_qt, nkb_max, scale_log2e);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2s
