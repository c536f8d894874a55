// attention_tc.cu -- variable-length self-attention for encoder models on the 5th-generation tensor cores
// (kernel K8, tcgen05 form; attention.cu keeps the mma.sync form for sequences longer than 384 tokens).
//     ctx[t, h*64:(h+1)*64] = softmax(Q_h K_h^T / sqrt(64) + key_mask) V_h      (non-causal, packed tokens)
// This is the attention block tritonserver's backends run as cuBLAS / cuDNN calls for the reference's transformer
// endpoint (examples/huggingface; model placed by engines/triton/triton_helper.py:159-186).
//
// PERSISTENT kernel, one CTA per SM, 10 warps, walking the work items (sequence, head, 128-query tile) round-robin:
//   warp 8 (one lane)  TMA producer: Q tile + every key / value block of the sequence (<= 3 blocks of 128 keys; box
//                      128 tokens x 64 halfs of the packed [T, 3H] qkv matrix, 128-byte swizzle) into TWO rings: Q + K
//                      (2 entries, released as soon as the score MMAs have read them, so the operands of item n+2 are
//                      on chip before a score buffer frees up) and V (3 entries, released after the P V MMAs).
//   warps 9-12 (one lane each)  MMA issuers, a PAIR per score buffer (issuing a tcgen05.mma costs ~125 cycles of
//                      register -> uniform-register traffic, more than these small MMAs run): S blocks alternate between
//                      the pair; P V is split by key range into TWO output accumulators (added in the epilogue).
//                      S = Q K^T     tcgen05.mma 128 x N x 16 (N = keys of the block rounded up to 16),
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
#include <stdlib.h>

namespace b2s {

using namespace sm100;

int make_tmap_2d_kmajor(CUtensorMap *out, const void *base, int64_t rows, int64_t K, int64_t ld_elems, int box_rows,
                        int is_bf16);

constexpr int AT_BM = 128;                  // queries per work item
constexpr int AT_BN = 128;                  // keys per block
constexpr int AT_D = 64;                    // head dim
constexpr int AT_MAX_KB = 3;                // <= 384 keys: S fits tensor memory
constexpr int AT_TILE = AT_BM * AT_D * 2;   // 16 KB: one [128 x 64] fp16 tile
constexpr int AT_QK_STAGES = 2;
constexpr int AT_THREADS = 416;              // 2 softmax warpgroups (warps 0-3, 4-7) + TMA producer (8) + 2 x 2 MMA issuers (9-12)

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
    uint32_t next, total, stride;                         // n_seq * heads * max_qt < 2^31 (checked by the launcher)
    __device__ AtWalker(const int64_t *cu_, int n_seq_, int heads_, int max_qt_)
        : cu(cu_), n_seq(n_seq_), heads(heads_), max_qt(max_qt_), next(blockIdx.x), total((uint32_t)(n_seq_ * heads_ * max_qt_)), stride(gridDim.x) {}
    __device__ bool pop(AtItem &it)
    {
        while (next < total) {
            const uint32_t i = next;
            next += stride;
            const int qt = (int)(i % (uint32_t)max_qt);
            const uint32_t r = i / (uint32_t)max_qt;
            const int h = (int)(r % (uint32_t)heads), b = (int)(r / (uint32_t)heads);
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
                    float scale_log2e, long long *__restrict__ dbg)
{
    // developer aid (B2S_ATTN_TIMING=1): SM-clock stamps of CTA 0's first 16 items, [item][16]
#define AT_STAMP(n, k)                                                                  \
    do {                                                                                \
        if (dbg && blockIdx.x == 0 && (n) < 16) dbg[(n) * 16 + (k)] = clock64();        \
    } while (0)
    extern __shared__ unsigned char at_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = heads * AT_D;
    const int qk_bytes = (1 + nkb_max) * AT_TILE;                // one Q + K entry
    const int v_bytes = nkb_max * AT_TILE;                       // one V entry
    const int NV = nkb_max <= 2 ? 3 : 2;                         // V ring depth (shared memory budget)
    unsigned char *v_ring = smem + AT_QK_STAGES * qk_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(v_ring + NV * v_bytes);
    uint64_t *bar_full = bars, *bar_empty = bars + 2, *bar_s = bars + 4, *bar_p = bars + 6, *bar_o = bars + 8, *bar_sfree = bars + 10;
    uint64_t *bar_vfull = bars + 12, *bar_vempty = bars + 15;    // [3] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 18);
    uint32_t *kbits = tmem_slot + 2;                             // [2 groups][2][12]: validity word of every 32-key chunk (masked path)
    // tensor-memory columns of one score buffer: scores [0, nkb*128) (P overwrites the front), the two output accumulators
    // behind them: NSB == 2: 256 columns, O_b at [128,192) and O_a at [192,256) (block 1's scores are dead by then);
    // NSB == 1: scores in [0,384), O_a at [384,448), O_b at [448,512)
    const int sbuf_cols = NSB == 2 ? 256 : 512;
    const uint32_t oa_off = NSB == 2 ? 192u : 384u, ob_off = NSB == 2 ? 128u : 448u;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 2);       // both issuers of the pair commit
            mbar_init(&bar_s[i], 2);
            mbar_init(&bar_p[i], 128);
            mbar_init(&bar_o[i], 2);
            mbar_init(&bar_sfree[i], 128);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&bar_vfull[i], 1);
            mbar_init(&bar_vempty[i], 2);
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
                const int st = n & 1, sv = (int)(n % (uint32_t)NV);
                const int nkb = (it.S + AT_BN - 1) / AT_BN;
                AT_STAMP(n, 0);
                mbar_wait(&bar_empty[st], ((n >> 1) & 1) ^ 1);          // the score MMAs that read this Q / K entry have retired
                AT_STAMP(n, 1);
                unsigned char *Qs = smem + st * qk_bytes, *Ks = Qs + AT_TILE;
                mbar_arrive_expect_tx(&bar_full[st], (uint32_t)((1 + nkb) * AT_TILE));
                tma_load_2d(Qs, &tm_qkv, &bar_full[st], it.h * AT_D, (int)(it.s0 + it.q0));
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(Ks + kb * AT_TILE, &tm_qkv, &bar_full[st], H + it.h * AT_D, (int)(it.s0 + kb * AT_BN));
                mbar_wait(&bar_vempty[sv], ((n / (uint32_t)NV) & 1) ^ 1);  // the P V MMAs that read this V entry have retired
                unsigned char *Vs = v_ring + sv * v_bytes;
                mbar_arrive_expect_tx(&bar_vfull[sv], (uint32_t)(nkb * AT_TILE));
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(Vs + kb * AT_TILE, &tm_qkv, &bar_vfull[sv], 2 * H + it.h * AT_D, (int)(it.s0 + kb * AT_BN));
            }
        }
        __syncwarp();
    } else if (warp >= 9) {
        // =========================================================================== MMA issuers: pair (9,10) -> buffer 0, (11,12) -> buffer 1
        const int pair = (warp - 9) >> 1, role = (warp - 9) & 1;         // role 0: even key blocks / first half of the key range
        if (lane == 0 && (NSB == 2 || pair == 0)) {
            AtWalker w(cu_seqlens, n_seq, heads, max_qt);
            uint32_t n_pop = 0;
            auto pop_own = [&](AtItem &o, uint32_t &n_o) -> bool {       // next item of THIS pair's buffer (n counts every item)
                while (w.pop(o)) {
                    const uint32_t n_cur = n_pop++;
                    if (NSB == 2 && (int)(n_cur & 1u) != pair) continue;
                    n_o = n_cur;
                    return true;
                }
                return false;
            };
            AtItem it, nxt;
            uint32_t n = 0, n_nxt = 0;
            bool have = pop_own(it, n);
            if (have) mbar_wait(&bar_full[n & 1], (n >> 1) & 1);
            while (have) {
                const int st = n & 1, sv = (int)(n % (uint32_t)NV), sb = (NSB == 2) ? pair : 0;
                const uint32_t use = (NSB == 2) ? (n >> 1) : n;
                const int nkb = (it.S + AT_BN - 1) / AT_BN;
                const uint32_t sbase = tmem_base + (uint32_t)(sb * sbuf_cols);
                // ---- S = Q K^T: this issuer takes the key blocks kb == role (mod 2); Q / K are on chip already
                if (role == 0) AT_STAMP(n, 3);
                mbar_wait(&bar_sfree[sb], (use & 1) ^ 1);                // softmax threads are done with this score buffer
                if (role == 0) AT_STAMP(n, 4);
                tc_fence_after();
                unsigned char *Qs = smem + st * qk_bytes, *Ks = Qs + AT_TILE;
                const uint64_t qd = make_sw128_kmajor_desc(Qs);
                for (int kb = role; kb < nkb; kb += 2) {
                    const int keys = min(AT_BN, it.S - kb * AT_BN);
                    const uint32_t idesc = at_idesc(AT_BM, (keys + 15) & ~15, 0);
                    const uint64_t kd = make_sw128_kmajor_desc(Ks + kb * AT_TILE);
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k)
                        umma_f16(sbase + (uint32_t)(kb * AT_BN), desc_advance(qd, k * 32), desc_advance(kd, k * 32), idesc, k > 0 ? 1u : 0u);
                }
                umma_commit(&bar_s[sb]);                                 // (an issuer without blocks still arrives)
                umma_commit(&bar_empty[st]);                             // Q / K entry free once these MMAs retire
                if (role == 0) AT_STAMP(n, 5);
                // ---- while the softmax group works: find the next item of this buffer and wait for its Q / K and this item's V
                const bool more = pop_own(nxt, n_nxt);
                if (more) mbar_wait(&bar_full[n_nxt & 1], (n_nxt >> 1) & 1);
                mbar_wait(&bar_vfull[sv], (n / (uint32_t)NV) & 1);
                // ---- O_role = P[:, range] V[range]: role 0 the first half of the 16-key steps, role 1 the rest
                if (role == 0) AT_STAMP(n, 6);
                mbar_wait(&bar_p[sb], use & 1);                          // P of this item sits in tensor memory
                if (role == 0) AT_STAMP(n, 7);
                tc_fence_after();
                unsigned char *Vs = v_ring + sv * v_bytes;
                const uint32_t idesc_o = at_idesc(AT_BM, AT_D, 1);
                const int ksteps = (it.S + 15) >> 4;                     // 16 keys per MMA; P is 0 beyond the sequence
                const int half = (ksteps + 1) >> 1;
                const int k_lo = role == 0 ? 0 : half, k_hi = role == 0 ? half : ksteps;
                const uint32_t o_col = sbase + (role == 0 ? oa_off : ob_off);
                for (int k = k_lo; k < k_hi; ++k) {
                    const uint64_t vd = at_desc_mnmajor(Vs + (k >> 3) * AT_TILE);
                    at_umma_ts(o_col, sbase + (uint32_t)(k * 8), desc_advance(vd, (k & 7) * 2048), idesc_o, k > k_lo ? 1u : 0u);
                }
                umma_commit(&bar_o[sb]);
                umma_commit(&bar_vempty[sv]);                            // ... and the V entry may be refilled
                if (role == 0) AT_STAMP(n, 8);
                it = nxt;
                n = n_nxt;
                have = more;
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
            uint32_t n_pop = 0;
            auto pop_own = [&](AtItem &o, uint32_t &n_o) -> bool {           // next item of THIS group's buffer
                while (w.pop(o)) {
                    const uint32_t n_cur = n_pop++;
                    if (NSB == 2 && (int)(n_cur & 1u) != grp) continue;
                    n_o = n_cur;
                    return true;
                }
                return false;
            };
            AtItem it, nxt;
            uint32_t n = 0, n_nxt = 0;
            bool have = pop_own(it, n), more = false;
            for (; have; it = nxt, n = n_nxt, have = more) {
                const int sb = (NSB == 2) ? grp : 0;
                const uint32_t use = (NSB == 2) ? (n >> 1) : n;
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
                if (row == 0) AT_STAMP(n, 9);
                mbar_wait(&bar_s[sb], use & 1);
                if (row == 0) AT_STAMP(n, 10);
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
                if (row == 0) AT_STAMP(n, 11);
                more = pop_own(nxt, n_nxt);                                  // off the critical path: the P V MMAs are running

                // ---- epilogue: O / l -> fp16 -> one 128-byte row
                mbar_wait(&bar_o[sb], use & 1);
                if (row == 0) AT_STAMP(n, 12);
                tc_fence_after();
                const float inv = l > 0.f ? 1.f / l : 0.f;
                const bool store = it.q0 + row < it.S;
                __half *dst = out + (it.s0 + it.q0 + row) * (int64_t)H + it.h * AT_D;
                const bool two = ((it.S + 15) >> 4) >= 2;                    // the second issuer had key steps: O = O_a + O_b
                uint32_t vc[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld_32x32(sbase + oa_off + (uint32_t)(c * 32), c == 0 ? va : vb);
                    if (two) {
                        tmem_ld_32x32(sbase + ob_off + (uint32_t)(c * 32), vc);
                        tmem_ld_wait();
                        uint32_t (&dst_v)[32] = c == 0 ? va : vb;
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst_v[j] = __float_as_uint(__uint_as_float(dst_v[j]) + __uint_as_float(vc[j]));
                    }
                }
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bar_sfree[sb]);                                 // O is in registers: the score buffer may take the next item
                if (row == 0) AT_STAMP(n, 13);
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
#undef AT_STAMP
}

static long long *g_attn_dbg = nullptr;

static size_t at_smem_bytes(int nkb_max)
{
    return (size_t)AT_QK_STAGES * (1 + nkb_max) * AT_TILE + (size_t)(nkb_max <= 2 ? 3 : 2) * nkb_max * AT_TILE + 448 + 1024;
}

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
    if (items >= (int64_t)1 << 31) return fail(B2S_ERR_INVALID, "attention: too many work items");
    const int grid = (int)(items < n_sm ? items : n_sm);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)AT_D);
    static long long *dbg = []() -> long long * {
        const char *e = getenv("B2S_ATTN_TIMING");
        if (!(e && e[0] == '1')) return nullptr;
        long long *p = nullptr;
        if (cudaMalloc(reinterpret_cast<void **>(&p), 256 * sizeof(long long)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, 256 * sizeof(long long));
        return p;
    }();
    g_attn_dbg = dbg;
    if (nkb_max <= 2)
        attention_tc_kernel<2><<<grid, AT_THREADS, at_smem_bytes(nkb_max), st>>>(tm, cu_seqlens, key_mask, static_cast<__half *>(out), n_seq, heads,
                                                                                  max_qt, nkb_max, scale_log2e, dbg);
    else
        attention_tc_kernel<1><<<grid, AT_THREADS, at_smem_bytes(nkb_max), st>>>(tm, cu_seqlens, key_mask, static_cast<__half *>(out), n_seq, heads,
                                                                                  max_qt, nkb_max, scale_log2e, dbg);
    count_launch();
    B2S_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2s

// developer aid: the stamps of the last launch (256 int64), B2S_ATTN_TIMING=1
extern "C" B2S_API int b2s_debug_attention_stamps(long long *out256)
{
    if (!b2s::g_attn_dbg || !out256) return b2s::fail(B2S_ERR_INVALID, "set B2S_ATTN_TIMING=1 and run an attention launch first");
    cudaDeviceSynchronize();
    return cudaMemcpy(out256, b2s::g_attn_dbg, 256 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : b2s::fail(B2S_ERR_CUDA, "copy failed");
}
