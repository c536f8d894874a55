// skinny.cu -- weight-streaming GEMM for the decode step of the LLM endpoint (BASELINE.json configs[4];
// the reference hands this endpoint to vLLM: clearml_serving/serving/preprocess_service.py:1097-1348):
//     Y[m, n] += sum_k X[m, k] * W[n, k]         W: [N_out, K] bf16 K-major (nn.Linear weight as stored)
//                                                X: [M <= 32, K] bf16 (one row per running sequence)
//                                                Y: [32, N_out] fp32, accumulated in place
// A decode step multiplies every weight of the model by at most max_batch (32) activation rows: the op is
// bound by streaming W from HBM once (2 bytes per weight, 32 FMAs each), so the kernel is organised around
// keeping every SM's share of that stream in flight, not around tensor-pipe utilisation:
//   * operands are swapped: the WEIGHT tile (128 rows x 64 k) is the UMMA "A" operand and X the 32-wide "B"
//     operand, so one tcgen05.mma 128x32x16 consumes a full 128-row weight slab with no padding of the batch
//     to 128 rows;
//   * stream-K: the (weight tile, k-block) units of the whole GEMM are dealt evenly to one persistent CTA per
//     SM, each CTA owning a CONTIGUOUS range, so all 148 SMs pull the same number of bytes whatever N_out is
//     (a 4096 x 2048 projection has only 32 row tiles); a CTA's range covers at most a few (tile, k-range)
//     segments, each accumulated in TMEM (double-buffered) and added into Y with red.global.add.f32 (a warp's
//     32 lanes hold 32 consecutive n of one batch row: every reduction instruction is one coalesced 128-byte line);
//   * TMA rings of 20 KB stages (W 16 KB + X 4 KB), two CTAs x 5 stages per SM, keep ~200 KB per SM in flight;
//   * programmatic dependent launch: the kernel may start while its predecessor (a small normalisation /
//     attention kernel) still runs -- the producer fills the whole ring with WEIGHT tiles first and only then
//     waits for the predecessor (griddepcontrol.wait) before it requests the activation tiles, so barrier
//     set-up, TMEM allocation and the first 23 MB of the weight stream are off the critical path.
// The consumer kernels (llm.cu) read Y and zero it again, so no separate memset is needed.
// Roofline: HBM; algorithmic bytes = 2 * N_out * K (the weight), X and Y are L2-resident.
#include "common.cuh"
#include "sm100.cuh"

#include <cuda_bf16.h>

#include <stdlib.h>

#include <mutex>

namespace b2s {

using namespace sm100;

int make_tmap_2d_kmajor(CUtensorMap *out, const void *base, int64_t rows, int64_t K, int64_t ld_elems,
                        int box_rows, int is_bf16);

constexpr int SK_BM = 128;       // weight rows per tile (UMMA M)
constexpr int SK_BN = 32;        // activation rows (UMMA N)
constexpr int SK_BK = 64;
constexpr int SK_THREADS = 256;  // warp 0 TMA, 1 MMA, 2 TMEM owner, 3 idle, 4-7 epilogue

template <int SK_STAGES>
struct SkSmem {
    static constexpr int A_BYTES = SK_BM * SK_BK * 2;   // 16 KB
    static constexpr int B_BYTES = SK_BN * SK_BK * 2;   // 4 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = SK_STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

__device__ __forceinline__ void red_add(float *addr, float a)
{
    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(a) : "memory");
}
// system scope: the accumulator is also the target of the tensor-parallel peer's reductions (over NVLink)
__device__ __forceinline__ void red_add_sys(float *addr, float a)
{
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(a) : "memory");
}

// SK_STAGES = 10: one CTA per SM; SK_STAGES = 5: two CTAs per SM (same bytes in flight per SM) -- the second
// form lets the CTAs of the NEXT projection move in, and start streaming their weights, as soon as half an SM
// frees up at the tail of the current one.
template <int SK_STAGES>
__global__ void __launch_bounds__(SK_THREADS, SK_STAGES > 5 ? 1 : 2)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   float *__restrict__ y, int n_out, int m_rows, int num_k, int total_units, int l2_ahead,
                   const uint32_t *idle_flag, const uint32_t *gen, int idle_want, float *__restrict__ y_peer, uint32_t *peer_done)
{
    using S = SkSmem<SK_STAGES>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + S::BAR_OFFSET);
    uint64_t *empty_bar = full_bar + SK_STAGES;
    uint64_t *tmem_full_bar = empty_bar + SK_STAGES;   // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    griddep_launch_dependents();   // the (small) kernel after this one may get resident early; it waits for us
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // contiguous unit range of this CTA (unit = tile * num_k + kb)
    const int u0 = (int)(((int64_t)total_units * blockIdx.x) / gridDim.x);
    const int u1 = (int)(((int64_t)total_units * (blockIdx.x + 1)) / gridDim.x);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_w);
        prefetch_tensormap(&tmap_x);
        for (int s = 0; s < SK_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 2 * SK_BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // The weight does not depend on the kernel that precedes this one in the stream (only X does): with a
            // programmatic dependent launch the ring is filled with weight tiles while that kernel still runs.
            const int pre = min(u1 - u0, SK_STAGES);
            for (int i = 0; i < pre; ++i) {
                const int u = u0 + i, tile = u / num_k, kb = u - tile * num_k;
                mbar_arrive_expect_tx(&full_bar[i], S::STAGE_BYTES);
                tma_load_2d(smem + i * S::STAGE_BYTES, &tmap_w, &full_bar[i], kb * SK_BK, tile * SK_BM);
            }
            // ... and once HBM goes IDLE the units behind the ring are pulled into L2.  "Idle" is signalled by the small kernel
            // between the previous projection and this one (RMSNorm / SwiGLU): it raises `idle_flag` as soon as ITS dependency
            // has resolved, i.e. when the previous projection has stopped streaming; for the few microseconds that kernel then
            // needs, the ring (200 KB per SM) is full and nothing else would use the memory system.  Prefetching at kernel
            // start instead competes with the predecessor's tail and is slower at every depth
            // (profiles/r02_llm_attn_decode_stream.txt).  Bounded poll: a missing signal only costs the prefetch.
            if (idle_flag && l2_ahead > 0 && u0 + pre < u1) {
                const uint32_t want = *reinterpret_cast<const volatile uint32_t *>(gen) * 1024u + (uint32_t)idle_want;
                const long long t0 = clock64();
                bool idle = false;
                while (clock64() - t0 < 60000) {
                    if ((int32_t)(*reinterpret_cast<const volatile uint32_t *>(idle_flag) - want) >= 0) { idle = true; break; }
                    __nanosleep(100);
                }
                if (idle) {
                    const int ahead = min(u1 - u0, pre + l2_ahead);
                    for (int i = pre; i < ahead; ++i) {
                        const int u = u0 + i, tile = u / num_k, kb = u - tile * num_k;
                        tma_prefetch_l2_2d(&tmap_w, kb * SK_BK, tile * SK_BM);
                    }
                }
            }
            griddep_wait();
            for (int i = 0; i < pre; ++i) {
                const int u = u0 + i, tile = u / num_k, kb = u - tile * num_k;
                tma_load_2d(smem + i * S::STAGE_BYTES + S::A_BYTES, &tmap_x, &full_bar[i], kb * SK_BK, 0);
            }
            int stage = pre == SK_STAGES ? 0 : pre;
            uint32_t phase = pre == SK_STAGES ? 1 : 0;
            for (int u = u0 + pre; u < u1; ++u) {
                const int tile = u / num_k, kb = u - tile * num_k;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char *sa = smem + stage * S::STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                tma_load_2d(sa, &tmap_w, &full_bar[stage], kb * SK_BK, tile * SK_BM);
                tma_load_2d(sa + S::A_BYTES, &tmap_x, &full_bar[stage], kb * SK_BK, 0);
                if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(SK_BM, SK_BN, 1);   // bf16 operands
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            int u = u0;
            while (u < u1) {
                const int tile = u / num_k;
                const int seg_end = min(u1, (tile + 1) * num_k);
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * SK_BN);
                for (int first = u; u < seg_end; ++u) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    unsigned char *sa = smem + stage * S::STAGE_BYTES;
                    const uint64_t adesc = make_sw128_kmajor_desc(sa);
                    const uint64_t bdesc = make_sw128_kmajor_desc(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < SK_BK / 16; ++k)
                        umma_f16(d_tmem, desc_advance(adesc, k * 32), desc_advance(bdesc, k * 32), idesc,
                                 (uint32_t)((u != first) || (k != 0)));
                    umma_commit(&empty_bar[stage]);
                    if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full_bar[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int as = 0;
        uint32_t aphase = 0;
        int u = u0;
        while (u < u1) {
            const int tile = u / num_k;
            u = min(u1, (tile + 1) * num_k);
            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * SK_BN), v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
            const int n = tile * SK_BM + q * 32 + lane;
            if (n < n_out && !y_peer) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < m_rows) red_add(y + (size_t)j * n_out + n, __uint_as_float(v[j]));
            } else if (n < n_out) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < m_rows) red_add_sys(y + (size_t)j * n_out + n, __uint_as_float(v[j]));
                // tensor-parallel pair: the same partial is ALSO added into the peer's accumulator over NVLink (a row-parallel
                // projection's all-reduce as two pushes: afterwards both ranks hold the full sum locally, llm.cu)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < m_rows) red_add_sys(y_peer + (size_t)j * n_out + n, __uint_as_float(v[j]));
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (peer_done) {
            // this CTA's pushes are complete: count it on the PEER (release at system scope after a barrier of the four
            // epilogue warps, so every warp's reductions are ordered before the count).  The peer's consumer kernel waits
            // for (step + 1) x grid counts instead of a flag that a LATER kernel of this rank would have to send.
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            if (warp == 4 && lane == 0) asm volatile("red.release.sys.global.add.u32 [%0], 1;\n" ::"l"(peer_done) : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * SK_BN);
    }
}

static int sk_num_sms()
{
    static int n = []() {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

// tensor maps for one (weight, activation buffer) pair; both are cached by the caller (CUDA-graph friendly)
int skinny_make_maps(CUtensorMap *tw, CUtensorMap *tx, const void *W, int64_t n_out, int64_t K, const void *X, int64_t x_rows)
{
    if (K % 8 != 0) return fail(B2S_ERR_INVALID, "skinny gemm: K must be a multiple of 8");
    if (x_rows > SK_BN) return fail(B2S_ERR_INVALID, "skinny gemm: at most %d activation rows", SK_BN);
    B2S_TRY(make_tmap_2d_kmajor(tw, W, n_out, K, K, SK_BM, 1));
    B2S_TRY(make_tmap_2d_kmajor(tx, X, x_rows, K, K, SK_BN, 1));
    return 0;
}

template <int STAGES>
static int skinny_launch(cudaStream_t st, const CUtensorMap &tw, const CUtensorMap &tx, float *y, int n_out, int m_rows, int num_k,
                         int64_t units, int ctas_per_sm, const uint32_t *idle_flag, const uint32_t *gen, int idle_want, float *y_peer,
                         uint32_t *peer_done, int *grid_out)
{
    using S = SkSmem<STAGES>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, []() {
        attr_err = cudaFuncSetAttribute(skinny_gemm_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    });
    if (attr_err != cudaSuccess) return fail_cuda(attr_err, "cudaFuncSetAttribute(skinny gemm)");
    // at least 4 k-blocks per CTA so tiny problems do not pay hundreds of prologues for nothing
    int grid = sk_num_sms() * ctas_per_sm;
    if (units < (int64_t)grid * 4) grid = (int)((units + 3) / 4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = []() { const char *e = getenv("B2S_SKINNY_PDL"); return !(e && e[0] == '0'); }();
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    // units per CTA prefetched into L2 behind the ring once the predecessor signals an idle HBM (B2S_SKINNY_L2_AHEAD, 0 = off)
    static const int l2_ahead = []() { const char *e = getenv("B2S_SKINNY_L2_AHEAD"); return e ? atoi(e) : 4; }();
    B2S_CUDA(cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<STAGES>, tw, tx, y, n_out, m_rows < SK_BN ? m_rows : SK_BN, num_k, (int)units,
                                l2_ahead, idle_flag, gen, idle_want, y_peer, peer_done));
    if (grid_out) *grid_out = grid;
    count_launch();
    return 0;
}

int skinny_gemm_maps(cudaStream_t st, const CUtensorMap &tw, const CUtensorMap &tx, float *y, int n_out, int K, int m_rows,
                     const uint32_t *idle_flag, const uint32_t *gen, int idle_want, float *y_peer, uint32_t *peer_done, int *grid_out)
{
    if (n_out <= 0 || K <= 0) return 0;
    const int num_k = (K + SK_BK - 1) / SK_BK;
    const int tiles = (n_out + SK_BM - 1) / SK_BM;
    const int64_t units = (int64_t)tiles * num_k;
    if (units > INT32_MAX) return fail(B2S_ERR_INVALID, "skinny gemm: problem too large");
    static const int per_sm = []() { const char *e = getenv("B2S_SKINNY_CTAS"); return (e && e[0] == '1') ? 1 : 2; }();
    return per_sm == 1 ? skinny_launch<10>(st, tw, tx, y, n_out, m_rows, num_k, units, 1, idle_flag, gen, idle_want, y_peer, peer_done, grid_out)
                       : skinny_launch<5>(st, tw, tx, y, n_out, m_rows, num_k, units, 2, idle_flag, gen, idle_want, y_peer, peer_done, grid_out);
}

}  // namespace b2s

// C ABI (operator level, device pointers): y[m][n_out] fp32 += X[m,K] . W[n_out,K]^T ; y must be zeroed by the
// caller before the first accumulation.
extern "C" B2S_API int b2s_op_skinny_gemm(int device, void *cuda_stream, const void *W, const void *X, float *y,
                                           int n_out, int K, int m)
{
    using namespace b2s;
    B2S_CUDA(cudaSetDevice(device));
    CUtensorMap tw, tx;
    B2S_TRY(skinny_make_maps(&tw, &tx, W, n_out, K, X, m));
    return skinny_gemm_maps(static_cast<cudaStream_t>(cuda_stream), tw, tx, y, n_out, K, m, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
}
