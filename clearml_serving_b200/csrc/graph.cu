// graph.cu -- B2S_MODEL_GRAPH: a small op-list executor for encoder-style DL models (BERT class).
//
// Stands in for what tritonserver does with the model file the reference ships to it
// (clearml_serving/engines/triton/triton_helper.py:159-186 places model.pt / model.onnx, :378-385
// selects the libtorch / ONNX-Runtime backend): the graph is lowered OFFLINE by
// clearml_serving_b200/formats.py into a flat list of the hand-written sm_100a kernels of this
// library -- no ONNX Runtime, no libtorch, no backend dispatch at run time.
//
// Execution model: tokens of all requests of a batch are PACKED (ragged) -- `row_offsets`
// (cu_seqlens) delimits sequences -- so no FLOP is spent on padding and a request's result cannot
// depend on its batch-mates.  Activations live in per-stream scratch buffers sized for the
// stream's max tokens; weights stay resident in HBM as fp16 (GEMM operands) / fp32 (bias, LayerNorm).
// TMA tensor maps are built once per (stream, op): the A operand of every GEMM is a fixed scratch
// buffer whose row count is the stream's capacity (rows past the live batch are never stored).
//
// Blob layout ("B2SG"), little endian:
//   header  {magic, version, n_tensors, n_buffers, n_ops, n_inputs, n_outputs, max_pos, out_buffer[4], in_dtype[4],
//            in_row_elems[4]}
//   tensors n_tensors x {u32 dtype, u32 ndim, i64 shape[4], u64 offset, u64 nbytes}
//   buffers n_buffers x {u32 dtype, u32 rows_kind (0 = per token, k = k rows per batch item), i64 cols}
//   ops     n_ops x {u32 opcode, i32 a[15], f32 f[4]}
//   data    weights, each 256-byte aligned
#include "common.cuh"

#include <cuda.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

namespace b2s {

// kernels defined in gemm.cu / norm.cu / attention.cu
int make_tmap_2d_kmajor(CUtensorMap *out, const void *base, int64_t rows, int64_t K, int64_t ld_elems, int box_rows,
                        int is_bf16);
int gemm_tn_maps(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, int bn, int M, int N, int K,
                 const GemmEpilogue &ep);
int gemm_bn_for(int N);
bool gemm_prefer_bn256(int M, int N, int K);
bool gemm_pair_enabled();
bool gemm_prefer_bn192(int M, int N, int K, int out_f32, int act);
int gemm_tn_maps_2sm192(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb_half96, int M, int N, int K,
                        const GemmEpilogue &ep);
int gemm_tn_maps_pair(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb_half, int M, int N, int K,
                      const GemmEpilogue &ep);
int layernorm(cudaStream_t st, const float *in, int64_t rows, int H, const float *gamma, const float *beta, float eps,
              void *out16, float *out32);
int embed_layernorm(cudaStream_t st, const int32_t *ids, const int32_t *types, const int64_t *cu_seqlens, int n_seq,
                    int64_t n_tokens, int H, const void *word, const void *pos, const void *type, int vocab, int max_pos,
                    int n_types, const float *gamma, const float *beta, float eps, void *out16, float *out32);
int gather_first(cudaStream_t st, const void *in, const int64_t *cu_seqlens, int n_seq, int H, void *out);
int attention_varlen(cudaStream_t st, const void *qkv, const int64_t *cu_seqlens, const int32_t *key_mask, void *out,
                     int n_seq, int max_seqlen, int heads, int head_dim, int64_t total_tokens);
int make_tmap_im2col_nhwc(CUtensorMap *out, const void *base, int64_t n_img, int H, int W, int C, int KS, int stride, int pad);
ConvGeom make_conv_geom(int H, int W, int C, int KS, int stride, int pad);
int conv_implicit_maps(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, int bn, bool pair, int M, int N, int K,
                       const GemmEpilogue &ep, const ConvGeom &cg);
int make_tmap_stem_s2d(CUtensorMap *out, const void *base, int64_t n_img, int Hz, int Wz, int OH, int OW, int rows_per_tile);
ConvGeom make_stem_geom(int OH, int OW);
int conv_stem_maps(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, int bn, int M, int N, const GemmEpilogue &ep,
                   const ConvGeom &cg);
// conv.cu
int nchw_to_s2d(cudaStream_t st, const void *in, int in_dtype, int64_t n_img, int C, int H, int W, int Hz, int Wz, void *out);
int nchw_to_nhwc(cudaStream_t st, const void *in, int in_dtype, int64_t n_img, int C, int H, int W, int Cp, void *out);
int im2col_nhwc(cudaStream_t st, const void *in, int64_t n_img, int H, int W, int C, int KH, int KW, int stride, int pad,
                int OH, int OW, int Kp, void *out);
int maxpool3x3s2(cudaStream_t st, const void *in, int64_t n_img, int H, int W, int C, int OH, int OW, void *out);
int avgpool(cudaStream_t st, const void *in, int64_t n_img, int HW, int C, void *out);

namespace {

enum GraphOp {
    OP_EMBED_LN = 1, OP_LINEAR = 2, OP_LAYERNORM = 3, OP_ATTENTION = 4, OP_GATHER_FIRST = 5,
    OP_NCHW_TO_NHWC = 6, OP_IM2COL = 7, OP_MAXPOOL = 8, OP_AVGPOOL = 9, OP_CONV = 10,
    OP_STEM_S2D = 11, OP_CONV_STEM = 12
};

struct GHeader {
    char magic[4];
    uint32_t version, n_tensors, n_buffers, n_ops, n_inputs, n_outputs, max_pos;
    int32_t out_buffer[4];
    int32_t in_dtype[4];
    int64_t in_row_elems[4];   // elements per batch row of input i, -1 = variable length (packed tokens)
};
static_assert(sizeof(GHeader) == 96, "graph header layout");
struct GTensor {
    uint32_t dtype, ndim;
    int64_t shape[4];
    uint64_t offset, nbytes;
};
static_assert(sizeof(GTensor) == 56, "graph tensor layout");
struct GBuffer {
    uint32_t dtype, rows_kind;   // rows_kind 0: one row per packed token; k >= 1: k rows per batch item
    int64_t cols;
};
static_assert(sizeof(GBuffer) == 16, "graph buffer layout");
struct GOp {
    uint32_t opcode;
    int32_t a[15];
    float f[4];
};
static_assert(sizeof(GOp) == 80, "graph op layout");

struct Plan {  // per-stream state: buffer addresses + cached A-operand tensor maps
    std::vector<unsigned char *> buf;
    std::vector<CUtensorMap> amap;  // one per op (only LINEAR ops use it)
};

struct GraphModel : Model {
    GHeader h{};
    std::vector<GTensor> tensors;
    std::vector<GBuffer> buffers;
    std::vector<GOp> ops;
    std::vector<CUtensorMap> bmap;     // weight tensor maps (narrow tiles: box 64 / 128), one per op
    std::vector<CUtensorMap> bmap256;  // weight tensor maps for 128 x 256 tiles (N >= 256)
    std::vector<CUtensorMap> bmap96;   // box 96 = half of a 192-wide 2-SM tile (fp32 outputs whose width is a multiple of 192)
    unsigned char *d_data = nullptr;
    std::mutex mu;
    std::map<void *, Plan> plans;   // keyed by the stream's scratch base
    bool ragged = false;

    ~GraphModel() override
    {
        if (d_data) { cudaSetDevice(device); cudaFree(d_data); }
    }
    const void *tptr(int idx) const { return idx < 0 ? nullptr : d_data + tensors[idx].offset; }
    void on_stream_destroy(void *scratch) override
    {
        std::lock_guard<std::mutex> l(mu);
        plans.erase(scratch);
    }

    size_t buffer_bytes(const GBuffer &b, int64_t max_rows, int64_t max_tokens) const
    {
        const int64_t rows = b.rows_kind == 0 ? max_tokens : max_rows * (int64_t)b.rows_kind;
        return (size_t)round_up(rows * b.cols * (int64_t)dtype_size(b.dtype), 1024) + 1024;
    }
    size_t scratch_bytes(int64_t max_rows, int64_t max_row_elems) const override
    {
        const int64_t max_tokens = max_rows * (max_row_elems > 0 ? max_row_elems : 1);
        size_t total = 1024;
        for (const GBuffer &b : buffers) total += buffer_bytes(b, max_rows, max_tokens);
        return total;
    }

    int build_plan(Plan &pl, void *scratch, size_t scratch_sz, int64_t max_rows, int64_t max_tokens)
    {
        unsigned char *p = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
        pl.buf.resize(buffers.size());
        for (size_t i = 0; i < buffers.size(); ++i) {
            pl.buf[i] = p;
            p += buffer_bytes(buffers[i], max_rows, max_tokens);
        }
        if ((size_t)(p - static_cast<unsigned char *>(scratch)) > scratch_sz)
            return fail(B2S_ERR_INVALID, "graph: stream scratch too small");
        pl.amap.resize(ops.size());
        for (size_t i = 0; i < ops.size(); ++i) {
            const GOp &op = ops[i];
            if (op.opcode == OP_CONV) {   // im2col view of the NHWC input buffer, `max_rows` images
                B2S_TRY(make_tmap_im2col_nhwc(&pl.amap[i], pl.buf[op.a[0]], max_rows, op.a[8], op.a[9], op.a[10], op.a[11],
                                              op.a[12], op.a[13]));
                continue;
            }
            if (op.opcode == OP_CONV_STEM) {   // overlapping-stride view of the space-to-depth image, `max_rows` images
                const ConvGeom cg = make_stem_geom(op.a[10], op.a[11]);
                B2S_TRY(make_tmap_stem_s2d(&pl.amap[i], pl.buf[op.a[0]], max_rows, op.a[8], op.a[9], op.a[10], op.a[11],
                                           cg.rows_per_tile));
                continue;
            }
            if (op.opcode != OP_LINEAR) continue;
            const GBuffer &ab = buffers[op.a[0]];
            const int64_t rows = ab.rows_kind == 0 ? max_tokens : max_rows * (int64_t)ab.rows_kind;
            B2S_TRY(make_tmap_2d_kmajor(&pl.amap[i], pl.buf[op.a[0]], rows, op.a[7], ab.cols, 128, 0));
        }
        return 0;
    }

    int launch(cudaStream_t st, int64_t n_rows, const void *const *d_in, void *const *d_out,
               const int64_t *d_row_offsets, void *scratch, size_t scratch_sz, const LaunchInfo &li) override
    {
        if (n_rows <= 0) return 0;
        int64_t n_tokens = 0;
        int max_seqlen = 0;
        if (ragged) {
            if (!d_row_offsets || !li.h_row_offsets) return fail(B2S_ERR_INVALID, "graph: ragged batch needs row offsets");
            n_tokens = li.h_row_offsets[n_rows];
            for (int64_t r = 0; r < n_rows; ++r) {
                const int64_t s = li.h_row_offsets[r + 1] - li.h_row_offsets[r];
                if (s <= 0) return fail(B2S_ERR_INVALID, "graph: empty sequence in batch");
                if (s > (int64_t)h.max_pos) return fail(B2S_ERR_INVALID, "graph: sequence of %lld tokens exceeds the model's %u positions", (long long)s, h.max_pos);
                if (s > max_seqlen) max_seqlen = (int)s;
            }
        }
        Plan *pl;
        {
            std::lock_guard<std::mutex> l(mu);
            auto it = plans.find(scratch);
            if (it == plans.end()) {
                Plan fresh;
                B2S_TRY(build_plan(fresh, scratch, scratch_sz, li.max_rows, li.max_rows * (li.max_row_elems > 0 ? li.max_row_elems : 1)));
                it = plans.emplace(scratch, std::move(fresh)).first;
            }
            pl = &it->second;
        }
        if (ragged && n_tokens > li.max_rows * (li.max_row_elems > 0 ? li.max_row_elems : 1))
            return fail(B2S_ERR_INVALID, "graph: batch of %lld tokens exceeds the stream capacity", (long long)n_tokens);

        for (size_t i = 0; i < ops.size(); ++i) {
            const GOp &op = ops[i];
            switch (op.opcode) {
            case OP_EMBED_LN: {
                // a: in_ids, in_types(-1), W_word, W_pos, W_type, gamma, beta, out16, out32, H, vocab, max_pos, n_types
                B2S_TRY(embed_layernorm(st, static_cast<const int32_t *>(d_in[op.a[0]]),
                                        op.a[1] >= 0 ? static_cast<const int32_t *>(d_in[op.a[1]]) : nullptr, d_row_offsets,
                                        (int)n_rows, n_tokens, op.a[9], tptr(op.a[2]), tptr(op.a[3]), tptr(op.a[4]), op.a[10],
                                        op.a[11], op.a[12], static_cast<const float *>(tptr(op.a[5])),
                                        static_cast<const float *>(tptr(op.a[6])), op.f[0],
                                        op.a[7] >= 0 ? pl->buf[op.a[7]] : nullptr,
                                        op.a[8] >= 0 ? reinterpret_cast<float *>(pl->buf[op.a[8]]) : nullptr));
                break;
            }
            case OP_LINEAR: {
                // a: in_buf, W, bias(-1), residual_buf(-1), out_buf, act, N, K, out_f32
                const GBuffer &ab = buffers[op.a[0]];
                const int M = (int)(ab.rows_kind == 0 ? n_tokens : n_rows * (int64_t)ab.rows_kind);
                GemmEpilogue ep;
                ep.bias = tptr(op.a[2]);
                ep.residual = op.a[3] >= 0 ? pl->buf[op.a[3]] : nullptr;
                const bool to_output = op.a[4] < 0;  // -1 - k  => model output k
                ep.C = to_output ? d_out[-1 - op.a[4]] : pl->buf[op.a[4]];
                ep.ldc = op.a[6];
                ep.act = op.a[5];
                ep.out_f32 = op.a[8];
                ep.is_bf16 = 0;
                ep.act_after = op.a[9];
                const bool wide = gemm_prefer_bn256(M, op.a[6], op.a[7]);
                if (op.a[6] % 192 == 0 && gemm_prefer_bn192(M, op.a[6], op.a[7], op.a[8], op.a[5]))
                    B2S_TRY(gemm_tn_maps_2sm192(st, pl->amap[i], bmap96[i], M, op.a[6], op.a[7], ep));
                else if (wide && gemm_pair_enabled())
                    B2S_TRY(gemm_tn_maps_pair(st, pl->amap[i], bmap[i], M, op.a[6], op.a[7], ep));   // bmap: box 128 = half tile
                else if (wide)
                    B2S_TRY(gemm_tn_maps(st, pl->amap[i], bmap256[i], 256, M, op.a[6], op.a[7], ep));
                else
                    B2S_TRY(gemm_tn_maps(st, pl->amap[i], bmap[i], gemm_bn_for(op.a[6]), M, op.a[6], op.a[7], ep));
                break;
            }
            case OP_CONV: {
                // a: in_buf, W, bias(-1), residual_buf(-1), out_buf, act, N, K, H, W, Cin, KS, stride, pad, act_after
                const ConvGeom cg = make_conv_geom(op.a[8], op.a[9], op.a[10], op.a[11], op.a[12], op.a[13]);
                const int M = (int)(n_rows * (int64_t)cg.OH * cg.OW);
                GemmEpilogue ep;
                ep.bias = tptr(op.a[2]);
                ep.residual = op.a[3] >= 0 ? pl->buf[op.a[3]] : nullptr;
                ep.C = pl->buf[op.a[4]];
                ep.ldc = op.a[6];
                ep.act = op.a[5];
                ep.out_f32 = 0;
                ep.is_bf16 = 0;
                ep.act_after = op.a[14];
                const bool bn256 = gemm_prefer_bn256(M, op.a[6], 0);   // convolutions stay on the v3 pair kernel
                if (bn256 && gemm_pair_enabled())
                    B2S_TRY(conv_implicit_maps(st, pl->amap[i], bmap[i], 128, true, M, op.a[6], op.a[7], ep, cg));
                else if (bn256)
                    B2S_TRY(conv_implicit_maps(st, pl->amap[i], bmap256[i], 256, false, M, op.a[6], op.a[7], ep, cg));
                else
                    B2S_TRY(conv_implicit_maps(st, pl->amap[i], bmap[i], gemm_bn_for(op.a[6]), false, M, op.a[6], op.a[7], ep, cg));
                break;
            }
            case OP_CONV_STEM: {
                // a: z_buf, W'[N,256], bias(-1), -, out_buf, act, N, K (256), Hz, Wz, OH, OW
                const ConvGeom cg = make_stem_geom(op.a[10], op.a[11]);
                const int M = (int)(n_rows * (int64_t)cg.OH * cg.OW);
                GemmEpilogue ep;
                ep.bias = tptr(op.a[2]);
                ep.residual = nullptr;
                ep.C = pl->buf[op.a[4]];
                ep.ldc = op.a[6];
                ep.act = op.a[5];
                ep.out_f32 = 0;
                ep.is_bf16 = 0;
                ep.act_after = 0;
                B2S_TRY(conv_stem_maps(st, pl->amap[i], bmap[i], gemm_bn_for(op.a[6]), M, op.a[6], ep, cg));
                break;
            }
            case OP_STEM_S2D:
                // a: in_input, z_buf, C, H, W, Hz, Wz
                B2S_TRY(nchw_to_s2d(st, d_in[op.a[0]], h.in_dtype[op.a[0]], n_rows, op.a[2], op.a[3], op.a[4], op.a[5], op.a[6],
                                    pl->buf[op.a[1]]));
                break;
            case OP_LAYERNORM: {
                // a: in32_buf, gamma, beta, out16(-1), out32(-1), H
                const GBuffer &ib = buffers[op.a[0]];
                const int64_t rows = ib.rows_kind == 0 ? n_tokens : n_rows * (int64_t)ib.rows_kind;
                B2S_TRY(layernorm(st, reinterpret_cast<const float *>(pl->buf[op.a[0]]), rows, op.a[5],
                                  static_cast<const float *>(tptr(op.a[1])), static_cast<const float *>(tptr(op.a[2])), op.f[0],
                                  op.a[3] >= 0 ? pl->buf[op.a[3]] : nullptr,
                                  op.a[4] >= 0 ? reinterpret_cast<float *>(pl->buf[op.a[4]]) : nullptr));
                break;
            }
            case OP_ATTENTION: {
                // a: qkv_buf, mask_input(-1), out_buf, heads, head_dim
                B2S_TRY(attention_varlen(st, pl->buf[op.a[0]], d_row_offsets,
                                         op.a[1] >= 0 ? static_cast<const int32_t *>(d_in[op.a[1]]) : nullptr, pl->buf[op.a[2]],
                                         (int)n_rows, max_seqlen, op.a[3], op.a[4], n_tokens));
                break;
            }
            case OP_GATHER_FIRST: {
                // a: in_buf, out_buf, H
                B2S_TRY(gather_first(st, pl->buf[op.a[0]], d_row_offsets, (int)n_rows, op.a[2], pl->buf[op.a[1]]));
                break;
            }
            case OP_NCHW_TO_NHWC:
                // a: in_input, out_buf, C, H, W, Cp
                B2S_TRY(nchw_to_nhwc(st, d_in[op.a[0]], h.in_dtype[op.a[0]], n_rows, op.a[2], op.a[3], op.a[4], op.a[5], pl->buf[op.a[1]]));
                break;
            case OP_IM2COL:
                // a: in_buf, out_buf, H, W, C, KH, KW, stride, pad, OH, OW, Kp
                B2S_TRY(im2col_nhwc(st, pl->buf[op.a[0]], n_rows, op.a[2], op.a[3], op.a[4], op.a[5], op.a[6], op.a[7], op.a[8],
                                    op.a[9], op.a[10], op.a[11], pl->buf[op.a[1]]));
                break;
            case OP_MAXPOOL:
                // a: in_buf, out_buf, H, W, C, OH, OW
                B2S_TRY(maxpool3x3s2(st, pl->buf[op.a[0]], n_rows, op.a[2], op.a[3], op.a[4], op.a[5], op.a[6], pl->buf[op.a[1]]));
                break;
            case OP_AVGPOOL:
                // a: in_buf, out_buf, HW, C
                B2S_TRY(avgpool(st, pl->buf[op.a[0]], n_rows, op.a[2], op.a[3], pl->buf[op.a[1]]));
                break;
            default:
                return fail(B2S_ERR_INVALID, "graph: unknown opcode %u", op.opcode);
            }
        }
        return 0;
    }
};

}  // namespace

int graph_model_create(int device, const void *blob, size_t bytes, Model **out)
{
    if (bytes < sizeof(GHeader)) return fail(B2S_ERR_INVALID, "graph blob too small");
    GraphModel *m = new GraphModel();
    m->device = device;
    auto bail = [&](int code) {
        delete m;
        return code;
    };
    memcpy(&m->h, blob, sizeof(GHeader));
    const GHeader &h = m->h;
    if (memcmp(h.magic, "B2SG", 4) != 0 || h.version != 1) return bail(fail(B2S_ERR_INVALID, "graph blob: bad magic/version"));
    if (h.n_inputs == 0 || h.n_inputs > 4 || h.n_outputs == 0 || h.n_outputs > 4 || h.n_ops == 0)
        return bail(fail(B2S_ERR_INVALID, "graph blob: bad counts"));
    const size_t table_bytes = (size_t)h.n_tensors * sizeof(GTensor) + (size_t)h.n_buffers * sizeof(GBuffer) + (size_t)h.n_ops * sizeof(GOp);
    size_t data_off = (size_t)round_up((int64_t)(sizeof(GHeader) + table_bytes), 256);
    if (bytes < data_off) return bail(fail(B2S_ERR_INVALID, "graph blob truncated (tables)"));
    const unsigned char *p = static_cast<const unsigned char *>(blob) + sizeof(GHeader);
    m->tensors.resize(h.n_tensors);
    memcpy(m->tensors.data(), p, h.n_tensors * sizeof(GTensor));
    p += h.n_tensors * sizeof(GTensor);
    m->buffers.resize(h.n_buffers);
    memcpy(m->buffers.data(), p, h.n_buffers * sizeof(GBuffer));
    p += h.n_buffers * sizeof(GBuffer);
    m->ops.resize(h.n_ops);
    memcpy(m->ops.data(), p, h.n_ops * sizeof(GOp));
    size_t data_bytes = 0;
    for (const GTensor &t : m->tensors) {
        if (t.offset % 256 != 0 || data_off + t.offset + t.nbytes > bytes)
            return bail(fail(B2S_ERR_INVALID, "graph blob: tensor outside the data section"));
        if (t.offset + t.nbytes > data_bytes) data_bytes = t.offset + t.nbytes;
    }
    // validate op operands
    auto tok = [&](int idx, bool optional) { return (optional && idx < 0) || (idx >= 0 && idx < (int)h.n_tensors); };
    auto bok = [&](int idx, bool optional) { return (optional && idx < 0) || (idx >= 0 && idx < (int)h.n_buffers); };
    auto iok = [&](int idx, bool optional) { return (optional && idx < 0) || (idx >= 0 && idx < (int)h.n_inputs); };
    for (const GOp &op : m->ops) {
        bool ok = true;
        switch (op.opcode) {
        case OP_EMBED_LN:
            ok = iok(op.a[0], false) && iok(op.a[1], true) && tok(op.a[2], false) && tok(op.a[3], false) && tok(op.a[4], false) &&
                 tok(op.a[5], false) && tok(op.a[6], false) && bok(op.a[7], true) && bok(op.a[8], true);
            break;
        case OP_LINEAR:
            ok = bok(op.a[0], false) && tok(op.a[1], false) && tok(op.a[2], true) && bok(op.a[3], true) &&
                 (bok(op.a[4], false) || (op.a[4] < 0 && -1 - op.a[4] < (int)h.n_outputs)) && op.a[6] > 0 && op.a[7] > 0 &&
                 op.a[7] % 8 == 0;
            if (ok) {
                const GTensor &w = m->tensors[op.a[1]];
                ok = w.dtype == B2S_F16 && w.ndim == 2 && w.shape[0] == op.a[6] && w.shape[1] == op.a[7] &&
                     m->buffers[op.a[0]].dtype == B2S_F16 && m->buffers[op.a[0]].cols == op.a[7];
            }
            break;
        case OP_CONV: {
            ok = bok(op.a[0], false) && tok(op.a[1], false) && tok(op.a[2], true) && bok(op.a[3], true) && bok(op.a[4], false) &&
                 op.a[6] > 0 && op.a[6] % 8 == 0 && op.a[10] > 0 && op.a[10] % 64 == 0 && op.a[11] >= 1 && op.a[11] <= 7 &&
                 op.a[12] >= 1 && op.a[13] >= 0 && op.a[7] == op.a[11] * op.a[11] * op.a[10];
            if (ok) {
                const GTensor &w = m->tensors[op.a[1]];
                const int OH = (op.a[8] + 2 * op.a[13] - op.a[11]) / op.a[12] + 1, OW = (op.a[9] + 2 * op.a[13] - op.a[11]) / op.a[12] + 1;
                const GBuffer &ib = m->buffers[op.a[0]], &ob = m->buffers[op.a[4]];
                ok = w.dtype == B2S_F16 && w.ndim == 2 && w.shape[0] == op.a[6] && w.shape[1] == op.a[7] && ib.dtype == B2S_F16 &&
                     ib.cols == op.a[10] && (int64_t)ib.rows_kind == (int64_t)op.a[8] * op.a[9] && ob.dtype == B2S_F16 &&
                     ob.cols == op.a[6] && (int64_t)ob.rows_kind == (int64_t)OH * OW;
                if (ok && op.a[3] >= 0) {
                    const GBuffer &rb = m->buffers[op.a[3]];
                    ok = rb.dtype == B2S_F16 && rb.cols == op.a[6] && rb.rows_kind == ob.rows_kind;
                }
            }
            break;
        }
        case OP_STEM_S2D:
            ok = iok(op.a[0], false) && bok(op.a[1], false) && op.a[2] >= 1 && op.a[2] <= 4 && op.a[3] > 0 && op.a[4] > 0 &&
                 2 * op.a[5] >= op.a[3] + 6 && 2 * op.a[6] >= op.a[4] + 6 && m->buffers[op.a[1]].dtype == B2S_F16 &&
                 m->buffers[op.a[1]].cols == 16 && (int64_t)m->buffers[op.a[1]].rows_kind == (int64_t)op.a[5] * op.a[6] &&
                 h.in_row_elems[op.a[0]] == (int64_t)op.a[2] * op.a[3] * op.a[4];
            break;
        case OP_CONV_STEM: {
            ok = bok(op.a[0], false) && tok(op.a[1], false) && tok(op.a[2], true) && bok(op.a[4], false) && op.a[6] > 0 &&
                 op.a[6] % 8 == 0 && op.a[6] <= 128 && op.a[7] == 256 && op.a[10] > 0 && op.a[11] > 0 && op.a[11] <= 128 &&
                 op.a[10] + 3 <= op.a[8] && op.a[11] + 3 <= op.a[9];
            if (ok) {
                const GTensor &w = m->tensors[op.a[1]];
                const GBuffer &ib = m->buffers[op.a[0]], &ob = m->buffers[op.a[4]];
                ok = w.dtype == B2S_F16 && w.ndim == 2 && w.shape[0] == op.a[6] && w.shape[1] == 256 && ib.dtype == B2S_F16 &&
                     ib.cols == 16 && (int64_t)ib.rows_kind == (int64_t)op.a[8] * op.a[9] && ob.dtype == B2S_F16 &&
                     ob.cols == op.a[6] && (int64_t)ob.rows_kind == (int64_t)op.a[10] * op.a[11];
            }
            break;
        }
        case OP_LAYERNORM:
            ok = bok(op.a[0], false) && tok(op.a[1], false) && tok(op.a[2], false) && bok(op.a[3], true) && bok(op.a[4], true);
            break;
        case OP_ATTENTION:
            ok = bok(op.a[0], false) && iok(op.a[1], true) && bok(op.a[2], false);
            break;
        case OP_GATHER_FIRST:
            ok = bok(op.a[0], false) && bok(op.a[1], false);
            break;
        case OP_NCHW_TO_NHWC:
            ok = iok(op.a[0], false) && bok(op.a[1], false) && op.a[5] % 8 == 0 && op.a[5] >= op.a[2] &&
                 m->buffers[op.a[1]].cols == op.a[5] && (int64_t)m->buffers[op.a[1]].rows_kind == (int64_t)op.a[3] * op.a[4];
            break;
        case OP_IM2COL:
            ok = bok(op.a[0], false) && bok(op.a[1], false) && op.a[4] % 8 == 0 && op.a[11] % 8 == 0 &&
                 op.a[11] >= op.a[5] * op.a[6] * op.a[4] && m->buffers[op.a[1]].cols == op.a[11] &&
                 (int64_t)m->buffers[op.a[1]].rows_kind == (int64_t)op.a[9] * op.a[10] &&
                 (int64_t)m->buffers[op.a[0]].rows_kind == (int64_t)op.a[2] * op.a[3] && m->buffers[op.a[0]].cols == op.a[4];
            break;
        case OP_MAXPOOL:
            ok = bok(op.a[0], false) && bok(op.a[1], false) && op.a[4] % 8 == 0 &&
                 (int64_t)m->buffers[op.a[1]].rows_kind == (int64_t)op.a[5] * op.a[6] && m->buffers[op.a[1]].cols == op.a[4];
            break;
        case OP_AVGPOOL:
            ok = bok(op.a[0], false) && bok(op.a[1], false) && op.a[3] % 8 == 0 &&
                 (int64_t)m->buffers[op.a[0]].rows_kind == (int64_t)op.a[2] && m->buffers[op.a[1]].rows_kind == 1;
            break;
        default:
            ok = false;
        }
        if (!ok) return bail(fail(B2S_ERR_INVALID, "graph blob: malformed op (opcode %u)", op.opcode));
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_data), data_bytes ? data_bytes : 256);
    if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMalloc(graph weights)"));
    e = cudaMemcpy(m->d_data, static_cast<const unsigned char *>(blob) + data_off, data_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail_cuda(e, "cudaMemcpy(graph weights)"));
    // weight tensor maps
    m->bmap.resize(m->ops.size());
    m->bmap256.resize(m->ops.size());
    m->bmap96.resize(m->ops.size());
    int64_t flops_fixed = 0;
    for (size_t i = 0; i < m->ops.size(); ++i) {
        const GOp &op = m->ops[i];
        if (op.opcode != OP_LINEAR && op.opcode != OP_CONV && op.opcode != OP_CONV_STEM) continue;
        int rc = make_tmap_2d_kmajor(&m->bmap[i], m->tptr(op.a[1]), op.a[6], op.a[7], op.a[7], gemm_bn_for(op.a[6]), 0);
        if (rc == 0 && op.a[6] >= 256)
            rc = make_tmap_2d_kmajor(&m->bmap256[i], m->tptr(op.a[1]), op.a[6], op.a[7], op.a[7], 256, 0);
        if (rc == 0 && op.opcode == OP_LINEAR && op.a[8] && op.a[6] % 192 == 0)
            rc = make_tmap_2d_kmajor(&m->bmap96[i], m->tptr(op.a[1]), op.a[6], op.a[7], op.a[7], 96, 0);
        if (rc != 0) return bail(rc);
        flops_fixed += 2LL * op.a[6] * op.a[7];
    }
    b2s_model_info &info = m->info;
    info.kind = B2S_MODEL_GRAPH;
    info.n_inputs = (int32_t)h.n_inputs;
    info.n_outputs = (int32_t)h.n_outputs;
    for (uint32_t i = 0; i < h.n_inputs; ++i) {
        info.in_dtype[i] = h.in_dtype[i];
        info.in_row_elems[i] = h.in_row_elems[i];  // -1: one variable-length row per sequence
        if (h.in_row_elems[i] < 0) m->ragged = true;
    }
    for (uint32_t o = 0; o < h.n_outputs; ++o) {
        const GBuffer &ob = m->buffers[h.out_buffer[o]];
        info.out_dtype[o] = (int32_t)ob.dtype;
        info.out_row_elems[o] = ob.cols;
    }
    info.weight_bytes = (int64_t)data_bytes;
    info.algo_bytes_fixed = (int64_t)data_bytes;
    info.algo_bytes_per_row = flops_fixed;  // GEMM FLOPs per token (linear part); see DESIGN.md
    *out = m;
    return 0;
}

}  // namespace b2s
