// linear.cu -- linear / logistic decision function + label pick (kernel K4 of SURVEY.md 2.2).
//
// Replaces `self._model.predict(data)` of the reference's sklearn engine for linear classifiers
// (clearml_serving/serving/preprocess_service.py:459-464; BASELINE.json configs[0]):
//     scores = X . coef^T + intercept   (fp64)
//     label  = classes[ n_out == 1 ? (score > 0) : argmax(scores) ]
// fp64 multiplies and adds are kept un-fused (__dmul_rn / __dadd_rn, k ascending), the same
// arithmetic the oracle (oracle/forest_oracle.c: oracle_linear_predict) performs.
// One thread per row: the op is a few dozen flops per row and HBM-trivial; what matters is that
// it rides the same collate -> H2D -> kernel -> D2H -> scatter path as the other models.
#include "common.cuh"

#include <string.h>

namespace b2s {

struct LinearBlobHeader {
    char magic[4];  // "B2SL"
    uint32_t version;
    uint32_t n_features, n_out, n_classes, reserved;
};
static_assert(sizeof(LinearBlobHeader) == 24, "blob header layout");

__global__ void __launch_bounds__(128)
linear_predict_kernel(const double *__restrict__ X, int64_t n_rows, int n_features,
                      const double *__restrict__ W, const double *__restrict__ b, int n_out,
                      const int64_t *__restrict__ classes, int64_t *__restrict__ labels,
                      double *__restrict__ scores)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const double *x = X + i * n_features;
    int best = 0;
    double best_s = 0.0;
    for (int c = 0; c < n_out; ++c) {
        double s = 0.0;
        const double *w = W + (int64_t)c * n_features;
        for (int k = 0; k < n_features; ++k) s = __dadd_rn(s, __dmul_rn(x[k], __ldg(w + k)));
        s = __dadd_rn(s, __ldg(b + c));
        scores[i * n_out + c] = s;
        if (c == 0 || s > best_s) { best = c; best_s = s; }
    }
    const int idx = (n_out == 1) ? (best_s > 0.0 ? 1 : 0) : best;
    labels[i] = __ldg(classes + idx);
}

namespace {
struct LinearModel : Model {
    void *d_blob = nullptr;
    const double *W = nullptr, *b = nullptr;
    const int64_t *classes = nullptr;
    int n_features = 0, n_out = 0;
    ~LinearModel() override
    {
        if (d_blob) { cudaSetDevice(device); cudaFree(d_blob); }
    }
    size_t scratch_bytes(int64_t, int64_t) const override { return 256; }
    int launch(cudaStream_t st, int64_t n_rows, const void *const *d_in, void *const *d_out,
               const int64_t *, void *, size_t, const LaunchInfo &) override
    {
        if (n_rows <= 0) return 0;
        const unsigned grid = (unsigned)((n_rows + 127) / 128);
        linear_predict_kernel<<<grid, 128, 0, st>>>(static_cast<const double *>(d_in[0]), n_rows, n_features,
                                                    W, b, n_out, classes, static_cast<int64_t *>(d_out[0]),
                                                    static_cast<double *>(d_out[1]));
        count_launch();
        B2S_CUDA(cudaGetLastError());
        return 0;
    }
};
}  // namespace

int linear_model_create(int device, const void *blob, size_t bytes, Model **out)
{
    if (bytes < sizeof(LinearBlobHeader)) return fail(B2S_ERR_INVALID, "linear blob too small");
    LinearBlobHeader h;
    memcpy(&h, blob, sizeof(h));
    if (memcmp(h.magic, "B2SL", 4) != 0 || h.version != 1) return fail(B2S_ERR_INVALID, "linear blob: bad magic/version");
    if (h.n_features == 0 || h.n_out == 0) return fail(B2S_ERR_INVALID, "linear blob: empty model");
    const uint32_t need_classes = h.n_out == 1 ? 2u : h.n_out;
    if (h.n_classes != need_classes) return fail(B2S_ERR_INVALID, "linear blob: n_classes %u does not match n_out %u", h.n_classes, h.n_out);
    const size_t wb = (size_t)h.n_out * h.n_features * 8, bb = (size_t)h.n_out * 8, cb = (size_t)h.n_classes * 8;
    if (bytes < sizeof(h) + wb + bb + cb) return fail(B2S_ERR_INVALID, "linear blob truncated");
    LinearModel *m = new LinearModel();
    m->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, wb + bb + cb);
    if (e != cudaSuccess) { delete m; return fail_cuda(e, "cudaMalloc(linear)"); }
    e = cudaMemcpy(m->d_blob, static_cast<const unsigned char *>(blob) + sizeof(h), wb + bb + cb, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete m; return fail_cuda(e, "cudaMemcpy(linear)"); }
    unsigned char *d = static_cast<unsigned char *>(m->d_blob);
    m->W = reinterpret_cast<const double *>(d);
    m->b = reinterpret_cast<const double *>(d + wb);
    m->classes = reinterpret_cast<const int64_t *>(d + wb + bb);
    m->n_features = (int)h.n_features;
    m->n_out = (int)h.n_out;
    b2s_model_info &info = m->info;
    info.kind = B2S_MODEL_LINEAR;
    info.n_inputs = 1;
    info.n_outputs = 2;
    info.in_dtype[0] = B2S_F64;
    info.in_row_elems[0] = h.n_features;
    info.out_dtype[0] = B2S_I64;
    info.out_row_elems[0] = 1;
    info.out_dtype[1] = B2S_F64;
    info.out_row_elems[1] = h.n_out;
    info.weight_bytes = (int64_t)(wb + bb + cb);
    info.algo_bytes_fixed = (int64_t)(wb + bb);
    info.algo_bytes_per_row = (int64_t)h.n_features * 8 + 8 + (int64_t)h.n_out * 8;
    *out = m;
    return 0;
}

}  // namespace b2s
