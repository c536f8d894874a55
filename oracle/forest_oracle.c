/*
 * oracle/forest_oracle.c -- CPU restatement of the reference's tree / linear model prediction.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product path (clearml_serving_b200) never does.
 *
 * What it restates (the arithmetic lives in third-party libraries the reference calls):
 *
 *  (1) engine=xgboost  -- reference call site clearml_serving/serving/preprocess_service.py:478-483
 *      (`Booster.predict(DMatrix)`), model written by examples/xgboost/train_model.py:14-28
 *      (objective reg:squarederror).  xgboost is pinned `>=1.7.5,<1.8`
 *      (clearml_serving/serving/requirements.txt:16) and is NOT installable here, so this is a
 *      restatement of its published CPU predictor algorithm:
 *          margin = base_score;  for t in trees (in order): margin += leaf_t(x)   [all fp32]
 *          traversal: missing(NaN) -> default child, else (x[f] < split_cond ? left : right)
 *          reg:squarederror => prediction = margin.
 *      PARITY UNPINNED for this mode: the reference ships no golden vectors and xgboost cannot be
 *      imported in the build container.  The traversal code below is shared with mode (2), which
 *      IS pinned against the reference's own engine class.
 *
 *  (2) engine=sklearn with a tree ensemble -- reference call site preprocess_service.py:459-464
 *      (`self._model.predict(data)`), e.g. examples/ensemble/train_ensemble.py:16-21.
 *      sklearn GradientBoostingRegressor: raw = init (fp64); for each stage in order
 *      raw += learning_rate * value[leaf] (fp64, two roundings), traversal
 *      `X[i, feature] <= threshold` with X fp32 and threshold fp64.
 *      sklearn RandomForestRegressor: y = 0; y += value[leaf] per tree in order; y /= n_trees.
 *      PINNED: tests/golden/sk_gbr.npz and sk_rf.npz were produced by running the reference's
 *      SKLearnPreprocessRequest.process (oracle/gen_golden.py) and this file reproduces them
 *      bit for bit (tests/test_oracle.py).
 *
 *  (3) engine=sklearn LogisticRegression / linear models -- fp64 X.W^T + b then argmax / >0
 *      (BASELINE.json configs[0]).  PINNED by tests/golden/lr_iris.npz on the predicted labels.
 *
 * Build: `make -C oracle` -> oracle/liboracle.so  (plain C, -O2, no -ffast-math, no FMA
 * contraction: -ffp-contract=off so the fp32/fp64 roundings are exactly the ones written here).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Structure-of-arrays forest, node ids are global (tree_offset[t] = first node of tree t).
 * left[i] < 0  => node i is a leaf and value[i] is its output.
 * Children ids stored in left/right are LOCAL to the tree (as in both the XGBoost JSON schema and
 * sklearn's tree_.children_left). */
typedef struct {
    int32_t n_trees;
    int32_t n_features;
    const int32_t *tree_offset; /* [n_trees+1] */
    const int32_t *left;        /* [n_nodes] */
    const int32_t *right;       /* [n_nodes] */
    const int32_t *feat;        /* [n_nodes] */
    const double *thr;          /* [n_nodes] split threshold (fp64 container) */
    const uint8_t *default_left;/* [n_nodes] missing-value direction */
    const double *value;        /* [n_nodes] leaf value (fp64 container) */
} oracle_forest;

/* mode 0: XGBoost  -- go left iff (float)x < (float)thr ; NaN -> default_left
 * mode 1: sklearn  -- go left iff (double)x <= thr      ; NaN -> default_left */
static inline int32_t leaf_of(const oracle_forest *f, int32_t t, const float *x, int mode)
{
    const int32_t base = f->tree_offset[t];
    int32_t nid = 0;
    while (f->left[base + nid] >= 0) {
        const int32_t g = base + nid;
        const float xv = x[f->feat[g]];
        int go_left;
        if (isnan(xv)) {
            go_left = f->default_left[g] != 0;
        } else if (mode == 0) {
            go_left = xv < (float)f->thr[g];
        } else {
            go_left = (double)xv <= f->thr[g];
        }
        nid = go_left ? f->left[g] : f->right[g];
    }
    return base + nid;
}

/* XGBoost gbtree, single output group, identity link.  out[i] fp32.  */
void oracle_forest_predict_xgb(const oracle_forest *f, const float *X, int64_t n_rows,
                               float base_score, float *out, int n_threads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (n_threads != 1) num_threads(n_threads > 0 ? n_threads : omp_get_num_procs())
#endif
    for (int64_t i = 0; i < n_rows; ++i) {
        const float *x = X + i * f->n_features;
        float acc = base_score;
        for (int32_t t = 0; t < f->n_trees; ++t) {
            acc += (float)f->value[leaf_of(f, t, x, 0)];
        }
        out[i] = acc;
    }
}

/* sklearn tree ensembles: acc = init; acc += scale*value (scale*value rounded, then the add
 * rounded); out = acc / divisor.  GBR: init=DummyRegressor mean, scale=learning_rate, divisor=1.
 * RF: init=0, scale=1, divisor=n_trees.  */
void oracle_forest_predict_f64(const oracle_forest *f, const float *X, int64_t n_rows,
                               double init, double scale, double divisor, double *out,
                               int n_threads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (n_threads != 1) num_threads(n_threads > 0 ? n_threads : omp_get_num_procs())
#endif
    for (int64_t i = 0; i < n_rows; ++i) {
        const float *x = X + i * f->n_features;
        double acc = init;
        for (int32_t t = 0; t < f->n_trees; ++t) {
            const double v = scale * f->value[leaf_of(f, t, x, 1)];
            acc = acc + v;
        }
        out[i] = acc / divisor;
    }
}

/* Linear decision function, fp64: scores[i,c] = sum_k X[i,k]*W[c,k] (k ascending) + b[c].
 * labels[i] = n_classes==1 ? (score>0) : argmax_c (first max wins, as numpy argmax).  */
void oracle_linear_predict(const double *X, int64_t n_rows, int32_t n_features,
                           const double *W, const double *b, int32_t n_out,
                           double *scores, int64_t *labels)
{
    for (int64_t i = 0; i < n_rows; ++i) {
        int64_t best = 0;
        double best_s = 0.0;
        for (int32_t c = 0; c < n_out; ++c) {
            double s = 0.0;
            for (int32_t k = 0; k < n_features; ++k) {
                s = s + X[i * n_features + k] * W[(int64_t)c * n_features + k];
            }
            s = s + b[c];
            scores[i * n_out + c] = s;
            if (c == 0 || s > best_s) { best = c; best_s = s; }
        }
        labels[i] = (n_out == 1) ? (scores[i] > 0.0 ? 1 : 0) : best;
    }
}

/* ---- compact (array-of-structs) form of the same predictor, used for TIMING the CPU baseline ----
 * Same arithmetic as oracle_forest_predict_xgb (tests assert bit-equality); the node is packed the
 * way xgboost's RegTree::Node is (children, split index + default bit, fp32 condition/leaf), so
 * the timed CPU arm is not handicapped by the readable SoA layout above. */
typedef struct {
    int32_t left, right;   /* local child ids, left < 0 => leaf */
    uint32_t feat_dl;      /* feat | default_left << 31 */
    float cond;            /* threshold, or leaf value */
} oracle_cnode;

typedef struct {
    int32_t n_trees, n_features;
    int32_t *tree_offset;
    oracle_cnode *nodes;
} oracle_compiled;

#include <stdlib.h>

oracle_compiled *oracle_forest_compile(const oracle_forest *f)
{
    oracle_compiled *c = (oracle_compiled *)malloc(sizeof(*c));
    const int32_t n_nodes = f->tree_offset[f->n_trees];
    c->n_trees = f->n_trees;
    c->n_features = f->n_features;
    c->tree_offset = (int32_t *)malloc(sizeof(int32_t) * (f->n_trees + 1));
    c->nodes = (oracle_cnode *)malloc(sizeof(oracle_cnode) * n_nodes);
    for (int32_t t = 0; t <= f->n_trees; ++t) c->tree_offset[t] = f->tree_offset[t];
    for (int32_t i = 0; i < n_nodes; ++i) {
        c->nodes[i].left = f->left[i];
        c->nodes[i].right = f->right[i];
        c->nodes[i].feat_dl = (uint32_t)f->feat[i] | (f->default_left[i] ? 0x80000000u : 0u);
        c->nodes[i].cond = f->left[i] >= 0 ? (float)f->thr[i] : (float)f->value[i];
    }
    return c;
}

void oracle_compiled_free(oracle_compiled *c)
{
    if (!c) return;
    free(c->tree_offset);
    free(c->nodes);
    free(c);
}

void oracle_compiled_predict_xgb(const oracle_compiled *c, const float *X, int64_t n_rows,
                                 float base_score, float *out, int n_threads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (n_threads != 1) num_threads(n_threads > 0 ? n_threads : omp_get_num_procs())
#endif
    for (int64_t i = 0; i < n_rows; ++i) {
        const float *x = X + i * c->n_features;
        float acc = base_score;
        for (int32_t t = 0; t < c->n_trees; ++t) {
            const oracle_cnode *nodes = c->nodes + c->tree_offset[t];
            const oracle_cnode *n = nodes;
            while (n->left >= 0) {
                const float xv = x[n->feat_dl & 0x7fffffffu];
                const int go_left = isnan(xv) ? (int)(n->feat_dl >> 31) : (xv < n->cond);
                n = nodes + (go_left ? n->left : n->right);
            }
            acc += n->cond;
        }
        out[i] = acc;
    }
}

/* xgboost's logistic link and its inverse, as the CPU library computes them in fp32
 * (src/common/math.h Sigmoid: 1 / (expf(min(-x, 88.7f)) + 1 + 1e-16f); src/objective/regression_loss.h
 * LogisticRegression::ProbToMargin: -logf(1/base_score - 1)).  Third-party (xgboost >=1.7.5,<1.8,
 * clearml_serving/serving/requirements.txt:16), restated: PARITY UNPINNED like the rest of the xgboost mode. */
void oracle_xgb_sigmoid(float *y, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        float x = -y[i];
        if (x > 88.7f) x = 88.7f;
        const float denom = expf(x) + 1.0f + 1e-16f;
        y[i] = 1.0f / denom;
    }
}

float oracle_xgb_prob_to_margin(float base_score)
{
    return -logf(1.0f / base_score - 1.0f);
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
