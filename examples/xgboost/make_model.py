"""Synthetic regressor in XGBoost's JSON model schema (what `Booster.save_model("m.json")` writes): complete depth-6 trees,
32 features, reg:squarederror, base_score 0.5 -- BASELINE.json configs[1]; `--trees N` for a smaller one."""
import json
import os
import sys

import numpy as np


def tree(rng, depth, n_features):
    n = 2 ** (depth + 1) - 1
    first_leaf = 2 ** depth - 1
    left = [2 * i + 1 if i < first_leaf else -1 for i in range(n)]
    right = [2 * i + 2 if i < first_leaf else -1 for i in range(n)]
    feat = [int(rng.integers(0, n_features)) if i < first_leaf else 0 for i in range(n)]
    cond = [float(np.float32(rng.normal(0, 1))) if i < first_leaf else float(np.float32(rng.normal(0, 0.1))) for i in range(n)]
    dl = [int(rng.random() < 0.5) if i < first_leaf else 0 for i in range(n)]
    return dict(left_children=left, right_children=right, split_indices=feat, split_conditions=cond, default_left=dl,
                split_type=[0] * n, categories=[], base_weights=[0.0] * n)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_trees = int(sys.argv[sys.argv.index("--trees") + 1]) if "--trees" in sys.argv else 1000
    out = args[0] if args else "."
    if os.path.isdir(out) or not out.endswith(".json"):
        os.makedirs(out, exist_ok=True)
        out = os.path.join(out, "xgb_model.json")
    rng = np.random.default_rng(0)
    model = {"learner": {"objective": {"name": "reg:squarederror"},
                         "learner_model_param": {"base_score": "5E-1", "num_feature": "32", "num_class": "0"},
                         "gradient_booster": {"name": "gbtree", "model": {"trees": [tree(rng, 6, 32) for _ in range(n_trees)]}}},
             "version": [1, 7, 5]}
    with open(out, "wt") as f:
        json.dump(model, f)
    print(out)


if __name__ == "__main__":
    main()
