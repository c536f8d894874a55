"""LogisticRegression on the iris data bundled with scikit-learn -> joblib pickle (BASELINE.json configs[0])."""
import os
import sys

import joblib
from sklearn.datasets import load_iris
from sklearn.linear_model import LogisticRegression

out = sys.argv[1] if len(sys.argv) > 1 else "."
if os.path.isdir(out) or not out.endswith(".pkl"):
    os.makedirs(out, exist_ok=True)
    out = os.path.join(out, "sklearn_iris.pkl")
X, y = load_iris(return_X_y=True)
joblib.dump(LogisticRegression(max_iter=1000).fit(X, y), out)
print(out)
