"""Time the REFERENCE's own serving path on the host cores (test / measurement infrastructure only).

What runs is the unmodified reference code -- `ModelRequestProcessor.process_request`
(clearml_serving/serving/model_request_processor.py:253-304) -> `_process_request` (:1309-1369) ->
`SKLearnPreprocessRequest.process` (preprocess_service.py:459-464) -> sklearn `predict` -- imported under the
stubs of oracle/ref_harness.py from `baseline/_ref` (the pip --target install of /root/reference that travels to
the GPU box) or from /root/reference itself.  The xgboost engine cannot run (xgboost is not installable), so the
tree workload uses the reference's sklearn engine on a GradientBoostingRegressor of the BASELINE configs[1] shape
(1000 trees x depth 6 x 32 features) -- the model behind tests/golden/sk_gbr_cfg2.npz, re-fitted with the same seed and
checked bit-for-bit against that golden before it is timed.

bench.py's cpu legs are the only callers.  Nothing here is imported by the product package.
"""
import asyncio
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class _TreePreprocess(object):
    """examples/xgboost/preprocess.py:12-19 widened to 32 features (BASELINE configs[1])"""

    def preprocess(self, body, state, collect_custom_statistics_fn=None):
        return [[body.get("x{}".format(i), None) for i in range(32)]]

    def postprocess(self, data, state, collect_custom_statistics_fn=None):
        return dict(y=data.tolist() if isinstance(data, np.ndarray) else data)


class _IrisPreprocess(object):
    """examples/sklearn/preprocess.py:12-19 widened from x0,x1 to x0..x3 (BASELINE configs[0])"""

    def preprocess(self, body, state, collect_custom_statistics_fn=None):
        return [[body.get("x0", None), body.get("x1", None), body.get("x2", None), body.get("x3", None)]]

    def postprocess(self, data, state, collect_custom_statistics_fn=None):
        return dict(y=data.tolist() if isinstance(data, np.ndarray) else data)


def fit_cfg2_gbr():
    """the GradientBoostingRegressor of oracle/gen_golden.py:gen_trees_cfg2 (same data, same seed)"""
    from sklearn.ensemble import GradientBoostingRegressor
    rng = np.random.default_rng(11)
    F = 32
    Xtr = rng.standard_normal((3000, F))
    ytr = (Xtr[:, 0] * 2 + np.sin(Xtr[:, 1] * 3) + Xtr[:, 2] * Xtr[:, 3] + np.abs(Xtr[:, 4:12]).sum(1) * 0.3
           + 0.5 * rng.standard_normal(3000))
    return GradientBoostingRegressor(n_estimators=1000, max_depth=6, learning_rate=0.05, subsample=0.5,
                                     random_state=0).fit(Xtr, ytr)


def fit_iris_lr():
    from sklearn.datasets import load_iris
    from sklearn.linear_model import LogisticRegression
    Xtr, ytr = load_iris(return_X_y=True)
    return LogisticRegression(max_iter=1000).fit(Xtr, ytr)


def golden_check(gbr):
    """True when the re-fitted model reproduces tests/golden/sk_gbr_cfg2.npz bit for bit (same sklearn build)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "sk_gbr_cfg2.npz"))
    return bool(np.array_equal(gbr.predict(g["X"]), g["y"]))


def make_reference_processor(model, kind):
    """(ref modules, processor, url) with the reference's sklearn engine around `model`"""
    from oracle import ref_harness as rh
    ref = rh.load_reference()
    pre = _TreePreprocess() if kind == "trees" else _IrisPreprocess()
    ep = ref.endpoints.ModelEndpoint(engine_type="sklearn", serving_url="bench_ref")
    eng = rh.make_engine(ref, ref.ps.SKLearnPreprocessRequest, ep, model=model, preprocess=pre)
    return ref, rh.make_processor(ref, {"bench_ref": ep}, {"bench_ref": eng}), "bench_ref"


def bodies_for(kind, n=512, seed=5):
    rng = np.random.default_rng(seed)
    if kind == "trees":
        X = rng.standard_normal((n, 32)).astype(np.float32)
    else:
        X = rng.uniform(0, 8, (n, 4))
    return X, [{"x{}".format(j): float(X[i, j]) for j in range(X.shape[1])} for i in range(n)]


def closed_loop(proc, url, bodies, seconds, warmup=50):
    """serial closed loop through process_request (the reference's sklearn engine is synchronous: one request at a time
    per worker process, entrypoint.sh:47-73 scales by processes); returns req/s, p50 / p99 latency, last reply"""
    lat = []

    async def run():
        for i in range(warmup):
            await proc.process_request(base_url=url, version=None, request_body=bodies[i % len(bodies)], serve_type="process")
        t_end = time.perf_counter() + seconds
        i = 0
        reply = None
        while time.perf_counter() < t_end:
            t = time.perf_counter()
            reply = await proc.process_request(base_url=url, version=None, request_body=bodies[i % len(bodies)], serve_type="process")
            lat.append(time.perf_counter() - t)
            i += 1
        return reply
    t0 = time.perf_counter()
    reply = asyncio.run(run())
    a = np.asarray(lat) * 1e6
    return dict(req_s=len(lat) / float(np.sum(lat)), p50_us=float(np.percentile(a, 50)), p99_us=float(np.percentile(a, 99)),
                completed=len(lat), timed_s=float(np.sum(lat)), wall_s=time.perf_counter() - t0), reply


def multi_process(kind, model, n_proc, seconds):
    """N worker processes, each with its own copy of the model and its own event loop, as the reference deploys them
    (serving/entrypoint.sh:47-73, CLEARML_SERVING_NUM_PROCESS).  Returns aggregate req/s."""
    import json
    import subprocess
    import sys
    import tempfile

    import joblib
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.pkl")
        joblib.dump(model, path)
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="1")
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, "-m", "oracle.ref_bench", kind, path, str(seconds)], cwd=ROOT, env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(n_proc)]
        out = []
        for pr in procs:
            line = pr.communicate(timeout=seconds + 120)[0].strip().splitlines()
            if pr.returncode == 0 and line:
                out.append(json.loads(line[-1]))
        wall = time.perf_counter() - t0
    if not out:
        return dict(error="no worker process finished")
    done = sum(o["completed"] for o in out)
    return dict(req_s=sum(o["completed"] / o["timed_s"] for o in out), processes=len(out), completed=done, wall_s=wall)


if __name__ == "__main__":   # worker process of multi_process()
    import json
    import sys

    import joblib
    _kind, _path, _seconds = sys.argv[1], sys.argv[2], float(sys.argv[3])
    _ref, _proc, _url = make_reference_processor(joblib.load(_path), _kind)
    _X, _bodies = bodies_for(_kind)
    _res, _ = closed_loop(_proc, _url, _bodies, _seconds)
    print(json.dumps(dict(completed=_res["completed"], timed_s=_res["timed_s"])))
