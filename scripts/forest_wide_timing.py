"""Developer aid: phase stamps (SM clock) of forest_wide_kernel + CUDA-event timing, L2-cold and warm, at the
BASELINE configs[1] shape.   python scripts/forest_wide_timing.py   (B2S_FOREST_WIDE_C=8|4, B2S_FOREST_WIDE=0 to compare)"""
import ctypes
import os
import sys

import numpy as np

os.environ["B2S_FOREST_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clearml_serving_b200 import formats, native  # noqa: E402
from oracle import oracle as orc  # noqa: E402

native.ensure_init(0)
rng = np.random.default_rng(0)


def run(n_trees, rows, reps=200):
    f = orc.synth_xgb_forest(n_trees, 6, 32, seed=0)
    pm = formats.pack_forest(f, "xgb", base=0.5)
    m = native.Model(pm.kind, pm.blob, 0)
    s = native.Stream(m, rows, 0, 2)
    X = rng.standard_normal((rows, 32)).astype(np.float32)
    di = native.DeviceBuffer(X.nbytes); di.upload(X)
    do = native.DeviceBuffer(rows * 4)
    t = native.Timer(s)
    for _ in range(10):
        s.infer_device(rows, [di.ptr], [do.ptr])
    s.synchronize()
    ok = np.array_equal(do.download(np.float32, rows), orc.forest_predict_xgb(f, X, 0.5))
    cold = []
    for _ in range(reps):
        s.flush_l2(); t.start(); s.infer_device(rows, [di.ptr], [do.ptr]); t.stop(); cold.append(t.elapsed_ms())
    b = (ctypes.c_longlong * 64)()
    native.check(native.lib().b2s_debug_read(m.handle, b))
    st = np.array(b[:]).reshape(8, 8)
    t.start()
    for _ in range(reps):
        s.infer_device(rows, [di.ptr], [do.ptr])
    t.stop()
    warm = t.elapsed_ms() / reps
    native.check(native.lib().b2s_debug_read(m.handle, b))
    w0 = np.array(b[:]).reshape(8, 8)[0]
    print("   warm cta0: load(blk0)=%d traverse=%d wait1=%d publish+rest=%d own_wait=%d chain=%d total=%d" % (
        w0[1] - w0[0], w0[5] - w0[1], w0[6] - w0[5], w0[2] - w0[6], w0[3] - w0[2], w0[4] - w0[3], w0[4] - w0[0]))
    s0 = st[0]
    print("   cold cta0: load(blk0)=%d traverse=%d wait1=%d publish+rest=%d own_wait=%d chain=%d total=%d" % (
        s0[1] - s0[0], s0[5] - s0[1], s0[6] - s0[5], s0[2] - s0[6], s0[3] - s0[2], s0[4] - s0[3], s0[4] - s0[0]))
    print("trees=%d rows=%d ok=%s cold median=%.2fus p10=%.2f warm(back-to-back)=%.2fus | cta0 (cold) load=%d traverse+publish=%d barrier=%d chain=%d total=%d cycles" % (
        n_trees, rows, ok, np.median(cold) * 1e3, np.percentile(cold, 10) * 1e3, warm * 1e3,
        s0[1] - s0[0], s0[2] - s0[1], s0[3] - s0[2], s0[4] - s0[3], s0[4] - s0[0]))
    for r in range(1, 4):
        print("   cta%d: load=%d traverse+publish=%d barrier=%d chain=%d" % (r, st[r][1] - st[r][0], st[r][2] - st[r][1], st[r][3] - st[r][2], st[r][4] - st[r][3]))
    di.free(); do.free(); t.destroy(); s.destroy(); m.free()


for nt, rows in ((1000, 64), (1000, 256), (250, 64), (4000, 64)):
    run(nt, rows)
