"""Developer aid: phase breakdown of forest_cluster_kernel (SM-clock stamps) + event timing for a few
batch sizes.  B2S_FOREST_TIMING=1 python scripts/forest_timing.py"""
import ctypes
import os
import sys

import numpy as np

os.environ["B2S_FOREST_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clearml_serving_b200 import formats, native  # noqa: E402
from oracle import oracle as orc  # noqa: E402

native.ensure_init(0)
forest = orc.synth_xgb_forest(1000, 6, 32, seed=0)
pm = formats.pack_forest(forest, "xgb", base=0.5)
model = native.Model(pm.kind, pm.blob, 0)
stream = native.Stream(model, 4096, 0, 2)
timer = native.Timer(stream)
rng = np.random.default_rng(0)
def one_forest(n_trees):
    f = orc.synth_xgb_forest(n_trees, 6, 32, seed=0)
    pmm = formats.pack_forest(f, "xgb", base=0.5)
    mm = native.Model(pmm.kind, pmm.blob, 0)
    ss = native.Stream(mm, 64, 0, 2)
    X = rng.standard_normal((64, 32)).astype(np.float32)
    di = native.DeviceBuffer(X.nbytes); di.upload(X)
    do = native.DeviceBuffer(64 * 4)
    tt = native.Timer(ss)
    for _ in range(5):
        ss.infer_device(64, [di.ptr], [do.ptr])
    tt.start()
    for _ in range(50):
        ss.infer_device(64, [di.ptr], [do.ptr])
    tt.stop()
    w = tt.elapsed_ms() / 50
    b = (ctypes.c_longlong * 64)()
    native.check(native.lib().b2s_debug_read(mm.handle, b))
    s0 = np.array(b[:]).reshape(8, 8)[0]
    print("trees=%d rows=64 warm50=%.2fus  cta0: load=%d traverse+publish=%d first_wait=%d sum=%d total=%d" % (
        n_trees, w * 1e3, s0[1] - s0[0], s0[2] - s0[1], s0[3] - s0[2], s0[4] - s0[3], s0[4] - s0[0]))
    di.free(); do.free(); ss.destroy(); mm.free()


for nt in (16, 128, 256, 512):
    one_forest(nt)

for rows in (1, 64, 1024, 4096):
    X = rng.standard_normal((rows, 32)).astype(np.float32)
    d_in = native.DeviceBuffer(X.nbytes); d_in.upload(X)
    d_out = native.DeviceBuffer(rows * 4)
    for _ in range(5):
        stream.infer_device(rows, [d_in.ptr], [d_out.ptr])
    stream.synchronize()
    stream.flush_l2(); timer.start(); stream.infer_device(rows, [d_in.ptr], [d_out.ptr]); timer.stop(); cold = timer.elapsed_ms()
    for _ in range(3):
        stream.infer_device(rows, [d_in.ptr], [d_out.ptr])
    # warm, single launch behind a busy GPU (flush of L2 would evict; use a dummy launch as queue primer)
    stream.infer_device(rows, [d_in.ptr], [d_out.ptr])
    timer.start(); stream.infer_device(rows, [d_in.ptr], [d_out.ptr]); timer.stop(); warm1 = timer.elapsed_ms()
    timer.start()
    for _ in range(50):
        stream.infer_device(rows, [d_in.ptr], [d_out.ptr])
    timer.stop(); warm50 = timer.elapsed_ms() / 50
    buf = (ctypes.c_longlong * 64)()
    native.check(native.lib().b2s_debug_read(model.handle, buf))
    st = np.array(buf[:]).reshape(8, 8)
    print("rows=%d cold=%.2fus warm1=%.2fus warm50=%.2fus" % (rows, cold * 1e3, warm1 * 1e3, warm50 * 1e3))
    for r in range(min(8, st.shape[0])):
        s = st[r]
        if s[0] == 0:
            continue
        print("   cta%d: load=%d traverse+publish=%d first_wait=%d sum=%d (rank1 landed +%d, last rank landed +%d after first_wait) total=%d cycles" % (
            r, s[1] - s[0], s[2] - s[1], s[3] - s[2], (s[4] - s[3]) if s[4] else -1, (s[5] - s[3]) if s[5] else -1,
            (s[6] - s[3]) if s[6] else -1, (s[4] if s[4] else s[2]) - s[0]))
    d_in.free(); d_out.free()

# ---- host-side split of the serial e2e latency: submit call vs wait call (64 one-row requests)
import time
stream2 = native.Stream(model, 64, 0, 4)
X = rng.standard_normal((64, 32)).astype(np.float32)
tin = (native.Tensor * 64)(); tout = (native.Tensor * 64)()
ob = np.zeros((64, 1), np.float32)
for r in range(64):
    row = X[r:r + 1]
    tin[r].data = row.ctypes.data; tin[r].dtype = 0; tin[r].ndim = 2; tin[r].shape[0], tin[r].shape[1] = 1, 32
    tout[r].data = ob[r].ctypes.data
lib = native.lib()
ev = ctypes.c_uint64(0)
for _ in range(50):
    native.check(lib.b2s_infer_batch(model.handle, stream2.handle, 64, tin, tout, ctypes.byref(ev)))
    native.check(lib.b2s_event_wait(ev.value))
ts, tw = [], []
for _ in range(500):
    t0 = time.perf_counter()
    native.check(lib.b2s_infer_batch(model.handle, stream2.handle, 64, tin, tout, ctypes.byref(ev)))
    t1 = time.perf_counter()
    native.check(lib.b2s_event_wait(ev.value))
    t2 = time.perf_counter()
    ts.append(t1 - t0); tw.append(t2 - t1)
ts, tw = np.array(ts) * 1e6, np.array(tw) * 1e6
print("serial e2e (64 requests): submit p50=%.1fus p99=%.1fus | wait p50=%.1fus p99=%.1fus | total p50=%.1fus" % (
    np.percentile(ts, 50), np.percentile(ts, 99), np.percentile(tw, 50), np.percentile(tw, 99), np.percentile(ts + tw, 50)))
