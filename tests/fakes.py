"""Test doubles (CPU suite): a stream that 'executes' batches on the host so the Python scheduler,
engine marshalling and router can be exercised without a GPU.  Never used by the product."""
import threading
import time

import numpy as np


class FakeModel(object):
    def __init__(self, n_features=4, in_dtype=np.float32, out_dtype=np.float32, fn=None):
        self.n_inputs, self.n_outputs = 1, 1
        self.in_dtypes, self.out_dtypes = [np.dtype(in_dtype)], [np.dtype(out_dtype)]
        self.in_row_elems, self.out_row_elems = [n_features], [1]
        self.fn = fn or (lambda x: x.sum(axis=1))

        class _I(object):
            kind = 1
        self.info = _I()

    def free(self):
        pass


class _Slot(object):
    def __init__(self, index, model, max_rows):
        self.index = index
        self.inputs = [np.zeros((max_rows, model.in_row_elems[0]), model.in_dtypes[0])]
        self.outputs = [np.zeros(max_rows, model.out_dtypes[0])]


class FakeStream(object):
    """acquire/submit/wait/release with the semantics of native.Stream; records batch sizes."""

    def __init__(self, model, max_rows, n_slots=2, latency_s=0.0, fail_on=None):
        self.model = model
        self.slots = [_Slot(i, model, max_rows) for i in range(n_slots)]
        self.free = list(range(n_slots))
        self.lock = threading.Lock()
        self.batches = []
        self.latency_s = latency_s
        self.fail_on = fail_on
        self.pending = {}
        self.n = 0
        self.max_row_elems = 0

    def acquire(self):
        from clearml_serving_b200 import native
        with self.lock:
            if not self.free:
                raise native.B2SError(native.B2S_ERR_BUSY, "busy")
            return self.slots[self.free.pop(0)]

    def submit(self, slot, n_rows, row_offsets=None):
        self.n += 1
        self.batches.append(int(n_rows))
        if self.fail_on is not None and (self.n in self.fail_on if isinstance(self.fail_on, (set, list, tuple)) else self.n == self.fail_on):
            raise ValueError("injected submit failure")
        self.pending[self.n] = (slot, n_rows, time.perf_counter())
        return self.n

    def collate_submit(self, slot, requests):
        """host-side stand-in for b2s_slot_collate: rows back to back in request order"""
        n_rows = 0
        for r in requests:
            slot.inputs[0][n_rows:n_rows + r.rows] = r.inputs[0]
            n_rows += r.rows
        return self.submit(slot, n_rows), n_rows

    def wait(self, ev):
        slot, n_rows, t0 = self.pending.pop(ev)
        dt = self.latency_s - (time.perf_counter() - t0)
        if dt > 0:
            time.sleep(dt)
        slot.outputs[0][:n_rows] = self.model.fn(slot.inputs[0][:n_rows])

    def release(self, slot):
        with self.lock:
            self.free.append(slot.index)

    def destroy(self):
        pass


def make_fake_engine(endpoint, model, preprocess=None, policy=None, latency_s=0.0, n_replicas=1):
    """A B200PreprocessRequest whose native model/stream are the host-side fakes above (CPU suite
    only): everything above the C ABI -- marshalling, batching, futures -- is the real code."""
    from clearml_serving_b200.preprocess_service import B200PreprocessRequest
    from clearml_serving_b200.scheduler import BatchPolicy, DynamicBatcher
    policy = policy or BatchPolicy.from_auxiliary_cfg(getattr(endpoint, "auxiliary_cfg", None))
    e = B200PreprocessRequest.__new__(B200PreprocessRequest)
    e.model_endpoint = endpoint
    e._preprocess = preprocess
    e._timeout = 30
    e._native_model = model
    e._model = model
    e._policy = policy
    from clearml_serving_b200.router import Replica, ReplicaSet
    reps = []
    for dev in range(n_replicas):
        reps.append(Replica(dev, model, DynamicBatcher(
            model, policy, name="fake{}".format(dev),
            stream=FakeStream(model, policy.max_batch_size, n_slots=policy.n_slots, latency_s=latency_s))))
    e._replicas = ReplicaSet(reps)
    e._batcher = reps[0].batcher
    return e
