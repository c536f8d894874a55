"""Per-endpoint request queue and dynamic-batch scheduler (Python, as north_star requires).

The reference has no such component of its own: "auto-batching" is tritonserver's dynamic batcher,
configured by keys passed through `--aux-config` (examples/huggingface/readme.md:113) and merged
verbatim into config.pbtxt (clearml_serving/engines/triton/triton_helper.py:326-331).  This module
implements those semantics in front of libb200serve:

  * `max_batch_size = N`: requests carry their own leading batch dim; a batch holds <= N rows.
  * FIFO per endpoint; a batch is dispatched as soon as a staging slot is free, taking as many
    queued requests as fit.
  * `dynamic_batching.max_queue_delay_microseconds = D`: if the queue cannot fill a batch (or hit a
    `preferred_batch_size`), wait up to D us from the arrival of the oldest request, then dispatch
    what is there.  D = 0 (default) dispatches immediately.
  * responses are per request (the batch output is split on dim 0).

Threads: `dispatch` forms batches and hands the requests' host pointers to the library, which collates
their rows into a pinned staging slot and submits the batch (b2s_slot_collate: one C call per batch, GIL
released); `complete` blocks in b2s_event_wait (GIL released), takes one copy per output out of the slot,
gives every request row views of it and resolves the futures -- one `loop.call_soon_threadsafe` per event
loop per batch.
"""
import asyncio
import bisect
import collections
import concurrent.futures
import re
import threading
import time

import numpy as np

from . import native


class BatchPolicy(object):
    def __init__(self, max_batch_size=64, max_queue_delay_us=0, preferred_batch_size=None, n_slots=4):
        self.max_batch_size = max(1, int(max_batch_size))
        self.max_queue_delay_us = max(0, int(max_queue_delay_us))
        self.preferred_batch_size = sorted(int(x) for x in (preferred_batch_size or []) if int(x) > 0)
        self.n_slots = max(1, int(n_slots))

    def __repr__(self):
        return "BatchPolicy(max_batch_size={}, max_queue_delay_us={}, preferred_batch_size={}, n_slots={})".format(
            self.max_batch_size, self.max_queue_delay_us, self.preferred_batch_size, self.n_slots)

    @classmethod
    def from_auxiliary_cfg(cls, aux, default_max_batch_size=64):
        """`aux` is ModelEndpoint.auxiliary_cfg: a dict (possibly with dotted keys, as the CLI's
        key=value form produces: `dynamic_batching.max_queue_delay_microseconds=5000`), a nested
        dict, or config.pbtxt text.  Unknown keys (platform, default_model_filename ...) are ignored."""
        flat = {}
        if isinstance(aux, dict):
            _flatten(aux, "", flat)
        elif isinstance(aux, str):
            flat = _parse_pbtxt(aux)
        mbs = flat.get("max_batch_size", default_max_batch_size)
        delay = flat.get("dynamic_batching.max_queue_delay_microseconds", 0)
        pref = flat.get("dynamic_batching.preferred_batch_size", None)
        if isinstance(pref, str):
            pref = [int(x) for x in re.findall(r"-?\d+", pref)]
        elif isinstance(pref, (int, float)):
            pref = [int(pref)]
        slots = flat.get("b200.staging_slots", 4)
        mbs = int(mbs)
        if mbs <= 0:  # Triton: max_batch_size 0 == batching disabled -> one request per launch
            mbs = 1
        return cls(max_batch_size=mbs, max_queue_delay_us=int(float(delay)), preferred_batch_size=pref,
                   n_slots=int(slots))


def _flatten(d, prefix, out):
    for k, v in d.items():
        key = "{}{}".format(prefix, k)
        if isinstance(v, dict):
            _flatten(v, key + ".", out)
        else:
            out[key] = v


def _parse_pbtxt(text):
    """Tiny reader for the handful of config.pbtxt keys the scheduler consumes."""
    out = {}
    m = re.search(r"max_batch_size\s*:\s*(\d+)", text)
    if m:
        out["max_batch_size"] = int(m.group(1))
    m = re.search(r"dynamic_batching\s*:?\s*\{([^}]*)\}", text, re.S)
    if m:
        body = m.group(1)
        d = re.search(r"max_queue_delay_microseconds\s*:\s*(\d+)", body)
        if d:
            out["dynamic_batching.max_queue_delay_microseconds"] = int(d.group(1))
        p = re.search(r"preferred_batch_size\s*:\s*\[([^\]]*)\]", body)
        if p:
            out["dynamic_batching.preferred_batch_size"] = [int(x) for x in re.findall(r"\d+", p.group(1))]
        elif re.search(r"preferred_batch_size\s*:\s*(\d+)", body):
            out["dynamic_batching.preferred_batch_size"] = [int(x) for x in re.findall(r"preferred_batch_size\s*:\s*(\d+)", body)]
    return out


class _Request(object):
    __slots__ = ("inputs", "rows", "row_len", "ptrs", "t_enq", "loop", "afuture", "cfuture", "info")

    def __init__(self, inputs, rows):
        self.inputs = inputs          # C-contiguous arrays [rows, row_elems], kept alive until the batch is collated
        self.rows = rows
        self.row_len = inputs[0].shape[1] if inputs and inputs[0].ndim > 1 else 0
        self.ptrs = [a.__array_interface__["data"][0] for a in inputs]
        self.t_enq = time.perf_counter()
        self.loop = None
        self.afuture = None
        self.cfuture = None
        self.info = None   # optional dict the caller wants filled with this request's batch figures (sampled statistics)


def _resolve_many(pairs):
    for fut, result, exc in pairs:
        if fut.done():  # cancelled / timed out by the caller
            continue
        if exc is not None:
            fut.set_exception(exc)
        else:
            fut.set_result(result)


BATCH_ROWS_BUCKETS = (1, 2, 4, 8, 16, 32, 64, 128, 256)   # histogram upper bounds (le), + one overflow bucket


class DynamicBatcher(object):
    """Queue + scheduler of one endpoint in front of one native.Stream."""

    def __init__(self, model, policy, name="endpoint", stream=None, max_row_elems=0, request_timeout_s=None):
        self.model = model
        self.policy = policy
        self.name = name
        # a request that waited longer than this in the queue is failed instead of dispatched (the engine's
        # per-request deadline, preprocess_service.py:48-49, without a timer object per request)
        self.request_timeout_s = request_timeout_s
        self.max_row_elems = int(max_row_elems)
        self.ragged = any(e < 0 for e in model.in_row_elems)
        # `stream` is injectable so the queueing logic can be unit-tested without a GPU
        self.stream = stream if stream is not None else native.Stream(model, policy.max_batch_size, self.max_row_elems,
                                                                      policy.n_slots)
        self._queue = collections.deque()
        self._queued_rows = 0
        self._cond = threading.Condition()
        self._inflight = collections.deque()
        self._inflight_cond = threading.Condition()
        self._running = True
        self._fatal = None
        # counters (what tritonserver's :8002/metrics used to expose for the model)
        self.stats = dict(batches=0, requests=0, rows=0, failed_requests=0, queue_delay_us_sum=0.0, exec_us_sum=0.0,
                          in_bytes=0, out_bytes=0, max_batch_rows=0,
                          batch_rows_hist=[0] * (len(BATCH_ROWS_BUCKETS) + 1))
        self._dispatch_thread = threading.Thread(target=self._dispatch_loop, name="b2s-dispatch-" + name, daemon=True)
        self._complete_thread = threading.Thread(target=self._complete_loop, name="b2s-complete-" + name, daemon=True)
        self._dispatch_thread.start()
        self._complete_thread.start()

    # ------------------------------------------------------------------ submission
    def _enqueue(self, req):
        if not self._running:
            raise ValueError("b200 engine: endpoint '{}' is shut down".format(self.name))
        if req.rows > self.policy.max_batch_size:
            raise ValueError("b200 engine: request with {} rows exceeds max_batch_size {}".format(
                req.rows, self.policy.max_batch_size))
        with self._cond:
            self._queue.append(req)
            self._queued_rows += req.rows
            self._cond.notify()

    def submit_async(self, inputs, rows, info=None):
        """Called on an asyncio loop; returns an awaitable future of the list of output arrays.  `info`: a dict that
        receives this request's batch_rows / queue_us / exec_us before the future resolves."""
        req = _Request(inputs, rows)
        req.info = info
        req.loop = asyncio.get_running_loop()
        req.afuture = req.loop.create_future()
        self._enqueue(req)
        return req.afuture

    def submit(self, inputs, rows, info=None):
        """Thread-safe synchronous-style submission; returns a concurrent.futures.Future."""
        req = _Request(inputs, rows)
        req.info = info
        req.cfuture = concurrent.futures.Future()
        self._enqueue(req)
        return req.cfuture

    # ------------------------------------------------------------------ batch formation
    def _take_batch(self):
        """Blocks until a batch should be dispatched (SURVEY.md 5.9 rules); returns list of requests."""
        pol = self.policy
        with self._cond:
            while self._running and not self._queue:
                self._cond.wait()
            if not self._running and not self._queue:
                return None
            if pol.max_queue_delay_us > 0:
                deadline = self._queue[0].t_enq + pol.max_queue_delay_us * 1e-6
                while self._running:
                    rows = self._queued_rows
                    if rows >= pol.max_batch_size:
                        break
                    if pol.preferred_batch_size and rows in pol.preferred_batch_size and \
                            rows == pol.preferred_batch_size[-1]:
                        break  # largest preferred size reached exactly: go now
                    now = time.perf_counter()
                    if now >= deadline:
                        break
                    self._cond.wait(timeout=deadline - now)
            batch, rows = [], 0
            limit = pol.max_batch_size
            if pol.preferred_batch_size and self._queued_rows < pol.max_batch_size:
                # dispatch the largest preferred size that the queue can form, else everything
                fits = [p for p in pol.preferred_batch_size if p <= self._queued_rows]
                if fits:
                    limit = fits[-1]
            expired = []
            while self._queue and rows + self._queue[0].rows <= limit:
                r = self._queue.popleft()
                if self.request_timeout_s is not None and time.perf_counter() - r.t_enq > self.request_timeout_s:
                    self._queued_rows -= r.rows
                    expired.append(r)
                    continue
                rows += r.rows
                batch.append(r)
            if expired:
                self._fail(expired, ValueError("b200 engine: request timed out after {}s in the queue".format(
                    self.request_timeout_s)))
            while not batch and self._queue:  # head does not fit a preferred size: take it alone
                r = self._queue.popleft()
                if self.request_timeout_s is not None and time.perf_counter() - r.t_enq > self.request_timeout_s:
                    self._queued_rows -= r.rows
                    self._fail([r], ValueError("b200 engine: request timed out after {}s in the queue".format(
                        self.request_timeout_s)))
                    continue
                rows += r.rows
                batch.append(r)
            self._queued_rows -= rows
            return batch

    def _dispatch_loop(self):
        while True:
            batch = self._take_batch()
            if batch is None:
                break
            if not batch:      # every popped request had expired in the queue
                continue
            self.stats["requests"] += len(batch)   # every request taken off the queue; failures are counted on top
            slot = None
            try:
                slot = self._acquire_slot()
                t_disp = time.perf_counter()
                if self.request_timeout_s is not None:   # the wait for a free slot counts against the deadline too
                    late = [r for r in batch if t_disp - r.t_enq > self.request_timeout_s]
                    if late:
                        batch = [r for r in batch if t_disp - r.t_enq <= self.request_timeout_s]
                        self._fail(late, ValueError("b200 engine: request timed out after {}s in the queue".format(
                            self.request_timeout_s)))
                        if not batch:
                            self.stream.release(slot)
                            continue
                # the collate (rows back to back in request order, cu_seqlens for variable-length models) happens inside
                # the library, without the GIL
                ev, n_rows = self.stream.collate_submit(slot, batch)
                st = self.stats
                st["batches"] += 1
                st["rows"] += n_rows
                st["queue_delay_us_sum"] += sum((t_disp - r.t_enq) for r in batch) * 1e6
                st["in_bytes"] += sum(a.nbytes for r in batch for a in r.inputs)
                if n_rows > st["max_batch_rows"]:
                    st["max_batch_rows"] = n_rows
                st["batch_rows_hist"][bisect.bisect_left(BATCH_ROWS_BUCKETS, n_rows)] += 1
                for r in batch:
                    r.inputs = None    # the rows now live in the slot
                    if r.info is not None:
                        r.info["batch_rows"] = n_rows
                        r.info["queue_us"] = (t_disp - r.t_enq) * 1e6
                        r.info["t_dispatch"] = t_disp
                with self._inflight_cond:
                    self._inflight.append((ev, slot, batch, t_disp))
                    self._inflight_cond.notify()
            except Exception as ex:  # a failed batch fails only its own requests -- and gives its slot back
                if slot is not None:
                    try:
                        self.stream.release(slot)
                    except Exception:  # noqa
                        pass
                    with self._inflight_cond:
                        self._inflight_cond.notify_all()
                self._fail(batch, ex)

    def _acquire_slot(self):
        while True:
            try:
                return self.stream.acquire()
            except native.B2SError as ex:
                if ex.code != native.B2S_ERR_BUSY:
                    raise
            with self._inflight_cond:  # all slots in flight: wait for the completion thread
                self._inflight_cond.wait(timeout=0.001)

    # ------------------------------------------------------------------ completion
    def _complete_loop(self):
        m = self.model
        while True:
            with self._inflight_cond:
                while self._running and not self._inflight:
                    self._inflight_cond.wait()
                if not self._inflight:
                    if not self._running:
                        break
                    continue
                ev, slot, batch, t_disp = self._inflight.popleft()
            try:
                self.stream.wait(ev)
                exec_us = (time.perf_counter() - t_disp) * 1e6   # collate hand-off -> results in the pinned slot
                self.stats["exec_us_sum"] += exec_us
                for r in batch:
                    if r.info is not None:
                        r.info["exec_us"] = exec_us
                # one copy per output per batch out of the pinned slot; requests get row views of it
                n_rows = sum(r.rows for r in batch)
                outs = [slot.outputs[o][:n_rows].copy() for o in range(m.n_outputs)]
                row = 0
                results = []
                for r in batch:
                    results.append([out[row:row + r.rows] for out in outs])
                    row += r.rows
                self.stats["out_bytes"] += sum(o.nbytes for o in outs)
                self.stream.release(slot)
                with self._inflight_cond:
                    self._inflight_cond.notify_all()
                self._resolve(batch, results, None)
            except Exception as ex:  # noqa
                try:
                    self.stream.release(slot)
                except Exception:  # noqa
                    pass
                self._fail(batch, ex)

    def _resolve(self, batch, results, exc):
        by_loop = {}
        for k, r in enumerate(batch):
            res = results[k] if results is not None else None
            if r.cfuture is not None:
                if exc is not None:
                    r.cfuture.set_exception(exc)
                else:
                    r.cfuture.set_result(res)
            else:
                by_loop.setdefault(r.loop, []).append((r.afuture, res, exc))
        for loop, pairs in by_loop.items():
            try:
                loop.call_soon_threadsafe(_resolve_many, pairs)
            except RuntimeError:  # loop closed
                pass

    def _fail(self, batch, ex):
        if not isinstance(ex, Exception):
            ex = ValueError(str(ex))
        self.stats["failed_requests"] += len(batch)
        self._resolve(batch, None, ex)

    # ------------------------------------------------------------------ lifetime
    def shutdown(self):
        if not self._running:
            return
        self._running = False
        with self._cond:
            self._cond.notify_all()
        self._dispatch_thread.join(timeout=5)
        with self._inflight_cond:
            self._inflight_cond.notify_all()
        self._complete_thread.join(timeout=5)
        pending = list(self._queue)
        self._queue.clear()
        if pending:
            self._fail(pending, ValueError("b200 engine: endpoint '{}' shut down".format(self.name)))
        try:
            self.stream.destroy()
        except Exception:  # noqa
            pass

    def snapshot_stats(self):
        st = dict(self.stats)
        st["batch_rows_hist"] = list(st["batch_rows_hist"])
        st["mean_exec_us"] = st["exec_us_sum"] / st["batches"] if st["batches"] else 0.0
        st["mean_batch_rows"] = st["rows"] / st["batches"] if st["batches"] else 0.0
        st["mean_queue_delay_us"] = st["queue_delay_us_sum"] / st["requests"] if st["requests"] else 0.0
        return st
