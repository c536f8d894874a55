"""Shared-memory path between the REST worker processes and ONE engine process per GPU.

The reference starts N uvicorn / gunicorn workers (clearml_serving/serving/entrypoint.sh:47-73,
CLEARML_SERVING_NUM_PROCESS) and every worker builds its own engine objects -- for the in-process engines that is N
copies of the model and N independent request streams (no cross-worker batching); for Triton the workers share the
sidecar over gRPC (preprocess_service.py:348-422: one protobuf message per request over TCP).  Here the workers keep
what must stay per process -- HTTP parsing, user `Preprocess.preprocess / postprocess` code -- and hand the request
TENSORS to the engine process of the GPU through a shared-memory ring:

    worker (any number)                                       engine process (one per GPU)
      EngineClient.request(url, tensors)                        EngineServer(processor)
        slot <- own free list                                     doorbell (unix datagram, abstract namespace)
        tensor frame written IN PLACE into the slot               -> wire.decode_tensors: zero-copy views of the slot
        doorbell ---------------------------------------------->  -> processor.process_request(url, tensors): the
        await reply doorbell (loop.add_reader)                        b200 engine batches ACROSS workers
        reply frame decoded from the same slot, copied out   <--  reply frame written into the slot, doorbell back
        slot -> free list

One `multiprocessing.shared_memory` segment holds `n_clients x slots_per_client` slots of `slot_bytes`; a client owns
a contiguous range (handed out at HELLO), so no cross-process allocator or lock exists: the only cross-process traffic
besides the payload is an 8-byte datagram each way.  The frames are the REST edge's own (`wire.py`).

`RemoteB200PreprocessRequest` (engine_type "b200_remote") is the worker-side engine class: same plugin surface, `process`
forwards to the engine process named by `auxiliary_cfg["b200.engine_socket"]`.
"""
import asyncio
import os
import socket
import struct
import threading
import time
from multiprocessing import shared_memory

import numpy as np

from . import wire
from .preprocess_service import BasePreprocessRequest

_HDR = struct.Struct("<IIII")       # state / status, url length, frame length, reserved
ST_FREE, ST_REQUEST, ST_REPLY_OK, ST_REPLY_ERR = 0, 1, 2, 3
_MSG = struct.Struct("<II")         # kind, slot
MSG_HELLO, MSG_RANGE, MSG_REQ, MSG_REPLY, MSG_BYE = 1, 2, 3, 4, 5


def _addr(name, who):
    return "\0b2s-{}-{}".format(name, who)      # abstract unix socket: no file, gone with the process


class EngineServer(object):
    """Owns the segment; serves `processor` (a ModelRequestProcessor with b200 endpoints) to any number of clients."""

    def __init__(self, processor, name="gpu0", n_clients=16, slots_per_client=64, slot_bytes=1 << 20, serve_type="process"):
        self.processor, self.name, self.serve_type = processor, name, serve_type
        self.n_clients, self.spc, self.slot_bytes = int(n_clients), int(slots_per_client), int(slot_bytes)
        self.shm = shared_memory.SharedMemory(name="b2s-" + name, create=True, size=self.n_clients * self.spc * self.slot_bytes)
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
        self.sock.bind(_addr(name, "engine"))
        self.sock.settimeout(0.2)
        self._next_client = 0
        self._running = True
        self.stats = dict(requests=0, errors=0, clients=0)
        self._loop = asyncio.new_event_loop()
        self._loop_thread = threading.Thread(target=self._loop.run_forever, name="b2s-shm-loop", daemon=True)
        self._loop_thread.start()
        self._thread = threading.Thread(target=self._serve, name="b2s-shm-" + name, daemon=True)
        self._thread.start()

    def _slot(self, i):
        return memoryview(self.shm.buf)[i * self.slot_bytes:(i + 1) * self.slot_bytes]

    def _serve(self):
        while self._running:
            try:
                data, client = self.sock.recvfrom(64)
            except socket.timeout:
                continue
            except OSError:
                return
            kind, slot = _MSG.unpack_from(data)
            if kind == MSG_HELLO:
                if self._next_client >= self.n_clients:
                    self.sock.sendto(_MSG.pack(MSG_RANGE, 0xFFFFFFFF), client)
                    continue
                first = self._next_client * self.spc
                self._next_client += 1
                self.stats["clients"] += 1
                self.sock.sendto(_MSG.pack(MSG_RANGE, first) + struct.pack("<III", self.spc, self.slot_bytes, 0), client)
            elif kind == MSG_REQ:
                asyncio.run_coroutine_threadsafe(self._handle(slot, client), self._loop)

    async def _handle(self, slot, client):
        mv = self._slot(slot)
        status, frame = ST_REPLY_ERR, b""
        try:
            _st, url_len, frame_len, _r = _HDR.unpack_from(mv, 0)
            url = bytes(mv[16:16 + url_len]).decode()
            off = 16 + url_len + (-url_len % 8)
            tensors = wire.decode_tensors(mv[off:off + frame_len])                # views of the slot: no copy
            body = tensors[0] if len(tensors) == 1 else tensors
            base, _, version = url.partition("//")
            out = await self.processor.process_request(base_url=base, version=version or None, request_body=body,
                                                       serve_type=self.serve_type)
            outs = out if isinstance(out, (list, tuple)) else [out]
            frame = wire.encode_tensors([np.asarray(o) for o in outs])
            status = ST_REPLY_OK
            self.stats["requests"] += 1
        except Exception as ex:  # noqa -- the worker re-raises it as the reference's engines would
            frame = "{}: {}".format(type(ex).__name__, ex).encode()[:4096]
            self.stats["errors"] += 1
        if 16 + len(frame) > self.slot_bytes:
            status, frame = ST_REPLY_ERR, b"ValueError: reply does not fit the shared-memory slot"
        mv[16:16 + len(frame)] = frame
        _HDR.pack_into(mv, 0, status, 0, len(frame), 0)
        try:
            self.sock.sendto(_MSG.pack(MSG_REPLY, slot), client)
        except OSError:
            pass

    def close(self):
        self._running = False
        self._thread.join(timeout=2)
        self._loop.call_soon_threadsafe(self._loop.stop)
        self._loop_thread.join(timeout=2)
        try:
            self.sock.close()
        finally:
            for fn in (self.shm.close, self.shm.unlink):
                try:
                    fn()
                except Exception:  # noqa -- a request still holding a view of the segment: the OS reclaims it at exit
                    pass


class EngineClient(object):
    """Worker-side end: attach, get a slot range, `await request(url, tensors)` from any asyncio loop of this process."""

    def __init__(self, name="gpu0", timeout_s=30.0):
        self.name, self.timeout_s = name, float(timeout_s)
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
        self.sock.bind(_addr(name, "c{}-{}".format(os.getpid(), id(self) & 0xFFFF)))
        self.sock.connect(_addr(name, "engine"))
        self.sock.settimeout(5.0)
        self.sock.send(_MSG.pack(MSG_HELLO, 0))
        data = self.sock.recv(64)
        kind, first = _MSG.unpack_from(data)
        if kind != MSG_RANGE or first == 0xFFFFFFFF:
            raise RuntimeError("b200 engine process '{}' has no free client range".format(name))
        self.first = first
        self.n_slots, self.slot_bytes, _ = struct.unpack_from("<III", data, _MSG.size)
        self.shm = shared_memory.SharedMemory(name="b2s-" + name)
        try:   # Python < 3.13 registers ATTACHED segments with this process's resource tracker, which would unlink the
            from multiprocessing import resource_tracker   # engine's segment when this worker exits
            resource_tracker.unregister(self.shm._name, "shared_memory")
        except Exception:  # noqa
            pass
        self._free = list(range(self.first, self.first + self.n_slots))
        self._waiters = {}
        self._lock = threading.Lock()
        self._readers = set()
        self.sock.setblocking(False)

    def _slot(self, i):
        return memoryview(self.shm.buf)[i * self.slot_bytes:(i + 1) * self.slot_bytes]

    def _on_readable(self):
        while True:
            try:
                data = self.sock.recv(64)
            except (BlockingIOError, InterruptedError):
                return
            except OSError:
                return
            kind, slot = _MSG.unpack_from(data)
            if kind == MSG_REPLY:
                with self._lock:
                    w = self._waiters.pop(slot, None)
                if w is not None:
                    loop, fut = w
                    loop.call_soon_threadsafe(lambda f=fut: f.done() or f.set_result(True))

    def _ensure_reader(self, loop):
        if loop not in self._readers:
            loop.add_reader(self.sock.fileno(), self._on_readable)
            self._readers.add(loop)

    async def request(self, url, tensors, version=None):
        loop = asyncio.get_running_loop()
        self._ensure_reader(loop)
        t0 = time.perf_counter()
        while True:
            with self._lock:
                slot = self._free.pop() if self._free else None
            if slot is not None:
                break
            if time.perf_counter() - t0 > self.timeout_s:
                raise ValueError("b200 engine: no shared-memory slot became free within {}s".format(self.timeout_s))
            await asyncio.sleep(0.0005)
        try:
            mv = self._slot(slot)
            key = (url if not version else "{}//{}".format(url, version)).encode()
            frame = wire.encode_tensors(tensors)
            off = 16 + len(key) + (-len(key) % 8)
            if off + len(frame) > self.slot_bytes:
                raise ValueError("b200 engine: request of {} bytes does not fit the shared-memory slot of {}".format(
                    len(frame), self.slot_bytes))
            mv[16:16 + len(key)] = key
            mv[off:off + len(frame)] = frame
            _HDR.pack_into(mv, 0, ST_REQUEST, len(key), len(frame), 0)
            fut = loop.create_future()
            with self._lock:
                self._waiters[slot] = (loop, fut)
            self.sock.send(_MSG.pack(MSG_REQ, slot))
            await asyncio.wait_for(fut, self.timeout_s)
            status, _u, n, _r = _HDR.unpack_from(mv, 0)
            if status == ST_REPLY_OK:
                outs = [a.copy() for a in wire.decode_tensors(mv[16:16 + n])]      # the slot is recycled: own the memory
                return outs[0] if len(outs) == 1 else outs
            msg = bytes(mv[16:16 + n]).decode("utf-8", "replace")
            raise ValueError(msg)
        finally:
            with self._lock:
                self._waiters.pop(slot, None)
                self._free.append(slot)

    def close(self):
        for loop in list(self._readers):
            try:
                loop.remove_reader(self.sock.fileno())
            except Exception:  # noqa
                pass
        try:
            self.sock.close()
        finally:
            try:
                self.shm.close()
            except Exception:  # noqa
                pass


@BasePreprocessRequest.register_engine("b200_remote", modules=["numpy"])
class RemoteB200PreprocessRequest(BasePreprocessRequest):
    """Worker-side engine: the model lives in the engine process of `auxiliary_cfg["b200.engine_socket"]` (default
    "gpu0"); `process` ships the request tensors through shared memory.  Same async flags as the Triton engine
    (preprocess_service.py:289-291)."""
    is_preprocess_async = False
    is_process_async = True
    is_postprocess_async = False
    _clients = {}
    _clients_lock = threading.Lock()

    def __init__(self, model_endpoint, task=None):
        super(RemoteB200PreprocessRequest, self).__init__(model_endpoint=model_endpoint, task=task)
        aux = getattr(model_endpoint, "auxiliary_cfg", None)
        self._engine_name = (aux or {}).get("b200.engine_socket", "gpu0") if isinstance(aux, dict) else "gpu0"

    def _client(self):
        with self._clients_lock:
            c = self._clients.get(self._engine_name)
            if c is None:
                c = self._clients[self._engine_name] = EngineClient(self._engine_name, timeout_s=self._timeout)
            return c

    async def process(self, data, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "process"):
            return await self._preprocess.process(data, state, collect_custom_statistics_fn)
        tensors = list(data) if isinstance(data, (list, tuple)) and data and isinstance(data[0], np.ndarray) else [np.asarray(data)]
        ep = self.model_endpoint
        return await self._client().request(str(ep.serving_url), tensors, getattr(ep, "version", None) or None)


def main():
    """Engine process of one GPU:  python -m clearml_serving_b200.shm_ipc --endpoints endpoints.json [--name gpu0] [--device 0]
    The REST workers run `uvicorn clearml_serving_b200.main:app --workers N` with the same endpoints declared as
    engine_type "b200_remote" + `"b200.engine_socket": "<name>"` in their auxiliary_cfg."""
    import argparse
    import signal
    from .model_request_processor import ModelRequestProcessor
    from .preprocess_service import B200EngineMixin
    ap = argparse.ArgumentParser()
    ap.add_argument("--endpoints", required=True)
    ap.add_argument("--name", default="gpu0")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--clients", type=int, default=16)
    ap.add_argument("--slots-per-client", type=int, default=64)
    ap.add_argument("--slot-bytes", type=int, default=1 << 20)
    args = ap.parse_args()
    B200EngineMixin._default_device = args.device
    proc = ModelRequestProcessor()
    proc.load_endpoints_file(args.endpoints)
    for url, ep in list(proc._endpoints.items()):      # build the engines now: the first request should not pay the model load
        proc._get_engine(url, ep)
    srv = EngineServer(proc, name=args.name, n_clients=args.clients, slots_per_client=args.slots_per_client, slot_bytes=args.slot_bytes)
    stop = threading.Event()
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *_a: stop.set())
    print("b200 engine process '{}' on cuda:{} serving {}".format(args.name, args.device, sorted(proc._endpoints)), flush=True)
    stop.wait()
    srv.close()
    proc.shutdown()


if __name__ == "__main__":
    main()
