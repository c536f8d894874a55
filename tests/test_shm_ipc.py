"""Shared-memory worker <-> engine-process path (clearml_serving_b200/shm_ipc.py) on CPU: the engine
process side runs the real ModelRequestProcessor + B200 engine class over the host-side fake native layer; clients live in
this process AND in separate worker processes (as uvicorn workers would)."""
import asyncio
import json
import os
import subprocess
import sys
import uuid

import numpy as np
import pytest

from clearml_serving_b200 import ModelEndpoint, shm_ipc
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from clearml_serving_b200.scheduler import BatchPolicy
from tests.fakes import FakeModel, make_fake_engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import asyncio, json, sys
import numpy as np
sys.path.insert(0, {root!r})
from clearml_serving_b200 import shm_ipc
async def main():
    c = shm_ipc.EngineClient({name!r}, timeout_s=20)
    xs = [np.full((1 + i % 3, 4), {base} + i, np.float32) for i in range({n})]
    outs = await asyncio.gather(*[c.request("sum", [x]) for x in xs])
    bad = None
    try:
        await c.request("nope", [xs[0]])
    except ValueError as ex:
        bad = str(ex)
    c.close()
    print(json.dumps(dict(ok=[bool(np.array_equal(o, x.sum(axis=1))) for o, x in zip(outs, xs)], bad=bad)))
asyncio.run(main())
"""


@pytest.fixture()
def server():
    name = "t" + uuid.uuid4().hex[:8]
    proc = ModelRequestProcessor()
    ep = ModelEndpoint(engine_type="b200", serving_url="sum", auxiliary_cfg={"max_batch_size": 16})
    eng = make_fake_engine(ep, FakeModel(n_features=4), policy=BatchPolicy(max_batch_size=16, max_queue_delay_us=3000, n_slots=2),
                           latency_s=0.002)
    proc._endpoints["sum"] = ep
    proc._engine_processor_lookup["sum"] = eng
    srv = shm_ipc.EngineServer(proc, name=name, n_clients=4, slots_per_client=8, slot_bytes=1 << 16)
    yield srv, name, eng
    srv.close()
    proc.shutdown()


def test_workers_in_other_processes_share_one_engine(server):
    srv, name, eng = server
    workers = [subprocess.Popen([sys.executable, "-c", _WORKER.format(root=ROOT, name=name, base=100 * k, n=20)],
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for k in range(2)]

    async def local():
        c = shm_ipc.EngineClient(name, timeout_s=20)
        try:
            xs = [np.full((2, 4), 7 + i, np.float32) for i in range(30)]           # more requests than this client's 8 slots
            outs = await asyncio.gather(*[c.request("sum", [x]) for x in xs])
            assert all(np.array_equal(o, x.sum(axis=1)) for o, x in zip(outs, xs))
            with pytest.raises(ValueError, match="does not fit"):
                await c.request("sum", [np.zeros((1, 1 << 15), np.float32)])
            with pytest.raises(ValueError, match="features per row|model expects"):
                await c.request("sum", [np.zeros((1, 3), np.float32)])             # the engine's own validation error comes back
        finally:
            c.close()
    asyncio.run(local())
    for w in workers:
        out, err = w.communicate(timeout=60)
        assert w.returncode == 0, err[-2000:]
        res = json.loads(out.strip().splitlines()[-1])
        assert all(res["ok"]) and len(res["ok"]) == 20
        assert "not found" in res["bad"]                                          # EndpointNotFound travels back as the message
    st = eng.engine_stats()
    assert st["requests"] == 70 and st["batches"] < 70                            # one batcher served every worker: cross-process batching
    assert srv.stats["clients"] == 3 and srv.stats["requests"] == 70 and srv.stats["errors"] >= 3


def test_remote_engine_class_behind_the_plugin_surface(server):
    srv, name, eng = server
    from clearml_serving_b200.preprocess_service import BasePreprocessRequest
    cls = BasePreprocessRequest.get_engine_cls("b200_remote")
    assert cls is shm_ipc.RemoteB200PreprocessRequest and cls.is_process_async
    worker = ModelRequestProcessor()                                              # what a uvicorn worker holds: no model, no GPU
    worker.add_endpoint(ModelEndpoint(engine_type="b200_remote", serving_url="sum", auxiliary_cfg={"b200.engine_socket": name}))

    async def run():
        xs = [np.full((1, 4), i, np.float32) for i in range(12)]
        outs = await asyncio.gather(*[worker.process_request("sum", None, x) for x in xs])
        return [float(np.asarray(o).ravel()[0]) for o in outs]
    try:
        assert asyncio.run(run()) == [4.0 * i for i in range(12)]
    finally:
        for c in list(shm_ipc.RemoteB200PreprocessRequest._clients.values()):
            c.close()
        shm_ipc.RemoteB200PreprocessRequest._clients.clear()
        worker.shutdown()
