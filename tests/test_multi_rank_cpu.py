"""The N>1 path of bench.py on CPU: two processes, gloo backend (the GPU box runs the same helpers over
NCCL).  Requests are independent, so ranks are replicas: no data-path collective, only the timing
barrier, the max-over-ranks reduction and the whole-job aggregation are shared."""
import os
import subprocess
import sys
import textwrap

from tests.conftest import ROOT


def test_two_rank_gloo_timing_and_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent('''
        import json, os, sys, time
        sys.path.insert(0, %r)
        import bench
        rank, world, local, dist = bench._dist_setup(2)
        assert world == 2 and dist is not None and dist.get_backend() == "gloo"
        bench._barrier_sync(dist, local)
        # each rank is a replica serving its own 64-request batches; rank 1 is slower
        my_seconds = 0.010 * (1 + rank)
        mx = bench._max_over_ranks(dist, local, my_seconds)
        total = bench._sum_over_ranks(dist, local, 100.0 * (rank + 1))
        bench._barrier_sync(dist, local)
        out = dict(rank=rank, mx=mx, total=total, value=bench.whole_job_value(world, 64, 10, mx))
        print("RESULT " + json.dumps(out), flush=True)
        dist.destroy_process_group()
    ''' % ROOT))
    env = dict(os.environ, B2S_DIST_BACKEND="gloo", MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    import json
    res = []
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        line = [l for l in o.splitlines() if l.startswith("RESULT ")][0]
        res.append(json.loads(line[7:]))
    for r in res:
        assert abs(r["mx"] - 0.020) < 1e-12          # the slowest rank defines the step time
        assert r["total"] == 300.0
        assert abs(r["value"] - 2 * 64 * 10 / 0.020) < 1e-6


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29732")
    outs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "20", "--warmup", "3"], env=e, capture_output=True, text=True, timeout=240)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout.strip())
    import json
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and outs[1] == ""


def test_tensor_parallel_leader_and_follower_replay_the_same_calls(tmp_path):
    """LLM endpoint, tensor_parallel_size 2 (host logic on CPU, gloo): rank 0 (the serving process) announces every
    engine call, rank 1 replays it with identical arguments, so both ranks would issue the same kernel sequence;
    the pair's gloo sub-group is the one bench.py builds for ranks (2i, 2i+1)."""
    script = tmp_path / "tp_worker.py"
    script.write_text(textwrap.dedent('''
        import json, os, sys
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from clearml_serving_b200 import llm_service as S

        class Rec(object):
            def __init__(self):
                self.calls, self.closed = [], False
            def generate(self, prompts, n, on_progress=None, chunk=8):
                # a streamed wave must be chunked alike on both ranks (they synchronise at the same decode steps)
                self.calls.append(([np.asarray(p).tolist() for p in prompts], int(n), int(chunk) if on_progress else 0))
                return np.zeros((len(prompts), n), np.int32)
            def close(self):
                self.closed = True

        dist.init_process_group("gloo")
        rank = dist.get_rank()
        group = dist.new_group(ranks=[0, 1], backend="gloo")
        eng = Rec()
        if rank == 0:
            lead = S.TensorParallelLeader(eng, group)
            lead.generate([[1, 2, 3], [4]], 5)
            lead.generate([[9] * 7], 2, on_progress=lambda w0, toks: None, chunk=4)
            lead.close()
        else:
            S.follower_loop(eng, group)
        print("RESULT " + json.dumps(dict(rank=rank, calls=eng.calls, closed=eng.closed)), flush=True)
        dist.destroy_process_group()
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29733", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    import json
    res = []
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        res.append(json.loads([l for l in o.splitlines() if l.startswith("RESULT ")][0][7:]))
    assert res[0]["calls"] == res[1]["calls"] == [[[[1, 2, 3], [4]], 5, 0], [[[9] * 7], 2, 4]]
    assert res[0]["closed"] and res[1]["closed"]


def test_continuous_batching_iterations_are_replayed_by_the_follower(tmp_path):
    """tensor_parallel_size 2 under the continuous scheduler: every scheduler iteration (page-table updates, prefill into
    slots, decode rows) is announced once and replayed by rank 1 with identical arguments -- both ranks end with the same
    emulated paged cache and the same tokens"""
    script = tmp_path / "tp_cb_worker.py"
    script.write_text(textwrap.dedent('''
        import json, os, sys
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from clearml_serving_b200 import llm_service as S
        from tests.test_llm_service import FakePagedLlm, _expected

        dist.init_process_group("gloo")
        rank = dist.get_rank()
        group = dist.new_group(ranks=[0, 1], backend="gloo")
        eng = FakePagedLlm(max_batch=4, max_ctx=256, n_pages=9)
        ok = True
        if rank == 0:
            lead = S.TensorParallelLeader(eng, group)
            b = S.ContinuousBatcher(lead, max_batch=4, max_ctx=256, chunk=3)
            rng = np.random.default_rng(0)
            reqs = [(rng.integers(0, 1000, int(rng.integers(1, 120))), int(rng.integers(1, 30))) for _ in range(10)]
            futs = [b.submit(p, n) for p, n in reqs]
            ok = all(f.result(timeout=60).tolist() == _expected(p, n) for (p, n), f in zip(reqs, futs))
            b.close()
            lead.close()
        else:
            S.follower_loop(eng, group)
        print("RESULT " + json.dumps(dict(rank=rank, ok=ok, calls=eng.calls, pool=int(eng.pool.sum()), closed=eng.closed)), flush=True)
        dist.destroy_process_group()
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29734", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    import json
    res = []
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        res.append(json.loads([l for l in o.splitlines() if l.startswith("RESULT ")][0][7:]))
    assert res[0]["ok"] and res[0]["calls"] == res[1]["calls"] and len(res[0]["calls"]) > 3
    assert res[0]["pool"] == res[1]["pool"] and res[0]["closed"] and res[1]["closed"]
