"""Round-robin replica router (north_star: "model replicas shard across the 8 GPUs of
one box with a Python-side round-robin router").

The reference scales by running more containers behind an external load balancer
(/root/reference/README.md:39-47); inside one serving process the equivalent is one engine replica
per GPU -- its own model copy in that GPU's HBM, its own CUDA stream and batcher -- and a counter
that deals requests out in turn.  Requests are independent, so there is no cross-GPU traffic and no
collective on the data path.
"""
import itertools


class Replica(object):
    __slots__ = ("device", "model", "batcher")

    def __init__(self, device, model, batcher):
        self.device, self.model, self.batcher = device, model, batcher


class ReplicaSet(object):
    def __init__(self, replicas):
        if not replicas:
            raise ValueError("ReplicaSet needs at least one replica")
        self.replicas = list(replicas)
        self._turn = itertools.count()   # next() is atomic under the GIL

    def __len__(self):
        return len(self.replicas)

    def pick(self):
        return self.replicas[next(self._turn) % len(self.replicas)]

    def shutdown(self):
        for r in self.replicas:
            try:
                r.batcher.shutdown()
            finally:
                r.model.free()
        self.replicas = []

    def snapshot_stats(self):
        per = [r.batcher.snapshot_stats() for r in self.replicas]
        tot = dict(replicas=len(per), devices=[r.device for r in self.replicas])
        for k in ("batches", "requests", "rows", "failed_requests", "queue_delay_us_sum", "exec_us_sum", "in_bytes", "out_bytes"):
            tot[k] = sum(p.get(k, 0) for p in per)
        tot["batch_rows_hist"] = [sum(col) for col in zip(*[p["batch_rows_hist"] for p in per])] if per else []
        tot["max_batch_rows"] = max(p["max_batch_rows"] for p in per)
        tot["mean_batch_rows"] = tot["rows"] / tot["batches"] if tot["batches"] else 0.0
        tot["mean_exec_us"] = tot["exec_us_sum"] / tot["batches"] if tot["batches"] else 0.0
        tot["mean_queue_delay_us"] = tot["queue_delay_us_sum"] / tot["requests"] if tot["requests"] else 0.0
        tot["per_replica_requests"] = [p["requests"] for p in per]
        return tot


def parse_devices(aux, default_device):
    """`b200.devices` ([0,1,2,3] or "0,1,2,3" or "all") / `b200.device` in the endpoint's auxiliary_cfg."""
    if isinstance(aux, dict):
        v = aux.get("b200.devices", None)
        if v is None and isinstance(aux.get("b200"), dict):
            v = aux["b200"].get("devices")
        if v is not None:
            if isinstance(v, str):
                if v.strip().lower() == "all":
                    from . import native
                    return list(range(max(1, native.device_count())))
                return [int(x) for x in v.replace("[", "").replace("]", "").split(",") if x.strip()]
            if isinstance(v, (list, tuple)):
                return [int(x) for x in v]
            return [int(v)]
        if "b200.device" in aux:
            return [int(aux["b200.device"])]
    return [int(default_device)]
