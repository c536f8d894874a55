"""Turn the ncu captures in gpurun_out/ into the compact, committed summaries under profiles/.
    python scripts/summarise_profiles.py r01
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches():
    p = os.path.join(SRC, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(list)
    for r in rows:
        try:
            agg[r["Kernel Name"].split("(")[0][:70]].append(float(r["Metric Value"].replace(",", "")))
        except Exception:
            pass
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, "%s_launch_list_summary.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 20 --warmup 3 --no-plugin --no-llama\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("%-72s %7s %12s %12s %7s\n" % ("kernel", "count", "avg_ns", "total_ns", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-72s %7d %12.0f %12.0f %6.1f%%\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))
    print("launch list:", len(rows), "launches,", len(agg), "kernels")


def full(name, cmd="python bench.py --steps 6 --warmup 3 --no-plugin --no-llama", out_name=None):
    rep = os.path.join(SRC, "ncu_%s.ncu-rep" % name)
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    out = {}
    for i, h in enumerate(hdr):
        if h in KEYS or h == "Kernel Name":
            out[h] = dict(unit=units[i], values=[r[i] for r in rows[2:]])
    with open(os.path.join(OUT, "%s_ncu_%s.txt" % (tag, out_name or name)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:%s (%s)\n" % (name, cmd))
        for k, v in out.items():
            f.write("%-80s %-14s %s\n" % (k, v["unit"], " | ".join(x[:60] for x in v["values"])))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    launches()
    for name in ("forest_staged", "gemm_tn_persistent", "gemm_tn_pair", "gemm_tn_2sm", "attention_varlen", "layernorm_kernel",
                 "embed_layernorm", "nchw_to_s2d", "maxpool3x3s2"):
        # the pair-kernel capture of the LLM prefill shapes keeps its own file (scripts/gpu_llm_ncu.sh)
        o = full(name, out_name="gemm_tn_pair_bert_resnet" if name == "gemm_tn_pair" else None)
        print(name, "ok" if o else "missing")
        if o and name == "forest_staged":
            try:
                def num(key):
                    v = o[key]["values"][0].replace(",", "")
                    u = o[key]["unit"]
                    x = float(v)
                    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                with open(os.path.join(OUT, "forest_traffic.json"), "w") as f:
                    json.dump(dict(dram_bytes_per_launch=traffic, source="%s_ncu_forest_staged.txt" % tag,
                                   note="cold caches (ncu flushes between replays); algorithmic bytes 768448"), f)
            except Exception as ex:  # noqa
                print("traffic:", ex)


def llm():
    """captures made by scripts/gpu_llm_ncu.sh"""
    cmd = "python scripts/llm_bench.py --layers 2 --waves 1 --gen 4 --no-graph: Llama-3-8B shapes, TP 1, batch 32, prompt 512"
    for name in ("skinny_gemm", "llm_attn_decode", "llm_attn_prefill", "llm_reduce_rms", "gemm_tn_pair", "gemm_tn_2sm_llm"):
        print(name, "ok" if full(name, cmd) else "missing")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "llm":
        os.makedirs(OUT, exist_ok=True)
        llm()
        sys.exit(0)
    main()
