"""Per-kernel SASS opcode summary of the built libb200serve.so (no GPU needed: cuobjdump reads the cubin).
Writes profiles/r02_sass_opcode_summary.txt -- the evidence that the contraction kernels are Blackwell-native
(UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, UCGABAR = cluster barrier; HMMA = legacy mma.sync).

    python scripts/sass_summary.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clearml_serving_b200", "libb200serve.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR",
        "HMMA", "IMMA", "FADD", "DADD", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "BAR"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcode_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()   # noqa: E731
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
    lines = ["# per-kernel SASS opcode counts of clearml_serving_b200/libb200serve.so (cuobjdump -sass, sm_100a)",
             "# UTCHMMA=tcgen05.mma  .2CTA=cta_group::2  LDTM=tcgen05.ld  UTMALDG/UTMASTG=TMA load/store  UBLKCP=cp.async.bulk",
             "# SYNCS=mbarrier  UCGABAR=cluster barrier  HMMA=legacy mma.sync", ""]
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", demangle(name))
        hot = ["{}={}".format(k, c[k]) for k in KEYS if c.get(k)]
        lines.append("{:<70s} instr={:<6d} {}".format(short[:70], c["_total"], " ".join(hot)))
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:4]))
    print("{} kernels -> {}".format(len(kernels), out_path))


if __name__ == "__main__":
    main()
