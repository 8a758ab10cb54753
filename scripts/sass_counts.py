"""Per-kernel SASS opcode counts of libgrail_cuda.so (static, from cuobjdump): the instructions that show what the
kernels are made of -- packed FP32 (FFMA2 / FADD2 / FMUL2), MUFU, async copies (LDGSTS = cp.async, UBLKCP = TMA bulk),
mbarrier traffic (SYNCS), 256-bit global accesses, shuffles.   usage: sass_counts.py [lib] > profiles/r2_sass_counts.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "grail-rs_b200/libgrail_cuda.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(2)
        counts[cur][op] += 1
        if op in ("LDG", "STG") and ".256" in (m.group(3) or ""):
            counts[cur][op + ".256"] += 1
        counts[cur]["_total"] += 1
cols = ["_total", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "LDG.256", "STG.256", "LDS", "STS", "SHFL", "BAR", "DADD"]
print("# static SASS instruction counts per kernel (cuobjdump -sass " + lib + ")")
print(f"{'kernel':60s} " + " ".join(f"{c:>8s}" for c in cols))
for k, c in counts.items():
    print(f"{k[:60]:60s} " + " ".join(f"{c[x]:8d}" for x in cols))
print("# no UTMALDG / UBLKCP: the walks' data path is 32-byte sectors per lane with lane-dependent ranges (cp.async = LDGSTS);")
print("# per-lane TMA bulk copies were built and measured in the first version of grail_phase.cuh (DESIGN.md 4a / 5).")
