"""profiles/r2_roofline_inputs.json from the round's own ncu captures (never typed in by hand; bench.py reads it).

For each captured kernel of one config-2 launch: the FP32 operations it EXECUTES (thread-level, predicated-on, from the
SASS opcode histogram of `ncu --set full --import-source on`: FADD / FMUL = 1 flop, FFMA = 2, FADD2 / FMUL2 = 2,
FFMA2 = 4; MUFU counted separately), the DRAM bytes ncu saw, and the pipe / issue utilisation figures.

usage: roofline_from_ncu.py out.json name=report.ncu-rep [name=report.ncu-rep ...]
"""
import collections, csv, io, json, re, subprocess, sys

FLOPS = {"FADD": 1, "FMUL": 1, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FFMA2": 4, "FMNMX": 1, "FSET": 1, "FSETP": 1, "FSEL": 0}
ARITH = ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def kernel(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    thr = collections.Counter()
    warp = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
        op = m.group(2) if m else "?"
        thr[op] += num(r[col["Predicated-On Thread Instructions Executed"]])
        warp[op] += num(r[col["Instructions Executed"]])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rr[0], rr[2]))
    g = lambda k: num(d.get(k, "0"))   # noqa: E731
    unit = dict(zip(rr[0], rr[1]))
    def byts(k):
        v, u = g(k), unit.get(k, "")
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    flops = sum(thr[o] * FLOPS[o] for o in ARITH)
    return {
        "kernel": d.get("Kernel Name", "?"), "report": rep.split("/")[-1],
        "executed_flops_per_launch": flops, "executed_fp32_compare_minmax_per_launch": thr["FMNMX"] + thr["FSET"] + thr["FSETP"],
        "mufu_per_launch": thr["MUFU"], "thread_instructions_per_launch": sum(thr.values()), "warp_instructions_per_launch": sum(warp.values()),
        "opcode_thread_instructions": {o: thr[o] for o in ARITH + ("MUFU", "IMAD", "MOV", "LDG", "STG", "LDS", "STS", "LDGSTS")},
        "dram_bytes_per_launch": byts("dram__bytes_read.sum") + byts("dram__bytes_write.sum"),
        "duration_us_under_ncu": g("gpu__time_duration.sum") * {"usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "us": 1, "ms": 1e3}.get(unit.get("gpu__time_duration.sum", "usecond"), 1),
        "issue_slots_busy_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_busy_pct": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": g("launch__registers_per_thread"),
    }


def main():
    out = sys.argv[1]
    res = {"_how": "scripts/roofline_from_ncu.py over `ncu --set full --clock-control none --import-source on` captures of one "
                   "config-2 launch (1 024 utterances x 220 476 samples); flops = predicated-on thread instructions x "
                   "{FADD 1, FMUL 1, FFMA 2, FADD2 2, FMUL2 2, FFMA2 4}"}
    step = 0.0
    for a in sys.argv[2:]:
        name, rep = a.split("=", 1)
        mult = 1
        if "*" in name:
            name, m = name.split("*")
            mult = int(m)
        res[name] = kernel(rep)
        res[name]["launches_per_step"] = mult
        step += res[name]["executed_flops_per_launch"] * mult
    res["step_executed_flops"] = step
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: v[kk] for kk in ("executed_flops_per_launch", "dram_bytes_per_launch", "issue_slots_busy_pct", "fma_pipe_busy_pct")}) for k, v in res.items() if k != "_how"}, indent=1))


if __name__ == "__main__":
    main()
