"""Hot spots of one kernel from an ncu report with source info: top SASS instructions by stall samples, and the opcode
histogram weighted by executed instructions.   usage: ncu_hot.py report.ncu-rep [top N]"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def num(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
tot_s = sum(num(r[col["# Samples"]]) for r in data)
tot_i = sum(num(r[col["Instructions Executed"]]) for r in data)
print(f"{len(data)} SASS instructions, {tot_s:.0f} samples, {tot_i:.0f} warp instructions executed")
print("--- top by samples: idx samples% execs  long_sb wait branch  source")
order = sorted(range(len(data)), key=lambda i: -num(data[i][col["# Samples"]]))
for i in order[:top]:
    r = data[i]
    print(f"{i:5d} {100*num(r[col['# Samples']])/tot_s:6.2f}% {num(r[col['Instructions Executed']]):11.0f}  "
          f"{num(r[col['stall_long_sb']]):6.0f} {num(r[col['stall_wait']]):6.0f} {num(r[col['stall_branch_resolving']]):6.0f}  {r[col['Source']][:90]}")
hist = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
    op = m.group(2) if m else "?"
    hist[op] += num(r[col["Instructions Executed"]]); samp[op] += num(r[col["# Samples"]])
    thr[op] += num(r[col["Predicated-On Thread Instructions Executed"]])
print("--- opcode: warp-instr%  samples%  thread-instr(pred on)")
for op, c in hist.most_common(22):
    print(f"{op:10s} {100*c/tot_i:6.2f}% {100*samp[op]/tot_s:6.2f}% {thr[op]:.4g}")
# cumulative samples by region of 50 instructions
print("--- samples by 40-instruction region")
for a in range(0, len(data), 40):
    s = sum(num(r[col["# Samples"]]) for r in data[a:a + 40]); e = sum(num(r[col["Instructions Executed"]]) for r in data[a:a + 40])
    if s / tot_s > 0.02: print(f"{a:5d}-{a+39:5d}: {100*s/tot_s:5.1f}% samples, {100*e/tot_i:5.1f}% instr")
