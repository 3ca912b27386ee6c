"""Top source lines of an ncu report by stall samples: ncu_lines.py report.ncu-rep [launch index] [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
def col(n): return hdr.index(n)
r = rows[2 + which]
for n in ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    if n in hdr: print(f"{n}: {r[col(n)][:90]} {rows[1][col(n)]}")
st = {h: r[col(h)] for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
print({k.split("stalled_")[1]: v for k, v in sorted(st.items(), key=lambda kv: -float(kv[1] or 0))[:8]})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "File Path"] + [len(rows)]
# the sections of one launch are consecutive (one per source file); a new launch starts when the function name changes
# or a (function, file) pair repeats
launches, cur, seen = [], [], set()
for s_ in range(len(secs) - 1):
    a, b = secs[s_], secs[s_ + 1]
    f = rows[a][1]
    fn = rows[a + 1][1] if len(rows[a + 1]) > 1 else ""
    if cur and ((fn, f) in seen or fn != cur_fn):
        launches.append(cur); cur, seen = [], set()
    cur.append((a, b, f)); seen.add((fn, f)); cur_fn = fn
if cur: launches.append(cur)
sel = launches[which]
agg = []; tot = 0; toti = 0
for a, b, f in sel:
    h = rows[a + 2]
    if "Line No" not in h: continue
    iL, iS, iI = h.index("Line No"), h.index("# Samples"), h.index("Instructions Executed")
    for r in rows[a + 3:b]:
        try: ln = int(r[iL]); smp = int(r[iS] or 0); ins = int(r[iI] or 0)
        except Exception: continue
        if smp or ins: agg.append((f.split("/")[-1], ln, smp, ins, r[1][:100]))
        tot += smp; toti += ins
agg.sort(key=lambda t: -t[2])
print("samples", tot, "inst", toti)
for t in agg[:top]: print(f"{t[0]}:{t[1]:4d} samp {100*t[2]/max(tot,1):4.1f}% inst {100*t[3]/max(toti,1):4.1f}% | {t[4]}")
