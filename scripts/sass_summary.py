"""SASS evidence for the tcgen05 / TMEM / TMA kernels: per kernel, the counts of UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG (TMA loads), UTCBAR (tcgen05.commit), FFMA2 / FADD2 (packed fp32) and a short excerpt around the
first UTCHMMA.   python scripts/sass_summary.py > profiles/sass_tcgen05_r2.md"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "mimrl_b200/lib/libmimrl_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = []
        continue
    if cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        kern[cur].append(line.rstrip())
def demangle(sym):
    full = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
    return full.replace("(anonymous namespace)::", "").replace("void ", "").replace("mimrl::", "").split("(")[0]
ops = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "FFMA2", "FADD2", "HMMA")
print("# SASS evidence, `cuobjdump -sass mimrl_b200/lib/libmimrl_b200.so` (sm_100a), round 2\n")
print("tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR, "
      "fma/add.f32x2 -> FFMA2/FADD2; no legacy HMMA anywhere.\n")
print("| kernel | instructions | " + " | ".join(ops) + " |\n|---|---:|" + "---:|" * len(ops))
rows = []
for k, lines in kern.items():
    cnt = {o: sum(1 for l in lines if re.search(r"\b" + o + r"\b", l.split("*/", 1)[1] if "*/" in l else l)) for o in ops}
    if cnt["UTCHMMA"] or cnt["UTMALDG"] or cnt["LDTM"]:
        rows.append((demangle(k), len(lines), cnt, lines))
for name, n, cnt, _ in sorted(rows, key=lambda r: -r[2]["UTCHMMA"]):
    print(f"| `{name[:90]}` | {n} | " + " | ".join(str(cnt[o]) for o in ops) + " |")
print(f"\ntotals: " + ", ".join(f"{o} {sum(r[2][o] for r in rows)}" for o in ops))
for name, n, cnt, lines in rows:
    if "sep_wsum_tc_kernel<0, true>" in name or "concat_fwd_kernel<true>" in name:
        i = next(j for j, l in enumerate(lines) if "UTCHMMA" in l)
        print(f"\n## excerpt: `{name}` around its first UTCHMMA\n\n```")
        for l in lines[max(0, i - 6): i + 8]:
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l))
        print("```")
