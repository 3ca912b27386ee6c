"""Turn gpurun_out ncu artefacts into the committed summaries under profiles/.

    python scripts/summarize_ncu.py <round-tag> [<dir under gpurun_out> [<file suffix>]]

Reads gpurun_out/<dir>/launches_<suffix>.csv and gpurun_out/<dir>/prof_{sep,concat,cube,knn}_<suffix>.ncu-rep, writes
profiles/*_<round-tag>.md and merges per-kernel dram bytes / tensor-pipe activity into profiles/ncu_kernels.json
(read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
sub = sys.argv[2] if len(sys.argv) > 2 else ""
suf = sys.argv[3] if len(sys.argv) > 3 else tag
GO = os.path.join("gpurun_out", sub)

# ---- launch list -> share of the step per kernel
rows = list(csv.reader(open(f"{GO}/launches_{suf}.csv", errors="ignore")))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1e-3)
    a = agg.setdefault(d["Kernel Name"][:110], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/launches_{tag}.md", "w") as f:
    f.write(f"# ncu launch list, `bench.py --steps 2 --warmup 1 --no-cpu --no-extras` ({tag})\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and\n"
            "serialised, so read the SHARES. 3 steps (1 warm-up + 2 timed) plus the kernel-timing leg are in the window.\n\n"
            "| total ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        f.write(f"| {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[0]} | `{k}` |\n")
    f.write(f"\ntotal {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")

# ---- full captures -> key metrics per kernel
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_barrier"]


NCU_JSON = "profiles/ncu_kernels.json"


def summarize(rep, dst, title, how):
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {title} ({tag})\n\n`{how}`\n"
                "(numbers under the profiler are not bench values; CUDA-event timings are in bench.py's JSON line and\n"
                "scripts/bench_components.py)\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')][:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    # per-kernel record for bench.py (first launch of each kernel name wins; online / plain variants kept apart)
    try:
        rec = json.load(open(NCU_JSON))
    except Exception:
        rec = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].split("::")[-1]
        key = short if short not in rec or rec[short].get("source", "").endswith(os.path.basename(dst)) is False else short
        get = lambda m: float(r[hdr.index(m)].replace(",", "")) if m in hdr and r[hdr.index(m)] not in ("", "n/a") else None
        unit = lambda m: units[hdr.index(m)] if m in hdr else ""
        to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        entry = {"dram_bytes": (rd * to_bytes.get(unit("dram__bytes_read.sum"), 1.0) + wr * to_bytes.get(unit("dram__bytes_write.sum"), 1.0))
                 if rd is not None and wr is not None else None,
                 "tensor_pipe_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 "duration": get("gpu__time_duration.sum"), "duration_unit": unit("gpu__time_duration.sum"),
                 "source": f"profiles/{os.path.basename(dst)} (ncu --set full, {tag})", "kernel": name[:160]}
        k = key
        i = 1
        while k in rec and rec[k].get("source") == entry["source"] and rec[k].get("kernel") != entry["kernel"]:
            i += 1
            k = f"{key}#{i}"
        rec.setdefault(k, entry) if rec.get(k, {}).get("source") == entry["source"] else rec.__setitem__(k, entry)
    json.dump(rec, open(NCU_JSON, "w"), indent=1, sort_keys=True)


summarize(f"{GO}/prof_sep_{suf}.ncu-rep", f"profiles/sep_kernels_{tag}.md",
          "ncu --set full, tcgen05 sweep kernels at B = 65536, E = 128",
          "ncu --set full --clock-control none --import-source on -k regex:sep_wsum_tc|sep_stats_tc python bench.py ...")
summarize(f"{GO}/prof_concat_{suf}.ncu-rep", f"profiles/concat_kernels_{tag}.md",
          "ncu --set full, fused concat-critic kernels, 2048 x 4096 pairs",
          "ncu --set full --clock-control none --import-source on -k regex:concat_(fwd|bwd)_kernel -c 2 python scripts/prof_concat.py")
summarize(f"{GO}/prof_cube_{suf}.ncu-rep", f"profiles/cubemlp_kernels_{tag}.md",
          "ncu --set full, CubeMLP tensor-core mixes, [1024,100,3,128] -> 50-3-128 -> 10-3-128",
          "ncu --set full --clock-control none --import-source on -k regex:cubemlp_tc_(fwd|bwd)_kernel -c 8 python scripts/cube_prof.py")
summarize(f"{GO}/prof_knn_{suf}.ncu-rep", f"profiles/knn_kernels_{tag}.md",
          "ncu --set full, k-NN filter / re-rank kernels, 4096 queries x 1M x 128 keys",
          "ncu --set full --clock-control none --import-source on -k regex:knn python scripts/bench_components.py knn")
print(open(f"profiles/launches_{tag}.md").read()[:2500])
