"""Key metrics of every kernel launch in an ncu report as a markdown table: ncu_kernel_table.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
W = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem"),
     ("smsp__inst_executed.sum", "warp inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
     ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %")]
ST = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
print("| kernel | " + " | ".join(n for _, n in W) + " | top stall classes (samples) |")
print("|---|" + "---:|" * len(W) + "---|")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    name = name.replace("void ", "").replace("unnamed>::", "").split("(")[0][:60]
    cells = []
    for k, _ in W:
        if k in hdr:
            v, u = r[hdr.index(k)], rows[1][hdr.index(k)]
            try:
                f = float(v)
                v = f"{f:.0f}" if f >= 100 else f"{f:.1f}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        else:
            cells.append("")
    st = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1]) for h in ST), reverse=True)[:4]
    print(f"| `{name}` | " + " | ".join(cells) + " | " + ", ".join(f"{n} {int(v)}" for v, n in st) + " |")
