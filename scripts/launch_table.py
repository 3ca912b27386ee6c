"""ncu launch-list csv -> per-kernel totals (stdout, markdown).  usage: launch_table.py file.csv [skip_fraction]
skip_fraction: ignore that leading fraction of the launches (warm-up)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr, recs = None, []
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
    recs.append((d["Kernel Name"][:100], v))
recs = recs[int(len(recs) * skip):]
agg = collections.OrderedDict()
for k, v in recs:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| total us | share | launches | us each | kernel |\n|---:|---:|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[0]} | {a[1] / a[0]:.1f} | `{k}` |")
print(f"\ntotal {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
