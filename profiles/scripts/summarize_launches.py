"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: summarize_launches.py FILE.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
k, v, g = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
tot, cnt, grids = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in rows:
    name = r[k].split("(")[0].split("::")[-1]
    tot[name] += float(r[v].replace(",", "")); cnt[name] += 1; grids[name][r[g]] += 1
allns = sum(tot.values())
print(f"{'kernel':28s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}  grid sizes (count)")
for name, ns in tot.most_common():
    gs = ", ".join(f"{a}x{b}" for a, b in grids[name].most_common(4))
    print(f"{name:28s} {cnt[name]:8d} {ns / 1e6:10.3f} {100 * ns / allns:6.1f}% {ns / cnt[name] / 1e3:9.1f}  {gs}")
