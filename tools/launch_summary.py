"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total us, share."""
import csv, sys, collections, re

def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ni, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ni])
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3
        a[2].add((r[gi], r[bi]))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot/1e3:.3f} ms of kernel time (ncu: serialised, cold caches -- compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}  grids")
    for name, (n, us, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {us:12.1f} {us/n:9.2f} {us/tot:7.1%}  {len(grids)} distinct (grid, block)")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
