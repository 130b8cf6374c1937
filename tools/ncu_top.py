"""Top stall lines / key metrics of every kernel in an ncu report (run where ncu is installed; no GPU needed).
usage: python tools/ncu_top.py report.ncu-rep [n_lines] [kernel substring]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 15
only = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    if only and only not in r[h.index("Kernel Name")]:
        continue
    print("---")
    for w in want:
        if w in h:
            print(f"  {w}: {r[h.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen = set()
for k, hi in enumerate(his):
    name = rows[hi - 1][1] if len(rows[hi - 1]) > 1 else ""
    if name in seen or (only and only not in name):
        continue
    seen.add(name)
    hh = rows[hi]
    ix, sx, st = hh.index("Instructions Executed"), hh.index("Source"), hh.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
    end = his[k + 1] - 1 if k + 1 < len(his) else len(rows)
    body = [r for r in rows[hi + 1:end] if len(r) > st]
    tot = sum(int(r[st]) for r in body) or 1
    ninst = sum(int(r[ix]) for r in body)
    print(f"=== {name}: {tot} samples, {ninst} warp instructions")
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[hh[i]] = agg.get(hh[i], 0) + int(r[i])
    print("   stall mix:", ", ".join(f"{k2[6:]} {100 * v / tot:.0f}%" for k2, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    for r in sorted(body, key=lambda r: -int(r[st]))[:topn]:
        reasons = sorted(((int(r[i]), hh[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{int(r[st]):7d} {100 * int(r[st]) / tot:5.1f}% exec={r[ix]:>8s} {r[sx].strip()[:64]:64s} {reasons}")
