"""Attribute ncu per-SASS-instruction counters to CUDA source lines without the GUI.

usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [rows-to-normalise-by]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g -c` line markers of the same kernel by instruction order."""
import csv, io, re, subprocess, sys, collections

rep, cubin, pat = sys.argv[1:4]
norm = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
import os
kf = os.environ.get('NCU_KERNEL')  # substring of the demangled kernel name -- needed when the report holds several kernels
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if kf:  # keep the first section whose "Kernel Name" row matches
    k0 = next(i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name' and kf in r[1])
    k1 = next((i for i in range(k0 + 1, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
    rows = rows[k0:k1]
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[h]
ci, si, sm = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
body = rows[h + 1:]
end = next((i for i, r in enumerate(body) if 'Source' in r and 'Address' in r), len(body))  # a second view may follow
sass = [(r[si].strip(), int(r[ci] or 0), int(r[sm] or 0)) for r in body[:end] if len(r) > ci]
stall_by_row = [{n: int(r[i] or 0) for i, n in stall_cols} for r in body[:end] if len(r) > ci]
dis = subprocess.run(['nvdisasm', '-gi' if os.environ.get('NCU_OUTER') else '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
# locate the kernel's text section
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and pat in l)
lines, cur, inl = [], None, None
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
    if m:
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
agg = collections.Counter(); smp = collections.Counter(); why = collections.defaultdict(collections.Counter)
for (src, n, s), loc, st in zip(sass, lines, stall_by_row):
    agg[loc] += n; smp[loc] += s
    why[loc].update(st)
tot, tots = sum(agg.values()), sum(smp.values())
print(f'total warp-instructions {tot} ({tot / norm:.1f} per unit), samples {tots}')
order = smp.most_common(40) if os.environ.get('NCU_SORT') == 'samples' else agg.most_common(40)
for loc, _ in order:
    n = agg[loc]
    top = ', '.join(f'{k[6:]} {v}' for k, v in why[loc].most_common(3) if v)
    print(f'{n / norm:9.1f} inst {100 * n / tot:5.1f}%   samples {100 * smp[loc] / max(1, tots):5.1f}%   {loc}   [{top}]')
