"""Attribute ncu per-SASS-instruction counters to CUDA source lines without the GUI.

usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [rows-to-normalise-by]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g -c` line markers of the same kernel by instruction order."""
import csv, io, re, subprocess, sys, collections

rep, cubin, pat = sys.argv[1:4]
norm = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
import os
kf = os.environ.get('NCU_KERNEL')  # e.g. regex:lc_count -- needed when the report holds several kernels
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['-k', kf, '-c', '1'] if kf else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[h]
ci, si, sm = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
body = rows[h + 1:]
end = next((i for i, r in enumerate(body) if 'Source' in r and 'Address' in r), len(body))  # a second view may follow
sass = [(r[si].strip(), int(r[ci] or 0), int(r[sm] or 0)) for r in body[:end] if len(r) > ci]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
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
agg = collections.Counter(); smp = collections.Counter()
for (src, n, s), loc in zip(sass, lines):
    agg[loc] += n; smp[loc] += s
tot, tots = sum(agg.values()), sum(smp.values())
print(f'total warp-instructions {tot} ({tot / norm:.1f} per unit), samples {tots}')
for loc, n in agg.most_common(40):
    print(f'{n / norm:9.1f} inst {100 * n / tot:5.1f}%   samples {100 * smp[loc] / max(1, tots):5.1f}%   {loc}')
