"""SASS evidence: per-kernel instruction histograms of the in-tree library (cuobjdump -sass), with the mnemonics that prove
what a kernel is built on (UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld/st, UTMALDG = TMA tensor load,
UBLKCP = cp.async.bulk (1-D TMA copy), SYNCS = mbarrier, LDGSTS = cp.async, HSET2 / HSETP2 = packed 16-bit compares, REDUX, ATOMS / RED) counted explicitly.
usage: python tools/sass_histogram.py > profiles/r3/sass_histogram.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ecoflap_b200", "libecoflap_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "HSET2", "HSETP2", "IDP4A", "REDUX", "ATOMS",
       "RED", "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "VOTE", "BAR", "FMUL", "HFMA2", "HADD2", "LOP3", "MUFU")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.split("\n")
print(f"# {os.path.relpath(LIB, ROOT)}: {len(hist)} kernels (cuobjdump -sass, sm_100a)\n")
for (k, h), name in zip(hist.items(), demangle):
    if not name.startswith("ecf::") and "ecf::" not in name:
        continue
    total = sum(h.values())
    keys = "  ".join(f"{m}={h[m]}" for m in KEY if h[m])
    print(f"{name[:150]}\n    {total} instructions   {keys}")
    top = ", ".join(f"{m} {c}" for m, c in h.most_common(8))
    print(f"    top: {top}\n")
