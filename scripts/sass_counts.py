"""SASS evidence for profiles/: per-kernel counts of the Blackwell-specific mnemonics in libyelprec_b200.so
(tcgen05 MMA = UTCHMMA/UTCQMMA..., TMEM loads = LDTM, TMA = UTMALDG / UBLKCP, packed fp32 FMA = FFMA2, vector reductions = REDG, atomics = ATOMG,
FP64 = DFMA). Usage: python scripts/sass_counts.py > profiles/r02_sass_counts.txt   (runs without a GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "yelprecommendation_b200", "lib", "libyelprec_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "REDG", "ATOMG", "ATOMS", "MEMBAR", "DFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
cur, counts, archs = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)[:80]
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        archs.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        for p in pats:
            if op == p:
                counts[cur][p] += 1
print(f"# {os.path.relpath(lib, ROOT)}: architectures {sorted(archs)}; {len(counts)} kernels")
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
print("# totals: " + ", ".join(f"{p} {tot[p]}" for p in pats if tot[p]))
print(f"{'kernel':82s} {'instr':>7s} " + " ".join(f"{p:>8s}" for p in pats[:13]))
for k, c in counts.items():
    print(f"{k:82s} {c['_total']:7d} " + " ".join(f"{c[p]:8d}" for p in pats[:13]))
