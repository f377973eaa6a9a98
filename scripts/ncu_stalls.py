"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv [--launch-skip k]` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
num = lambda r, k: int(r[ix[k]] or 0)
print("kernel:", rows[0][1][:100] if rows and len(rows[0]) > 1 else "?", "| instructions", len(data), "| samples", sum(num(r, "# Samples") for r in data))
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
    main = sorted([(s[6:], num(r, s)) for s in stalls if num(r, s) > 0], key=lambda kv: -kv[1])[:3]
    print(str(r[-1]).rjust(5), r[ix["# Samples"]].rjust(6), r[ix["Source"]].strip()[:60].ljust(60), main)
agg = {s[6:]: sum(num(r, s) for r in data) for s in stalls}
print("all:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
