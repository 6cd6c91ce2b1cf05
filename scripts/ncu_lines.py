"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by source line: instructions executed + stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
data = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ci:
        continue
    if r[0] in ("", "-"):      # SASS line belonging to the previous source line
        continue
    try:
        n = int(r[ci]); s = int(r[cs])
    except ValueError:
        continue
    data.append((n, s, cur, r[0], r[1].strip()[:100]))
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print("total warp-instructions", tot, "samples", tots)
print("--- by instructions")
for n, s, f, l, src in sorted(data, reverse=True)[:top]:
    print(f"{n:>12} {100*n/tot:5.1f}%  smp {100*s/max(tots,1):5.1f}%  {f}:{l:<4} {src}")
print("--- by stall samples")
for n, s, f, l, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{n:>12} {100*n/tot:5.1f}%  smp {100*s/max(tots,1):5.1f}%  {f}:{l:<4} {src}")
