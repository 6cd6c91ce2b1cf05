"""Shared-memory wavefronts per source line from an `ncu --page source --print-source cuda,sass --csv` dump: total, ideal, excess."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None; cur = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; iw = hdr.index("L1 Wavefronts Shared"); ii = hdr.index("L1 Wavefronts Shared Ideal"); ie = hdr.index("L1 Wavefronts Shared Excessive"); ix = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iw or r[0] in ("", "-"): continue
    try: w = int(r[iw]); i = int(r[ii]); e = int(r[ie]); n = int(r[ix])
    except ValueError: continue
    if w: data.append((w, i, e, n, cur, r[0], r[1].strip()[:110]))
tot = sum(d[0] for d in data); ex = sum(d[2] for d in data)
print("shared wavefronts", tot, "excess", ex)
for w, i, e, n, f, l, s in sorted(data, reverse=True)[:top]:
    print(f"{w:>10} ideal {i:>10} excess {e:>9} inst {n:>9}  {f}:{l:<4} {s}")
