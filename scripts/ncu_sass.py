"""List SASS (address order) with per-frame executed counts for source lines [lo, hi] of a file from an ncu source CSV."""
import csv, sys
path, fname, lo, hi, frames = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
rows = list(csv.reader(open(path)))
hdr = None; cur = None; line = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] not in ("", "-"):
        try: line = int(r[0])
        except ValueError: pass
        continue
    if r[2] == "..." or cur != fname or line is None or not (lo <= line <= hi): continue
    try: out.append((int(r[2], 16), line, r[3].strip(), int(r[7]), int(r[6])))
    except ValueError: pass
out.sort()
tot = 0
for a, l, s, n, smp in out:
    tot += n
    print(f"{a & 0xfffff:05x} L{l:<4} {n/frames:7.1f} smp {smp:5d}  {s}")
print("total/frame", tot / frames)
