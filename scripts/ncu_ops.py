"""Executed warp-instructions per frame by opcode from an `ncu --page source --print-source cuda,sass --csv` dump (SASS rows de-duplicated by address)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1]))); frames = float(sys.argv[2])
hdr = None; ins = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] not in ("", "-") or r[2] == "...": continue
    try: a = int(r[2], 16); n = int(r[7])
    except ValueError: continue
    t = r[3].split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
    ins[a] = (n, op.split(".")[0])
c = collections.Counter()
for n, op in ins.values(): c[op] += n
print("total %.0f per frame: " % (sum(c.values()) / frames) + ", ".join("%s %.0f" % (op, n / frames) for op, n in c.most_common(28)))
