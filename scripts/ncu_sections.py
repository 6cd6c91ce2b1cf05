"""Aggregate an ncu source-page CSV (--print-source cuda,sass) of the warp feature kernel into pipeline phases.
SASS rows are de-duplicated by address (inlined code is listed under the callee line AND the call site) and attributed to the
phase of the kernel-file line (syg_frame_warp.cuh) that encloses them in address order."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
W = "sygnals_b200/csrc/syg_frame_warp.cuh"
def grep_line(fn, pat):
    for i, l in enumerate(open(fn), 1):
        if pat in l: return i
    return 10**9
marks = [(grep_line(W, p), n) for p, n in [("framing + window + time-domain", "load+window"), ("pass 1: radix E", "fft"),
         ("real split -> |X[k]|^2", "split+power"), ("time-domain features (unwindowed", "time feats"),
         ("per-frame spectral statistics", "spec stats"), ("---- mel energies", "mel"), ("---- spectral contrast", "contrast")]]
def phase(l):
    name = "prologue"
    for ln, n in marks:
        if l >= ln: name = n
    return name
hdr = None; cur = None; line = None
ins = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] not in ("", "-"):
        try: line = int(r[0])
        except ValueError: pass
        continue
    if r[2] == "...": continue
    try: a = int(r[2], 16); n = int(r[7]); s = int(r[6])
    except ValueError: continue
    d = ins.setdefault(a, {"n": n, "s": s, "k": None, "op": r[3].split()[0] if r[3].split() else ""})
    if cur == "syg_frame_warp.cuh": d["k"] = phase(line)
# instructions that never appear under a kernel-file line inherit the phase of the previous address
agg = {}; last = "prologue"
for a in sorted(ins):
    d = ins[a]
    if d["k"] is None: d["k"] = last
    last = d["k"]
    x = agg.setdefault(d["k"], [0, 0, 0]); x[0] += d["n"]; x[1] += d["s"]; x[2] += 1
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
for k, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:14s} {n:>12} {100*n/tot:5.1f}%  stall samples {100*s/max(ts,1):5.1f}%  static {c:5d}" + (f"   {n/frames:7.0f} inst/frame" if frames else ""))
print(f"{'total':14s} {tot:>12}" + (f"  static {sum(v[2] for v in agg.values())}   {tot/frames:7.0f} inst/frame" if frames else ""))
