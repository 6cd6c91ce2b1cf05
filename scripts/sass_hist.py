"""Static SASS instruction histogram of one kernel by source region (nvdisasm -g dump)."""
import re, collections, sys
lines = open(sys.argv[1]).read().split('\n')
pat = sys.argv[2]
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and pat in l][0]
end = [i for i, l in enumerate(lines) if i > start and '.section' in l and '.text.' in l]
end = end[0] if end else len(lines)
cur = None; cnt = collections.Counter()
for ln in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', ln): cnt[cur] += 1
print('total instr', sum(cnt.values()))
grp = collections.Counter()
for (f, l), v in cnt.items(): grp[(f, l // 20 * 20)] += v
for k, v in grp.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25): print(v, k)
