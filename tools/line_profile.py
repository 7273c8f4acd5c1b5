"""Join ncu's per-SASS-instruction counts with nvdisasm -g line info: executed warp instructions per source line.
usage: line_profile.py <ncu source csv> <nvdisasm -g text of the same function> <units> [top]"""
import csv, re, sys, collections
src_csv, sass, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
inst = []
for r in rows[2:]:
    try: inst.append((r[iS], int(r[iE]), int(r[iSm] or 0)))
    except Exception: pass
lines = []
cur = None
inl = None
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
assert len(lines) == len(inst), (len(lines), len(inst))
agg = collections.Counter(); samp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for (s, n, sm), ln in zip(inst, lines):
    agg[ln] += n; samp[ln] += sm
    t = s.split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[ln][op] += n
tot = sum(agg.values()); ts = sum(samp.values())
print("total", tot, "per unit", tot / units)
for ln, n in agg.most_common(top):
    print("%-22s %8.1f /unit (%4.1f%%) samples %4.1f%%  %s" % ("%s:%d" % ln if ln else "?", n / units, 100 * n / tot, 100 * samp[ln] / max(ts, 1),
          " ".join("%s:%.0f" % (o, c / units) for o, c in ops[ln].most_common(6))))
# ---- totals per source range -----------------------------------------------------------------
if len(sys.argv) > 5:
    for spec in sys.argv[5:]:
        f, lo, hi = spec.split(':')
        n = sum(v for (k, v) in agg.items() if k and k[0] == f and int(lo) <= k[1] <= int(hi))
        print("range %s  %.1f /unit" % (spec, n / units))
