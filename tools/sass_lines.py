"""Print the SASS (with executed counts per unit) attributed to a source line range.
usage: sass_lines.py <ncu source csv> <nvdisasm -g text> <units> <file> <lo> <hi>"""
import csv, re, sys
src_csv, sass, units, f, lo, hi = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
rows = list(csv.reader(open(src_csv))); hdr = rows[1]
iS, iE = hdr.index('Source'), hdr.index('Instructions Executed')
inst = [(r[iS], int(r[iE])) for r in rows[2:] if r[iE].isdigit()]
lines = []; cur = None
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
for i, ((s, n), ln) in enumerate(zip(inst, lines)):
    if ln and ln[0] == f and lo <= ln[1] <= hi and n > 0:
        print("%5d %s:%d %6.2f  %s" % (i, ln[0][:12], ln[1], n / units, s[:90]))
