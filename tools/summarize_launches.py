import csv, collections, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    name = name.split('(')[0][-60:]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit == 'ns': v /= 1000
    elif unit == 'ms': v *= 1000
    agg.setdefault(name, []).append(v)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-62s n=%3d  avg %9.1f us  total %10.1f us" % (k, len(v), sum(v) / len(v), sum(v)))
