import csv, io, sys
t = open(sys.argv[1]).read(); t = t[t.index('"ID"'):]
rows = list(csv.DictReader(io.StringIO(t)))
d = {}
for r in rows:
    d.setdefault((r['ID'], r['Kernel Name'].split('(')[0][-60:]), {})[r['Metric Name']] = r['Metric Value']
keys = list(d)
names = []
for k in keys:
    for m in d[k]:
        if m not in names: names.append(m)
print(' ' * 58, *['%14s' % k[1].split('::')[-1][:14] for k in keys])
for m in names:
    print('%-58s' % m[-58:], *['%14s' % d[k].get(m, '') for k in keys])
