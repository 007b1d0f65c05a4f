"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel share table kept under profiles/."""
import csv, sys, collections
path, title = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    name = r[ki].split('(')[0][:70]
    tot[name] += float(r[vi].replace(',', '')) / 1e6
    cnt[name] += 1
s = sum(tot.values())
print('# %s\n' % title)
print('| kernel | launches | total ms | share |\n|---|---|---|---|')
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('| %s | %d | %.3f | %.1f %% |' % (n, cnt[n], v, 100 * v / s))
