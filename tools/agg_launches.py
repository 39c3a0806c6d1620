"""Aggregate an ncu --metrics gpu__time_duration.sum --csv launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('stz::', '')
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} us {n:5d} launches {t / n:8.2f} us/launch {100 * t / tot:5.1f}%  {k[:100]}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
