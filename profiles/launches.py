"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ik]).split('::')[-1][:60]
    v = float(r[iv].replace(',', ''))
    v = v/1e3 if r[iu] == 'ns' else v*1e3 if r[iu] == 'ms' else v
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'{k:62s} n={len(v):4d} avg={sum(v)/len(v):10.1f} us  total={sum(v)/1e3:9.2f} ms  share={100*sum(v)/tot:5.1f}%')
