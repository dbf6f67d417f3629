"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; ix = {k: j for j, k in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).split('::')[-1]
    v = float(r[ix['Metric Value']].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ix['Metric Unit']], 1e-6)
    key = name if len(sys.argv) < 3 else (name, r[ix['Grid Size']])
    agg[key][0] += 1; agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{str(k):60s} n={v[0]:5d} ms={v[1]:9.3f} {v[1] / tot * 100:5.1f}%  avg_us={v[1] / v[0] * 1000:8.1f}")
