"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals / shares and (optionally) the
kernel sequence of one minibatch.  Usage: python scripts/launch_summary.py file.csv [--seq]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'ID')
H = rows[hdr]
agg = collections.defaultdict(lambda: [0, 0.0]); seq = []
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    d = dict(zip(H, r))
    name = re.sub(r'^.*::', '', re.sub(r'\(.*', '', d['Kernel Name']))
    try: v = float(d['Metric Value'].replace(',', ''))
    except ValueError: continue
    u = d['Metric Unit']
    v = v / 1000 if u in ('ns', 'nsecond') else (v * 1000 if u in ('ms', 'msecond') else v)
    agg[name][0] += 1; agg[name][1] += v; seq.append((name, v))
tot = sum(v[1] for v in agg.values())
print(f"{len(seq)} launches, {tot/1000:.2f} ms serialised")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:52]:52s} n={v[0]:5d} total={v[1]/1000:9.3f} ms avg={v[1]/v[0]:8.1f} us share={100*v[1]/tot:5.1f}%")
if '--seq' in sys.argv:
    st = [i for i, (n, v) in enumerate(seq) if n.startswith('gather_rows') or n.startswith('gather_meta')]
    a, b = st[len(st) // 2], st[len(st) // 2 + 1]
    for n, v in seq[a:b]: print(f"   {n[:60]:60s} {v:8.1f}")
    print(f"   one minibatch: {b - a} launches, {sum(v for n, v in seq[a:b]):.1f} us")
