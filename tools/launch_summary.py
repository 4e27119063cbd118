"""Dev tool: per-kernel summary of one frame from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(frame = from one GenSimple/GenJittered launch to the one `chunks` launches later)."""
import csv, collections, re, sys
path = sys.argv[1]; chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lines = [l for l in open(path) if not l.startswith('==')]
rows = []
for x in csv.DictReader(lines):
    try: rows.append((x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x['Metric Unit']))
    except Exception: pass
gen = [i for i, (k, _, _) in enumerate(rows) if 'Gen' in k]
start = gen[0]; end = gen[chunks] if len(gen) > chunks else len(rows)
fr = rows[start:end]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, v, u in fr:
    v = v / 1000 if u in ('ns', 'nsecond') else v
    k = re.sub(r'\(.*', '', k)
    a = agg[k]; a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f'{len(fr)} launches in one frame, sum of kernel durations {tot:.1f} us (cold-cache, serialised: compare SHARES)')
for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{a[1]:10.1f} us {100*a[1]/tot:5.1f}% n={a[0]:4d} avg={a[1]/a[0]:8.1f} max={a[2]:8.1f}  {k[:120]}")
