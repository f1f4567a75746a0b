"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
path = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.OrderedDict(); cnt = collections.Counter(); seq = []
for i, row in enumerate(csv.DictReader(lines)):
    if i < skip: continue
    full = row['Kernel Name']
    name = re.sub(r'\(.*', '', full)
    name = name if 'TileCfg' in name else re.sub(r'<.*', '', name)
    name = name.replace('void ', '').replace('bobe::', '')[:70]
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    tot[name] = tot.get(name, 0) + v; cnt[name] += 1; seq.append((name, v))
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print(f"{v/1e3:10.3f} ms {100*v/T:5.1f}%  x{cnt[k]:4d}  avg {v/cnt[k]:9.1f} us  {k}")
print(f"total {T/1e3:.3f} ms over {len(seq)} launches")
