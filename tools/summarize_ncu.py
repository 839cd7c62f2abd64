"""Summarise an ncu --metrics gpu__time_duration.sum launch list (csv) into a per-kernel table."""
import collections, csv, re, sys
path, out = sys.argv[1], sys.argv[2]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name']; v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
    if name.startswith('void native::') or 'elementwise' in name[:60]:
        inner = re.findall(r'native::(?:<unnamed>::|\(anonymous namespace\)::)?(\w+)', name)
        short = 'aten: ' + ' / '.join(list(dict.fromkeys(inner))[:3])
    else:
        short = re.sub(r'\(.*', '', name)[:90]
    tot[short] += v; cnt[short] += 1
total = sum(tot.values())
ours = sum(v for k, v in tot.items() if 'b2::' in k)
with open(out, 'w') as f:
    f.write("# per-kernel device time of ONE full-size pair-iteration (ncu gpu__time_duration.sum, --clock-control none;\n")
    f.write("# cold-cache, serialised launches: compare SHARES, not absolutes)\n")
    f.write("total %.2f ms over %d launches; libb2attack kernels %.2f ms (%.1f%%)\n\n" % (total, sum(cnt.values()), ours, 100 * ours / total))
    f.write("%10s %7s %6s  %s\n" % ("ms", "share", "count", "kernel"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v / total < 0.001:
            continue
        f.write("%10.3f %6.1f%% %6d  %s\n" % (v, 100 * v / total, cnt[k], k))
print(open(out).read()[:1500])
