"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS/source lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
si = [hdr.index(h) for h in stalls]


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


data = [r for r in rows[2:] if len(r) == len(hdr) and r[i_s] != '# Samples']
tot = sum(num(r[i_s]) for r in data)
print('rows', len(data), 'total samples', tot)
agg = {h: 0 for h in stalls}
for r in data:
    for h, i in zip(stalls, si):
        agg[h] += num(r[i])
s = max(1, sum(agg.values()))
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print('  %-24s %8d %5.1f%%' % (h, v, 100 * v / s))
for r in sorted(data, key=lambda r: -num(r[i_s]))[:topn]:
    st = sorted(((num(r[i]), h) for h, i in zip(stalls, si)), reverse=True)[:2]
    print(r[i_s].rjust(7), r[i_ex].rjust(10), r[i_src][:100].ljust(100), st)
