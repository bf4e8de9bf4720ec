"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals, or (--seq A B) the
launches with IDs in [A, B).  usage: python tools/launch_list.py file.csv [--seq A B] [--last-frame]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]
kn, mv, gs, bs, idc = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size'), h.index('ID')
L = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        L.append((int(r[idc]), r[kn], float(r[mv].replace(',', '')) / 1e3, r[gs], r[bs]))
    except ValueError:
        pass
def short(n):
    n = n.replace('hfagp::', '').replace('void ', '')
    return n[:n.index('(')][:48] if '(' in n else n[:48]
if '--seq' in sys.argv:
    a, b = int(sys.argv[sys.argv.index('--seq') + 1]), int(sys.argv[sys.argv.index('--seq') + 2])
    for i, n, t, g, bb in L:
        if a <= i < b:
            print(f'{i:5d} {short(n):50s} {t:9.1f} us  grid {g} block {bb}')
else:
    if '--last-frame' in sys.argv:
        # the last frame starts at the last nchw_to_nhwc launch
        start = max(i for i, n, *_ in L if 'nchw_to_nhwc' in n)
        L = [x for x in L if x[0] >= start]
        print('last frame: launches', len(L), 'first id', start)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for i, n, t, g, bb in L:
        agg[short(n)][0] += 1
        agg[short(n)][1] += t
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print(f'{k:50s} n={v[0]:4d} total={v[1]:9.1f} us  {100 * v[1] / tot:5.1f}%')
    print('total us', round(tot, 1))
