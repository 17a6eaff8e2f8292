import csv, sys
# attribute pc samples of a source-page CSV to line ranges of a file
rows = list(csv.reader(open(sys.argv[1])))
target = sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    name, lo, hi = a.split(':')
    ranges.append((name, int(lo), int(hi)))
cur = None; hdr = None
agg = {}; other = {}
tot = 0
for r in rows:
    if r and r[0] in ('File Name', 'File Path'):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or not r or not r[0].strip().isdigit():
        continue
    si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
    try: s = int(r[si]); ins = int(r[ii])
    except: continue
    tot += s
    ln = int(r[0])
    key = None
    if cur == target:
        for name, lo, hi in ranges:
            if lo <= ln <= hi: key = name; break
        if key is None: key = f'{cur}:other'
    else:
        key = f'{cur}'
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += ins
print('total samples', tot)
for k, (s, ins) in sorted(agg.items(), key=lambda t: -t[1][0]):
    print(f'{100*s/max(tot,1):5.1f}% smp  {ins:>12d} ins  {k}')
