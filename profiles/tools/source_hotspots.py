import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; data = []
for r in rows:
    if r and r[0] in ('File Name','File Path'):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or not r or not r[0].strip().isdigit():
        continue
    si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
    try: s = int(r[si]); ins = int(r[ii])
    except: continue
    data.append((s, ins, cur, r[0], r[1].strip()[:100]))
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print("total samples", tot, "total inst", toti)
for s, ins, f, ln, src in sorted(data, key=lambda t: -t[0])[:n]:
    print(f"{100*s/max(tot,1):5.1f}% smp {100*ins/max(toti,1):5.1f}% ins  {f}:{ln:>4s}  {src}")
