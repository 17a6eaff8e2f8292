import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None;cur=None;out=[]
for r in rows:
    if r and r[0] in('File Name','File Path'): cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr is None or not r or not r[0].strip().isdigit(): continue
    si=hdr.index('# Samples'); ii=hdr.index('Instructions Executed')
    try: out.append((cur,int(r[0]),int(r[si]),int(r[ii]),r[1][:100]))
    except: pass
tot=sum(o[2] for o in out); ti=sum(o[3] for o in out)
print('samples',tot,'instr',ti)
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
for f,ln,s,i,t in sorted(out,key=lambda o:-o[3])[:n]:
    print(f'{100*s/tot:5.1f}% smp {100*i/ti:5.1f}% ins {f}:{ln} {t[:95]}')
