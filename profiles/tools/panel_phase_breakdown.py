import csv,sys,re
rows=list(csv.reader(open(sys.argv[1])))
src=open('/root/repo/hl-vae_b200/csrc/kl_stream.cu').read().split('\n')
# phase boundaries from markers in the current source
marks=[('setup','// ---- per-CTA setup'),('P0','// ---- P0, software pipelined'),('P1','// ---- P1: K0xz rows'),('P2 V+rv','// ---- P2: V = B^-1'),('P3a','// ---- P3a:'),('P3b S','// ---- P3b:'),('P4 W','// ---- P4: W = V G'),('P5a+P6','// ---- P5a:'),('P5b','// ---- P5b:'),('flush','// ---- cached components: reduce')]
pos=[]
k0=[i for i,l in enumerate(src) if 'kl_panel_k(const __grid_constant__' in l][0]
for n,m in marks:
    idx=[i for i,l in enumerate(src) if m in l and i>k0][0]+1
    pos.append((n,idx))
end=[i for i,l in enumerate(src) if 'int launch_panel(' in l][0]
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
agg={}
for f,ln,s,i,t in out:
    k=f
    if f=='kl_stream.cu':
        k='kl_stream:other'
        for (n,lo),(n2,hi) in zip(pos,pos[1:]+[('end',end)]):
            if lo<=ln<hi: k=n
    a=agg.setdefault(k,[0,0]); a[0]+=s; a[1]+=i
for k,(s,i) in sorted(agg.items(),key=lambda t:-t[1][0]): print(f'{100*s/tot:5.1f}% smp {100*i/ti:5.1f}% ins {i:>12d} {k}')
print('--- common.cuh lines')
for f,ln,s,i,t in out:
    if f=='common.cuh' and (i>0 or s>0): print(f'{100*s/tot:5.1f}% {100*i/ti:5.1f}% {ln} {t[:80]}')
print('--- top kl_stream lines')
for f,ln,s,i,t in sorted([o for o in out if o[0]=='kl_stream.cu'],key=lambda o:-o[2])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f'{100*s/tot:5.1f}% {100*i/ti:5.1f}% {ln} {t[:90]}')
