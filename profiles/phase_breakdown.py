import csv, sys, re
src='/root/repo/gsimcli_b200/csrc/k2_grid_csort.cu'
lines=open(src).read().split('\n')
def find(pat, start=0):
    for i in range(start, len(lines)):
        if pat in lines[i]: return i+1
    raise KeyError(pat)
marks=[('load+sanitize','// ---- column c of the tile -> registers'),('pass0','// ---- pass 0'),('pass1','// ---- pass 1'),('scan','// bins of more than HALO+1'),('momfinal','// ---- moment finals'),('pass2','// ---- pass 2'),('generic','// ---- generic path: bitonic'),('bigbins','// ---- big bins'),('winload','// ---- windows:'),('OET','// max-bin-size phases'),('writeback','float* wq = A'),('mode','// ---- planes'),('planes','auto plane_value'),('zero','// histogram back to zero'),('end','template <int IPT, bool MODE, int HALO>\nint cs_launch_one')]
pos=[]
for name,pat in marks[:-1]:
    pos.append((name,find(pat)))
pos.append(('end',find('int cs_launch_one')))
rows=list(csv.reader(open(sys.argv[1])))
cols=int(sys.argv[2])
cur=None; agg={}; hdr=None
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>3 and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<8: continue
    if r[0].isdigit():
        try:
            d=dict(zip(hdr,r))
            agg[(cur,int(r[0]))]=(int(r[4]),int(r[7]),d)
        except Exception as e: pass
ti=sum(v[1] for v in agg.values()); ts=sum(v[0] for v in agg.values())
print('total inst/col %.1f  per elem %.2f ; samples %d'%(ti/cols, ti/cols*32/1000, ts))
stallkeys=[k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
first=pos[0][1]
def bucket(a,b):
    i=s=0; st={}
    for (f,l),(sa,ins,d) in agg.items():
        if f=='k2_grid_csort.cu' and a<=l<b:
            i+=ins; s+=sa
            for k in stallkeys:
                try: st[k]=st.get(k,0)+int(d[k])
                except: pass
    return i,s,st
rows_out=[('setup',1,first)]+[(pos[k][0],pos[k][1],pos[k+1][1]) for k in range(len(pos)-1)]
for name,a,b in rows_out:
    i,s,st=bucket(a,b)
    top=sorted(st.items(), key=lambda kv:-kv[1])[:4]
    print('%-13s inst %5.1f%% (%5.2f/elem) samples %5.1f%%  %s'%(name,100*i/ti,i/cols*32/1000,100*s/ts,' '.join('%s=%.1f%%'%(k[6:],100*v/ts) for k,v in top)))
oth={}
for (f,l),(sa,ins,d) in agg.items():
    if f!='k2_grid_csort.cu':
        o=oth.setdefault(f,[0,0]); o[0]+=ins; o[1]+=sa
for f,o in oth.items(): print(f,'inst %.1f%% samples %.1f%%'%(100*o[0]/ti,100*o[1]/ts))
