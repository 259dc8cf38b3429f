"""Per-phase instruction count and warp-state samples (by stall reason) of the window backward from an ncu report captured
with --import-source on.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:msda_bwd_d32_win > src.csv
       python profiles/stall_budget.py src.csv
The line ranges below are those of csrc/msda_d32_win.cuh at the end of round 2.
"""
import csv,sys,collections
path=sys.argv[1]
cur=None; hdr=None
phases=[(36,453,'top/setup'),(454,515,'F0 loads/init'),(516,536,'F1 decode/bbox'),(537,625,'F2 alloc/staging/records/counts'),(626,663,'F3 scan'),(664,684,'F4 place/wait'),(685,725,'S setup'),(726,831,'S sorted pass'),(837,912,'D direct pass'),(913,972,'W write-out')]
special={('msda_d32_win.cuh',310):'S row loads',('msda_d32_win.cuh',311):'S row loads',('msda_d32_win.cuh',367):'flush red',('msda_d32_win.cuh',369):'flush red',('msda_d32_win.cuh',160):'F2 alloc/staging/records/counts',('msda_d32_win.cuh',162):'F2 alloc/staging/records/counts'}
agg=collections.defaultdict(lambda: collections.Counter())
inst=collections.Counter()
for r in csv.reader(open(path)):
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if r[0] in ("","Function Name"): continue
    try: line=int(r[0])
    except: continue
    name=special.get((cur,line))
    if name is None:
        if cur=='msda_d32_win.cuh':
            name='other win'
            for a,b,n in phases:
                if a<=line<=b: name=n;break
        else: name='inl:'+cur
    ie=hdr.index("Instructions Executed")
    try: inst[name]+=int(r[ie])
    except ValueError: continue
    for i,h in enumerate(hdr):
        if h.startswith('stall_') and '(Not Issued)' not in h:
            try: agg[name][h]+=int(r[i])
            except: pass
tot=sum(sum(c.values()) for c in agg.values())
print('total samples',tot)
for name,c in sorted(agg.items(), key=lambda kv:-sum(kv[1].values())):
    s=sum(c.values())
    top=', '.join('%s %.1f'%(k.replace('stall_',''),100*v/tot) for k,v in c.most_common(5))
    print('%-34s inst %6.2fM  samples %5.1f%%  | %s'%(name,inst[name]/1e6,100*s/tot,top))
