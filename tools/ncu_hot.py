"""Group SASS of each kernel in an ncu report by execution count to find the hot regions."""
import csv, subprocess, sys, io
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
txt = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern=None; hdr=None; data={}
for r in rows:
    if r and r[0]=='Kernel Name': kern=r[1][:50]; data[kern]=[]; hdr=None; continue
    if r and r[0]=='Address': hdr=r; continue
    if hdr and len(r)>=len(hdr)-2: data[kern].append(dict(zip(hdr,r)))
for kname,v in data.items():
    tot=sum(int(d['Instructions Executed']) for d in v); smp_tot=sum(int(d['# Samples']) for d in v)
    print('=====',kname,'inst',tot,'samples',smp_tot)
    i=0
    while i<len(v):
        n=int(v[i]['Instructions Executed']); j=i
        while j<len(v) and int(v[j]['Instructions Executed'])==n: j+=1
        cnt=j-i
        if n*cnt>tot*thresh:
            ops={}
            for d in v[i:j]:
                t=d['Source'].strip().split()
                op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op]=ops.get(op,0)+1
            smp=sum(int(d['# Samples']) for d in v[i:j])
            print(f"  lines {i:3d}-{j-1:3d} n={cnt:3d} exec={n/1e6:7.2f}M inst_share={n*cnt/tot*100:5.1f}% sample_share={smp/smp_tot*100:5.1f}% thr={v[i]['Avg. Threads Executed']:>4s} ", dict(sorted(ops.items(), key=lambda x:-x[1])[:7]))
        i=j
