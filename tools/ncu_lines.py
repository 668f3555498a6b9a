import re, csv, io, subprocess, collections, sys
rep, cubin, func, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
subprocess.run("cd /tmp && rm -rf cub && mkdir cub && cd cub && cuobjdump -xelf all /root/repo/maaco_path_planing_b200/libmpp_b200.so >/dev/null 2>&1 && nvdisasm --print-line-info %s > /tmp/all.sass" % cubin, shell=True)
lines=open('/tmp/all.sass').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+func+':')][0]
cur=None; ins=[]
for l in lines[start+1:]:
    if l.startswith('//------') : break
    m=re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1),16), cur, m.group(2)))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src))); hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}; data=rows[2:]
print(len(ins), len(data))
base=int(data[0][ci['Address']],16)
byaddr={a:(ln,t) for a,ln,t in ins}
cnt=collections.Counter(); smp=collections.Counter()
for r in data:
    off=int(r[ci['Address']],16)-base
    ln,_=byaddr.get(off,(None,None))
    cnt[ln]+=int(r[ci['Instructions Executed']]); smp[ln]+=int(r[ci['# Samples']])
files={}
def text(f,n):
    import glob
    if f not in files:
        c=glob.glob('/root/repo/maaco_path_planing_b200/csrc/'+f)
        files[f]=open(c[0]).read().split('\n') if c else []
    return files[f][n-1].strip()[:95] if n-1 < len(files[f]) else ''
tot=sum(smp.values())
for ln,c in sorted(cnt.items(), key=lambda kv:-smp[kv[0]])[:int(sys.argv[5]) if len(sys.argv)>5 else 40]:
    if ln is None: print('None',c/units, smp[ln]); continue
    f,n=ln
    print(f'{c/units:7.1f} inst  smp {100*smp[ln]/tot:5.1f}%  {f}:{n}  {text(f,n)}')
