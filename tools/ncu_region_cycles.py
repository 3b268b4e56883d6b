import sys,subprocess
sys.path.insert(0,'/root/repo/tools')
from ncu_by_line import parse_disasm, parse_ncu
tag=sys.argv[1]; cyc=float(sys.argv[2])
D=parse_disasm('disasm.txt','ant_env_kernelILi0')
N=parse_ncu('sass_%s.csv'%tag)
assert len(D)==len(N),(len(D),len(N))
regs=subprocess.check_output(['python','regions.py']).decode().split()
R=[]
for spec in regs:
    loc,name=spec.split('='); f,rng=loc.split(':'); lo,hi=map(int,rng.split('-')); R.append((f,lo,hi,name))
from collections import defaultdict
st=defaultdict(lambda:[0,0,0,0])
tot=sum(n['samples'] for n in N)
per=cyc/tot
for i,(a,(f,l),t) in enumerate(D):
    name='other'
    for (rf,lo,hi,rn) in R:
        if f==rf and lo<=l<=hi: name=rn;break
    s=st[name]; s[0]+=1; s[1]+=N[i]['inst']; s[2]+=N[i]['samples']
    if N[i]['inst']>0: s[3]+=1
print("%-20s %7s %8s %10s %8s"%('region','static','exec-static','dyn/warp','cyc/warp'))
for k,v in sorted(st.items(), key=lambda kv:-kv[1][2]):
    print("%-20s %7d %8d %10.0f %8.0f"%(k,v[0],v[3],v[1]/512,v[2]*per))
print('total dyn/warp %.0f'%(sum(v[1] for v in st.values())/512))
