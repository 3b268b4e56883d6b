import re,sys,subprocess
src=open('/root/repo/hrl_pybullet_envs_b200/csrc/hrl_ant.cuh').read().split('\n')
def L(pat):
    return next(i for i,l in enumerate(src) if pat in l)+1
marks=[('fk','LegKin leg_fk('),('x0','constexpr int sym6'),('sinv_mul','void sinv_mul('),('x1','float clampf('),('emit_row','void emit_row('),
       ('pgs_helpers','Blackwell packed fp32'),('x2','void ant_substep('),('contacts_detect','contacts: spheres vs'),('dynamics','smooth dynamics: bias'),
       ('rows_build','constraint rows: counts'),('pgs','// ---------------- projected Gauss-Seidel, Bullet'),('integrate','back to physical velocities')]
pos=[(n,L(p)) for n,p in marks]+[('end',len(src)+1)]
args=[]
for (n,lo),(n2,hi) in zip(pos,pos[1:]):
    if n.startswith('x'): continue
    args.append('hrl_ant.cuh:%d-%d=%s'%(lo,hi-1,n))
args+=['hrl_math.cuh:1-40=math_helpers','hrl_math.cuh:41-100=rng','hrl_sensors.cuh:1-200=sensors']
cu=open('/root/repo/hrl_pybullet_envs_b200/csrc/hrl_b200.cu').read().split('\n')
def C(pat): return next(i for i,l in enumerate(cu) if pat in l)+1
a,b,c,d=C('// ---- load ----'),C('// ---- task layer: observation'),C('// ---- store ----'),C('// PointGather: one thread per env')
args+=['hrl_b200.cu:%d-%d=load_physics_loop'%(a-20,b-1),'hrl_b200.cu:%d-%d=task_layer'%(b,c-1),'hrl_b200.cu:%d-%d=store'%(c,d)]
print(' '.join(args))
