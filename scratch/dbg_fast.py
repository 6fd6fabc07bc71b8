import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import pyoracle as po
from drone_b200 import capi
from drone_b200.vec import RaceVec
n,T,seed=1024,300,11
cpu=po.RefRace(n)
rng=np.random.default_rng(1234); tape=rng.uniform(-1.3,1.3,size=(16,n,4)).astype(np.float32)
cpu.reset(seed)
vec=RaceVec(n,math='fast'); vec.set_reset_mode(capi.RESET_INJECT)
dt=torch.from_numpy(tape).cuda()
names=['px','py','pz','vx','vy','vz','qw','qx','qy','qz','wx','wy','wz','r0','r1','r2','r3']
worst=np.zeros(17); worstrel=np.zeros(17); wobs=np.zeros(29); wobsrel=np.zeros(29)
for t in range(T):
    vec.put_state(cpu.get_state())
    cpu.step(tape[t%16])
    idx=np.flatnonzero(cpu.terminals); pl=np.zeros((n,cpu.blob),np.float32)
    if len(idx): pl[idx]=cpu.get_state(idx)
    vec.set_reset_payload(pl); vec.step(dt[t%16])
    st=vec.get_state(); rs=cpu.get_state(); keep=cpu.terminals==0
    d=np.abs(st[keep,:17]-rs[keep,:17]); rel=d/(np.abs(rs[keep,:17])+1e-30)
    # error beyond abs floor
    worst=np.maximum(worst,d.max(0)); 
    bad = d>1e-6+1e-5*np.abs(rs[keep,:17])
    if bad.any() and t<40:
        r,c=np.argwhere(bad)[0]; print('t',t,names[c],'got',st[keep][r,c],'ref',rs[keep][r,c],'vec', rs[keep][r,10:13] if c>=10 and c<13 else rs[keep][r,3:6])
    ob=vec.observations.cpu().numpy(); do=np.abs(ob-cpu.observations); wobs=np.maximum(wobs,do.max(0))
print('max abs state err', dict(zip(names,worst)))
print('max abs obs err', wobs)
