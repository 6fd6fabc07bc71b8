import numpy as np, torch
from oracle import pyoracle as po
from drone_b200.vec import SwarmVec
n,A,R,seed=64,16,5,31
orc=po.OrcSwarm(n,A,R,seed=seed); orc.reset(seed, mode=po.RESET_PHILOX)
vec=SwarmVec(n,A,R,math="strict",seed=seed); vec.reset(seed)
o=vec.observations.cpu().numpy(); r=orc.observations
bad=np.argwhere(o.view(np.uint32)!=r.view(np.uint32))
print('mismatches', len(bad), 'cols', np.unique(bad[:,1]), 'rows', np.unique(bad[:,0])[:20])
env,ag=vec.split_state(vec.get_state()); oenv,oag=orc.get_state()
print('env equal', np.array_equal(env.view(np.uint32),oenv.view(np.uint32)), 'tasks dev', env[:10,1], 'tasks orc', oenv[:10,1])
b2=np.argwhere(ag.view(np.uint32)!=oag.view(np.uint32)); print('agent blob mismatches', len(b2), 'fields', np.unique(b2[:,2]) if len(b2) else None)
if len(b2): 
    e,a,f=b2[0]; print(e,a,f, ag[e,a,f], oag[e,a,f], 'task', oenv[e,1])
be=np.argwhere(env.view(np.uint32)!=oenv.view(np.uint32)); print('env mismatch', be[:10])
