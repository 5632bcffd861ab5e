"""Randomised hunt (CPU, the oracle = the reference's arithmetic on the installed SciPy) for an n=3 candidate whose BFGS fallback
(Optimizer.py:151-161) MOVES away from its start point -- the one outcome theta_n3.cuh models only through a status code.
    python tools/hunt_fallback_moved.py <seed> <seconds>
Round 2: 4 seeds x 300 s = 4.3e5 candidates, 3.5e5 of them through the fallback (2.8e5 start point returned, 7.2e4 None): 0 moved."""
import sys, time, itertools, collections
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np, theta_oracle as O
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
cnt=collections.Counter(); t0=time.time(); found=[]
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 600:
    m=int(rng.integers(3,8)); k=int(rng.integers(2,5))
    rN=rng.integers(1000,2_000_000,m)
    style=rng.integers(0,3)
    if style==0:
        r=rng.integers(0,3_000_000,m)            # unrelated reads: badly fitting candidates -> roots out of range
    else:
        C=np.stack([np.full(m,2), rng.integers(0,k+1,m), rng.integers(0,k+1,m)],1)
        mu=rng.dirichlet([1,1,1]); p=(C*rN[:,None]).dot(mu)+1e-9; r=rng.multinomial(int(rN.sum()), p/p.sum())
    for _ in range(40):
        a=rng.integers(0,k+1,m); b=rng.integers(0,k+1,m)
        Cm=np.stack([np.full(m,2),a,b],1)
        try:
            res=O.solve_n3(Cm,[int(x) for x in r],[int(x) for x in rN],full=True)
        except Exception as e:
            cnt['exc']+=1; continue
        if res is None: cnt['none']+=1; continue
        cnt[res[4]]+=1
        if res[4]=='fallback_moved': found.append((Cm.tolist(),[int(x) for x in r],[int(x) for x in rN],res[0].tolist(),float(res[1])))
print(dict(cnt)); print('moved cases:',len(found))
import json; json.dump(found[:20],open('/tmp/moved_%s.json'%(sys.argv[1] if len(sys.argv)>1 else '0'),'w'))
