import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best=1e9
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
for (n,d,S,N) in ((2048,6,1,9472*8),(128,6,1,2_000_000),(512,4,64,100_000)):
    X = rng.random((n,d))*9+1; Y = np.sin(X).sum(1); Y=(Y-Y.mean())/Y.std()
    eng=Engine(0); eng.fit('rbf',X,Y,np.ones(S),np.full((S,1),2.0),np.full(S,1e-2))
    Xs=torch.from_numpy(rng.random((N,d))*9+1).cuda(); f=torch.empty(N,dtype=torch.float64,device='cuda'); df=torch.empty(N,d,dtype=torch.float64,device='cuda')
    kind='EI_MCMC' if S>1 else 'EI'
    tg=t(lambda: eng.acq(kind,0.01,Xs,grad=True,out=(f,df),stream=st)); tv=t(lambda: eng.acq(kind,0.01,Xs,out=(f,None),stream=st))
    print(json.dumps({'n':n,'d':d,'S':S,'N':N,'grad_ms':tg,'val_ms':tv,'grad_Mcand_s':N/tg/1e3,'val_Mcand_s':N/tv/1e3}))
    eng.close()
# latency
for n in (2048, 4096):
    d=6; X=rng.random((n,d))*9+1; Y=np.sin(X).sum(1); Y=(Y-Y.mean())/Y.std()
    eng=Engine(0); eng.fit('rbf',X,Y,[1.0],[[2.0]],[1e-2])
    x1=rng.random((1,d))*9+1
    import time
    for g in (True, False):
        for _ in range(5): eng.acq('EI',0.01,x1,grad=g)
        t0=time.perf_counter()
        for _ in range(50): eng.acq('EI',0.01,x1,grad=g)
        print(json.dumps({'latency_n':n,'grad':g,'ms':(time.perf_counter()-t0)/50*1e3}))
    eng.close()
