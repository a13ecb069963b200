import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
n, d, N = 2048, 6, 9472 * 16
X = rng.random((n,d))*9+1; Y = np.sin(X).sum(1); Y=(Y-Y.mean())/Y.std()
Xs=torch.from_numpy(rng.random((N,d))*9+1).cuda(); f=torch.empty(N,dtype=torch.float64,device='cuda'); df=torch.empty(N,d,dtype=torch.float64,device='cuda')
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best=1e9
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
for chunk in (1184, 2368, 3552, 4736, 9472, 14208):
    eng=Engine(0); eng.set_chunk(chunk); eng.fit('rbf',X,Y,[1.0],[[2.0]],[1e-2])
    tg=t(lambda: eng.acq('EI',0.01,Xs,grad=True,out=(f,df),stream=st)); tv=t(lambda: eng.acq('EI',0.01,Xs,out=(f,None),stream=st))
    print(json.dumps({'chunk':chunk,'grad_Mcand_s':N/tg/1e3,'val_Mcand_s':N/tv/1e3}))
    eng.close()
