import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
from nascargymnasium_b200.engine import Engine
from nascargymnasium_b200 import track as T
E=4096
names=list(T.BUILTIN_TRACK_NAMES)
for mode in (1,0):
    eng=Engine(E,1,tracks=names,auto_reset=True)
    eng.reset_host(track_id=(np.arange(E)*len(names)//E).astype(np.int32))
    last=torch.empty((E,38),device='cuda:0')
    t0=time.time()
    for i in range(60):
        eng.rollout(1000, seed=3+i, mode=mode, obs_last=last.view(-1))
    torch.cuda.synchronize()
    st=eng.read_stats(); o=last.cpu().numpy(); recs=eng.get_state_host()
    print('mode',mode,'time',round(time.time()-t0,2),st,'finite',np.isfinite(o).all(),np.isfinite(recs[:,:7]).all(),'obs range',o.min(),o.max())
    eng.close()
