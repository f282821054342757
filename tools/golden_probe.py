import sys, glob, os, numpy as np
sys.path.insert(0,'/root/repo')
from nascargymnasium_b200.engine import Engine
for path in sorted(glob.glob('/root/repo/tests/golden/traj_*.npz')):
    with np.load(path) as z: g={k:z[k] for k in z.files}
    track, C = str(g['track']), int(g['num_cars'])
    eng = Engine(1, C, tracks=[track], discrete=bool(g['discrete']), reset_on_lap=bool(g['reset_on_lap']), auto_reset=False)
    obs0 = eng.reset_host()
    d0 = np.abs(obs0.reshape(C,38) - g['obs0']).max()
    n=len(g['actions']); worst=0; wr=0; first_flag=None; devs=[]
    for t in range(n):
        a=g['actions'][t]
        a = a.astype(np.int32) if g['discrete'] else a.astype(np.float32)
        obs, rew, te, tr, _ = eng.step_host(a)
        d=float(np.abs(obs.reshape(C,38)-g['obs'][t]).max()); worst=max(worst,d); wr=max(wr,float(np.abs(rew-g['reward'][t]).max()))
        devs.append(d)
        if (bool(te[0])!=bool(g['terminated'][t]) or bool(tr[0])!=bool(g['truncated'][t])) and first_flag is None: first_flag=t
        if g['did_reset'][t]: eng.reset_host(fresh=False)
    devs=np.array(devs)
    print(os.path.basename(path), 'C',C,'n',n,'d0 %.1e'%d0,'worst obs %.3g'%worst,'worst rew %.3g'%wr,'first flag mismatch',first_flag, 'dev@100 %.2g @500 %.2g'%(devs[:100].max(), devs[:min(500,n)].max()))
    eng.close()
