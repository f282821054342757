import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.vector_env import NascarVectorEnv
from nascargymnasium_b200 import layout as L, track as T
R=L.R
E=2048
v = NascarVectorEnv(E, track_file=None, discrete_action_space=True)
obs,_ = v.reset(seed=123)
t0=v.track_id.copy()
coast=np.zeros(E,dtype=np.int64)
for step in range(1,601):
    obs,rew,te,tr,info=v.step(coast)
t1=v.track_id.copy()
rng=np.random.default_rng(0)
acts=rng.integers(0,5,size=(40,E))
for k in range(8):
    obs,rew,te,tr,info=v.step(acts[k])
    d=np.flatnonzero(te|tr)
    recs=v.engine.get_state_host()
    u=recs.view(np.uint32)
    print("k",k,"done",len(d), d[:10], "tracks changed", int((v.track_id!=t1).sum()))
    for e in d[:3]:
        print("   env",e,"track",u[e,R["NCG_R_TRACK"]],"step",u[e,R["NCG_R_STEP"]],"flags",hex(u[e,R["NCG_R_FLAGS"]]), "cumrew", recs[e,R["NCG_R_CUM_REWARD"]], "len", info["episode"]["l"][e] if info else None, "ret", info["episode"]["r"][e] if info else None)
    zero=np.flatnonzero((u[:,R["NCG_R_STEP"]]!=k+1))
    print("   envs with unexpected step count", len(zero), zero[:10], u[zero[:10],R["NCG_R_STEP"]])
