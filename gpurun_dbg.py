import numpy as np, sys
sys.path.insert(0,'.')
from oracle import rollout_oracle as O
from tests.helpers import *
def growth(skind, ckind, x0, T, integ="euler", fast=False):
    dyn=make_dynamics(skind); dyn.fast_trig=fast; ctl=make_controller(ckind,dyn); osys,octl=oracle_pair(skind,ckind)
    res=dyn.rollout(ctl,x0,T,integrator=integ,record_stride=1)
    xs,us,_,_=O.rollout(osys,octl,x0.astype(np.float64),T,integ,record_stride=1)
    d=np.abs(angle_diff(res.xs,xs,WRAP_IDX[skind]))
    e=d/np.maximum(1,np.abs(xs))
    for t in [1,10,25,40,50,60,75,90,100,200,300,400,500,999]:
        if t<=T:
            pe=e[:t+1].max(axis=(0,2))
            print(skind,ckind,integ,'fast' if fast else 'acc','t',t,'maxerr %.2e'%e[:t+1].max(),'env0 %.2e'%pe[0],'median %.2e q90 %.2e q99 %.2e'%tuple(np.quantile(pe,[.5,.9,.99])))
np.random.seed(0)
x0=make_dynamics("cartpole").get_initial_states(512).astype(np.float32)
growth("cartpole","cartpole_lqr",x0,500)
rng=np.random.default_rng(3); x0=rng.uniform(-0.1,0.1,size=(256,4)).astype(np.float32); x0[0]=[0.001,0,0,0]
growth("acrobot","acrobot_es",x0,90)
growth("acrobot","acrobot_es",x0,90,fast=True)
# per-step control acrobot
dyn=make_dynamics("acrobot"); ctl=make_controller("acrobot_es",dyn); osys,octl=oracle_pair("acrobot","acrobot_es")
x=rand_states("acrobot",4096,7).astype(np.float32)
u=ctl.get_control_efforts(x); uo=octl.control(osys,x.astype(np.float64))
big=O.OracleSystem(osys.kind,4,1,osys.dt,np.array([-1e30]),np.array([1e30]),osys.par)
uu=octl.control(big,x.astype(np.float64))
err=np.abs(u-uo)[:,0]
idx=np.argsort(-err)[:8]
for i in idx: print('acrobot ctl', x[i], 'gpu',u[i],'ref',uo[i],'unclipped',uu[i])
print('frac>1e-5', np.mean(err/np.maximum(1,np.abs(uo[:,0]))>1e-5), 'normwise', (err/np.maximum(1,np.abs(uu[:,0]))).max())
dyn=make_dynamics("cartpole"); ctl=make_controller("cartpole_lqr",dyn); osys,octl=oracle_pair("cartpole","cartpole_lqr")
x=rand_states("cartpole",4096,7).astype(np.float32)
u=ctl.get_control_efforts(x); uo=octl.control(osys,x.astype(np.float64))
err=np.abs(u-uo)[:,0]; idx=np.argsort(-err)[:5]
for i in idx: print('cartpole ctl', x[i], 'gpu',u[i],'ref',uo[i])
print('frac>1e-5', np.mean(err/np.maximum(1,np.abs(uo[:,0]))>1e-5))
