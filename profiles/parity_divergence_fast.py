"""Divergence of the fast-policy GPU trajectory from the CPU oracle as a function of step count (config 2,
200,000 envs, fp64) plus per-env contact-event count mismatches.  Output transcribed in r1_summary.md."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import c_oracle as co
from helpers import comp_rel_err
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import stepper, synth, scenes
import rigidbody_simulation_b200.mj as mj
E=200_000
s=synth.sphere_incline(E)
model=mj.MjModel.from_xml_string(scenes.single_body_xml("sphere",[0.2],plane_euler=(0.7,0,0)),nenv=E)
data=mj.MjData(model); data.set_state(s["qpos"],s["qvel"])
model.set_per_env(restitution=s["restitution"],friction=s["friction"])
qp,qv=s["qpos"].copy(),s["qvel"].copy()
cnt=(np.zeros(E,np.uint32),np.zeros(E,np.uint32))
kw=dict(geom="sphere",mass=model.body_mass[-1],inertia=model.body_inertia[-1],size=0.2,plane_pos=[0,0,0],plane_normal=model.plane_normal,gravity=[0,0,-9.8],dt=s["dt"],restitution=s["restitution"],friction=s["friction"],threshold=0.0,counters=cnt)
done=0
for upto in (1,2,5,10,50,100,300):
    co.step_body_plane(qp,qv,upto-done,**kw)
    stepper.step_body_plane(model,data,-1,s["dt"],None,None,0.0,substeps=upto-done,arith="fast")
    done=upto
    gq=data.qpos.torch().cpu().numpy(); gv=data.qvel.torch().cpu().numpy()
    eq=np.abs(gq-qp)/np.maximum(np.abs(qp),1e-3); ev=np.abs(gv-qv)/np.maximum(np.abs(qv),1e-3)
    calls,imps=data.counters()
    bad=(calls[:,0]!=cnt[0])|(imps[:,0]!=cnt[1])
    print(upto,'err q %.2e v %.2e'%(eq.max(),ev.max()),'count mismatches',int(bad.sum()), 'worst env', int(np.argmax(ev.max(axis=1))), 'comp', int(np.argmax(ev.max(axis=0))))
    if bad.any():
        i=np.where(bad)[0][0]; print('  env',i,calls[i,0],cnt[0][i],imps[i,0],cnt[1][i], 'e,mu',s['restitution'][i],s['friction'][i])
