#!/usr/bin/env python
"""Where the fp32 error of this recurrence comes from (CPU, numpy; fixture tests/golden/bench8.npz: 512 spins x 1000
steps with the bench distributions).  Three evaluations of the same step formulas as csrc/bloch_math.cuh:

  f64      everything in double (agrees with the reference's fp64 output to 2e-14)
  f32      everything in float, with CORRECTLY ROUNDED sqrt / sin / cos  (an ideal fp32 kernel)
  angle64  field component bz, |b|^2, phi, sin, cos, 1/phi evaluated in DOUBLE; only the rotation/relaxation
           arithmetic and the stored state are float

Result (max / rms |dM| vs f64):  f32 1.9e-5 / 3.7e-6,  angle64 1.6e-5 / 3.4e-6,  reference fp32 2.0e-5.
=> the error is the random walk of the per-step rounding of the STATE UPDATE (~1e-7 relative per step, x sqrt(1000),
max over 1536 components), not the trigonometry or the angle: no fp32 kernel gets below ~1.5e-5 here, which is why
the fp32 parity tests assert  <= max(1e-5, the reference's own fp32 error on the same inputs).
"""
import os
import numpy as np
g=dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'bench8.npz')))
print({k:v.shape for k,v in g.items() if k.startswith('in_')})
f32=np.float32; f64=np.float64
rf=g['in_rf'].astype(f64); gr=g['in_gr'].astype(f64); loc=g['in_loc'].astype(f64); df=g['in_df'].astype(f64); b1=g['in_b1'].astype(f64)
M0=g['in_M0'].astype(f64); T1=float(np.ravel(g['in_T1'])[0]); T2=float(np.ravel(g['in_T2'])[0]); gam=float(np.ravel(g['in_gam'])[0]); dt=float(np.ravel(g['in_dt'])[0])
N,nM,_=loc.shape; nT=rf.shape[2]
if b1.ndim==4: b1=b1[...,0]
if rf.ndim==4: rf=rf[...,0]
G=2*np.pi*gam*dt
def run(mode):
    # mode: 'f64' all double; 'f32' all float; 'angle64': field bz + p2 + phi + reduction in f64, rest f32; 'state64': only state update in f64
    T=f64 if mode=='f64' else f32
    cbr=(G*b1[0,:,0]).astype(T); cbi=(G*b1[0,:,1]).astype(T)
    gl=(G*loc[0]).astype(T); gbz0=(2*np.pi*dt*df[0]).astype(T)
    e1=T(np.expm1(-dt/T1)); e2=T(np.expm1(-dt/T2))
    m=M0[0].astype(T).copy()
    for t in range(nT):
        rx,ry=T(rf[0,0,t]),T(rf[0,1,t]); gx,gy,gz=(T(gr[0,k,t]) for k in range(3))
        bx=cbr*rx-cbi*ry; by=cbr*ry+cbi*rx
        if mode in('angle64',):
            bz64=gl[:,0].astype(f64)*f64(gx)+gl[:,1].astype(f64)*f64(gy)+gl[:,2].astype(f64)*f64(gz)+gbz0.astype(f64)
            p2=bx.astype(f64)**2+by.astype(f64)**2+bz64**2
            phi=np.sqrt(np.maximum(p2,1e-24)); s=np.sin(phi); c=np.cos(phi)
            rs=(1/phi)
            a=(s*rs).astype(T); d=((1-c)*rs*rs).astype(T); c=c.astype(T); bz=bz64.astype(T)
        else:
            bz=gl[:,0]*gx+gl[:,1]*gy+gl[:,2]*gz+gbz0
            p2=np.maximum(bx*bx+by*by+bz*bz,T(1e-24))
            if mode=='f32':
                phi=np.sqrt(p2.astype(f64)).astype(T)  # correctly rounded sqrt
                s=np.sin(phi.astype(f64)).astype(T); c=np.cos(phi.astype(f64)).astype(T)  # correctly rounded trig of fp32 phi
                rs=(T(1)/phi)
            else:
                phi=np.sqrt(p2); s=np.sin(phi); c=np.cos(phi); rs=1/phi
            a=s*rs; d=(T(1)-c)*rs*rs
        kk=d*(bx*m[:,0]+by*m[:,1]+bz*m[:,2])
        abx,aby,abz=a*bx,a*by,a*bz
        nx=abz*m[:,1]-aby*m[:,2]+kk*bx+c*m[:,0]
        ny=abx*m[:,2]-abz*m[:,0]+kk*by+c*m[:,1]
        nz=aby*m[:,0]-abx*m[:,1]+kk*bz+c*m[:,2]
        nx=nx+e2*nx; ny=ny+e2*ny; nz=nz+e1*(nz-T(1))
        m=np.stack([nx,ny,nz],1).astype(T)
    return m.astype(f64)
ref=run('f64')
print('f64 vs golden', np.abs(ref-g['Mo_f64'][0]).max())
for mode in ('f32','angle64'):
    r=run(mode); print(mode, np.abs(r-ref).max(), np.sqrt(((r-ref)**2).mean()))
print('reference fp32', np.abs(g['Mo_f32'][0]-g['Mo_f64'][0]).max())
