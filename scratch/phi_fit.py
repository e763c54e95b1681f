import numpy as np
from scipy.special import erfc
from scipy.optimize import least_squares
z=np.linspace(0,7.0,40001)
target=0.5*erfc(z/np.sqrt(2))
def model(q,deg,z=z,dt=np.float64):
    p=dt(q[0]); a=[dt(v) for v in q[1:]]
    zz=z.astype(dt)
    u=dt(1)/(dt(1)+p*zz)
    poly=np.zeros_like(zz)
    for k in range(deg,0,-1): poly=(poly+a[k-1])*u
    e=np.exp2((zz*zz*dt(-0.5*1.4426950408889634)).astype(dt)).astype(dt)
    return poly*e
deg=6
p0=np.array([0.3275911/np.sqrt(2)]+[0.5*v for v in [0.254829592,-0.284496736,1.421413741,-1.453152027,1.061405429]]+[0.0])
r=least_squares(lambda q: (model(q,deg)-target)*1e8, p0, method='lm', max_nfev=50000)
q=r.x
print('f64 max err',np.abs(model(q,deg)-target).max())
# crude minimax refinement: iteratively reweighted
w=np.ones_like(z)
for it in range(30):
    r=least_squares(lambda q_: (model(q_,deg)-target)*1e8*w, q, method='lm', max_nfev=20000)
    q=r.x
    e=np.abs(model(q,deg)-target); 
    w*= (1+ 2*e/e.max()); w/=w.mean()
print('f64 minimax-ish err',np.abs(model(q,deg)-target).max())
q32=q.astype(np.float32)
m32=model(q32,deg,dt=np.float32)
print('f32 eval max err',np.abs(m32.astype(np.float64)-target).max())
xs=np.linspace(-7,7,80001)
h=model(q32,deg,z=np.abs(xs),dt=np.float32).astype(np.float64)
phi=np.where(xs>=0,1-h,h)
from scipy.special import erf
gel=xs*phi; ref=0.5*xs*(1+erf(xs/np.sqrt(2)))
print('gelu max abs err',np.abs(gel-ref).max(), 'rel to max(|x|,1)', (np.abs(gel-ref)/np.maximum(np.abs(xs),1)).max())
print('coeffs', ['%.9e'%v for v in q32])
