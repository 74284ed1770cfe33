import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
import b200bo as bo
rng=np.random.default_rng(4)
X=rng.random((8,4096)); y=np.sin(3*X.sum(0))+0.1*rng.standard_normal(4096)
m=bo.B200GPE(8, mean=bo.MeanConst(0.0), kernel=bo.SEArd(np.zeros(8),0.0), logNoise=-2.0, capacity=4096)
m.fit(X,y)
grid=[(a,l) for a in np.linspace(-3,0,8) for l in np.linspace(-1.5,0.5,8)]
Theta=np.stack([np.concatenate([[a,0.0],np.full(8,l),[0.0]]) for a,l in grid],axis=1)
ref=None
for K in (0,1,2,4,6,8,12):
    m.set_knob("sweep_workers",K)
    m.mll_sweep(Theta[:,:max(K,1)])
    t0=time.perf_counter(); a,b=m.mll_sweep(Theta); t1=time.perf_counter()-t0
    t0=time.perf_counter(); a2,_=m.mll_sweep(Theta,want_grad=False); t2=time.perf_counter()-t0
    if ref is None: ref=(a,b)
    print(f"workers={K}: with gradient {t1:.3f} s, values only {t2:.3f} s, same bits: {np.array_equal(a,ref[0]) and np.array_equal(b,ref[1])}")
