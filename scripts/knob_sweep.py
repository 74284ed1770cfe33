"""developer sweep of the K6 chunk / lane knobs on one workload: python scripts/knob_sweep.py [workload]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import b200bo
from b200bo import _lib
import ctypes as C
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "metric"]
X, y, ll = bench.synth(w)
m = b200bo.B200GPE(w["D"], mean=b200bo.MeanConst(0.0), kernel=b200bo.gp._Kernel(w["kernel"], ll, 0.0), logNoise=-2.0, capacity=w["N"])
m.fit(X, y)
par = np.array(bench.acq_params(w, y), float)
Xs = torch.from_numpy(bench.candidates(w, 0).T.copy()).cuda()
dbest = torch.zeros(2, dtype=torch.float64, device="cuda")
dgrad = torch.empty((w["M"], w["D"]), dtype=torch.float64, device="cuda") if w["grad"] else None
pp = par.ctypes.data_as(C.POINTER(C.c_double)) if par.size else None
def step():
    _lib.check(_lib.lib.b200bo_acquire_dev(m._h, _lib.ACQ_KINDS[w["acq"]], pp, par.size, C.c_void_p(Xs.data_ptr()), w["M"], 50, 0, None,
                                           C.c_void_p(dgrad.data_ptr()) if dgrad is not None else None, None, None, C.c_void_p(dbest.data_ptr())), m._h)
for lanes in (2, 1):
    for mb in (32, 64, 96, 128, 192):
        m.set_knob("acq_lanes", lanes); m.set_knob("acq_chunk_mb", mb)
        for _ in range(3): step()
        _lib.check(_lib.lib.b200bo_sync(m._h), m._h)
        t0 = time.perf_counter()
        for _ in range(5): step()
        _lib.check(_lib.lib.b200bo_sync(m._h), m._h)
        dt = (time.perf_counter() - t0) / 5
        print(f"lanes={lanes} chunk_mb={mb}: {dt * 1e3:.3f} ms/step  {w['M'] / dt / 1e6:.2f} M cand/s")
