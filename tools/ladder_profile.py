"""Development aid: C2 ladder through Annealer.anneal() (device-resident) with a phase breakdown.
python tools/ladder_profile.py [B] [nbeta]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from varanneal_b200 import va_ode, _devicemin
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 20
_, Y = bench.twin_data()
X0, P0 = bench.initial_paths(B, 1000)
an = va_ode.Annealer()
an.set_model("lorenz96", bench.D)
an.set_data(Y, t=bench.DT * np.arange(bench.N_MODEL))
import torch
from varanneal_b200 import _lib
lib = _lib.load()
orig = lib.vab_anneal
tt = {}
class W(object):
    def __call__(self, *a):
        t = time.time(); r = orig(*a); tt["native"] = time.time() - t; return r
lib.vab_anneal = W()
oi = an.anneal_init
def timed_init(*a, **k):
    t = time.time(); oi(*a, **k); torch.cuda.synchronize(); tt["init"] = time.time() - t
an.anneal_init = timed_init
t0 = time.time()
an.anneal(X0, P0, bench.ALPHA, np.arange(nb), bench.RM, bench.RF0, bench.LIDX, [0], disc="SimpsonHermite",
          init_to_data=True, opt_args={"gtol": 1e-8, "ftol": 1e-8, "maxfun": 1000000, "maxiter": 1000000})
dt = time.time() - t0
l0 = 0
lib.vab_anneal = orig
print("anneal %.2f s  (init %.2f s, native ladder %.2f s, host layout %.2f s)  launches %d" % (dt, tt.get("init", 0), tt.get("native", 0), dt - tt.get("native", 0) - tt.get("init", 0), an.gpu_launches - l0))
nf = an.nfev_array
print("nfev total %d  per-path ladder totals: min %d mean %.0f max %d" % (nf.sum(), nf.sum(1).min(), nf.sum(1).mean(), nf.sum(1).max()))
print("sum over rungs of per-rung max: %d" % nf.max(0).sum())
for i in range(nb):
    print("beta %2d nfev mean %.0f max %d  A mean %.4e  flags %s" % (i, nf[:, i].mean(), nf[:, i].max(), an.A_array[:, i].mean(), np.bincount(an.exitflags[:, i], minlength=3)))
