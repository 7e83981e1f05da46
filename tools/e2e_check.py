"""Development aid: host->device->host evaluation rate through A_gradA(pinned), C2 shape."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varanneal_b200 import va_ode
D, N, B = 100, 5001, 64
rng = np.random.RandomState(0)
Lidx = [i for i in range(D) if i % 5 in (0, 2)]
Y = rng.randn(N, len(Lidx))
X0 = rng.randn(B, N, D); P0 = np.full((B, 1), 8.17)
an = va_ode.Annealer(); an.set_model("lorenz96", D); an.set_data(Y, t=0.025 * np.arange(N))
an.anneal_init(X0, P0, 2.0, [3], 4.0, 4e-3, Lidx, [0], disc="SimpsonHermite", init_to_data=False)
n = an._n
XP = torch.empty(B, n, dtype=torch.float64, pin_memory=True)
XP[:, :N * D] = torch.from_numpy(X0.reshape(B, -1)); XP[:, N * D] = 8.17
for chunks in (1, 2, 4, 8, 16):
    import functools
    an._pipelined_eval(XP, *an._pinned()[1:], chunks=chunks)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        an._pipelined_eval(XP, *an._pinned()[1:], chunks=chunks)
    dt = (time.perf_counter() - t0) / 5
    print("chunks %2d: %.2f ms per batch -> %.0f evals/s, %.1f GB/s each way" % (chunks, 1e3 * dt, B / dt, B * n * 8 / dt / 1e9))
# raw copies
d = torch.empty(B, n, dtype=torch.float64, device="cuda"); h = torch.empty(B, n, dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(XP, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
with torch.cuda.stream(s1): d.copy_(XP, non_blocking=True)
with torch.cuda.stream(s2): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); t3 = time.perf_counter()
print("raw H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f GB/s each" % (B*n*8/(t1-t0)/1e9, B*n*8/(t2-t1)/1e9, B*n*8/(t3-t2)/1e9))
