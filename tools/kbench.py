"""Kernel micro-benchmark (development aid): times vab_ode_action_grad on random data of a given
shape through the C ABI.  python tools/kbench.py [disc] [D] [N] [B] [reps]"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varanneal_b200 import va_ode  # noqa: E402

disc = sys.argv[1] if len(sys.argv) > 1 else "SimpsonHermite"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100
N = int(sys.argv[3]) if len(sys.argv) > 3 else 5001
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
rng = np.random.RandomState(0)
Lidx = [i for i in range(D) if i % 5 in (0, 2)]
Y = rng.randn(N, len(Lidx))
X0 = rng.randn(B, N, D)
P0 = np.full((B, 1), 8.17)
an = va_ode.Annealer()
an.set_model("lorenz96", D)
an.set_data(Y, t=0.025 * np.arange(N))
an.anneal_init(X0, P0, 2.0, [3], 4.0, 4e-3, Lidx, [0], disc=disc, init_to_data=False)
an._XP[:, :N * D].copy_(torch.from_numpy(X0.reshape(B, -1)))
an._XP[:, N * D] = 8.17
for _ in range(3):
    an._action_grad_native(8.0)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    an._action_grad_native(8.0)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])
byt = B * 16.0 * N * D + 8.0 * N * len(Lidx)
print("%s D=%d N=%d B=%d: median %.4f ms  min %.4f ms  -> %.0f GB/s (X+G bytes), %.0f evals/s"
      % (disc, D, N, B, np.median(ms), ms.min(), byt / np.median(ms) / 1e6, B / np.median(ms) * 1e3))
