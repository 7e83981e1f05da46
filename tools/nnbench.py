"""Development aid: NN action+gradient kernel timing. args: config(c4|c5|twin20) B reps"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varanneal_b200 import va_nnet
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
st, M = {"c4": ([100] * 5, 1000), "c5": ([25, 30, 4], 10000), "twin20": ([10] * 20, 1000)}[cfg]
st = np.array(st)
NDnet = int(st.sum()); NP = int(sum(st[n] * st[n + 1] + st[n + 1] for n in range(len(st) - 1)))
rng = np.random.RandomState(0)
din, dout = rng.rand(M, st[0]), rng.rand(M, st[-1])
X0 = rng.rand(B, M * NDnet); P0 = 0.3 * rng.randn(B, NP)
an = va_nnet.Annealer(); an.set_structure(st); an.set_activation("sigmoid")
an.set_input_data(din); an.set_output_data(dout)
an.anneal_init(X0, P0, 1.1, [20.0], 1.0, 1e-2, np.arange(NP), init_to_data=False)
an._XP[:, :an._n].copy_(torch.from_numpy(np.concatenate([X0, P0], axis=1)))
for _ in range(3): an._action_grad_native(6.7)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    an._action_grad_native(6.7); ev[i + 1].record()
torch.cuda.synchronize()
ms = np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])
flops = 6.0 * M * sum(st[n] * st[n + 1] for n in range(len(st) - 1)) * B
byt = (16.0 * M * NDnet + 16.0 * NP) * B
print("%s B=%d: %.3f ms -> %.0f evals/s, %.2f TFLOP/s fp64, %.0f GB/s algorithmic" % (
    cfg, B, ms, B / ms * 1e3, flops / ms / 1e9, byt / ms / 1e6))
