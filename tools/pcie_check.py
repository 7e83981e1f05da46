"""Development aid: PCIe host<->device rates with pinned buffers: each direction alone, both at once,
chunked, and via the pipelined A_gradA seam."""
import os, sys, time
import torch
n = 64 * 500101
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).normal_()
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
GB = n * 8 / 1e9
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def both_chunked(k=8):
    m = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_in[i*m:(i+1)*m].copy_(h_in[i*m:(i+1)*m], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*m:(i+1)*m].copy_(d_out[i*m:(i+1)*m], non_blocking=True)
t = timed(h2d); print("H2D alone   %.2f ms  %.1f GB/s" % (1e3*t, GB/t))
t = timed(d2h); print("D2H alone   %.2f ms  %.1f GB/s" % (1e3*t, GB/t))
t = timed(both); print("both        %.2f ms  %.1f GB/s each way" % (1e3*t, GB/t))
t = timed(both_chunked); print("both x8     %.2f ms  %.1f GB/s each way" % (1e3*t, GB/t))
os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv")
os.system("nvidia-smi topo -m | head -5; nproc; numactl -H 2>/dev/null | head -5")
