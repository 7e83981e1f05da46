"""Runs the tcgen05 Ozaki-split GEMM prototype (csrc/ozaki_gemm.cu) at the C4 contraction shape and
prints accuracy + timing as JSON lines:  python tools/ozaki_probe.py > gpurun_out/ozaki.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from varanneal_b200 import _lib                # noqa: E402

ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
print(json.dumps({"fp64_fma_peak_tflops": ctx.fp64_peak_tflops()}), flush=True)
SHAPES = ((1, 128, 64, 32, 0.0), (2, 200, 100, 100, 4.0), (256, 1000, 100, 100, 0.0),
          (256, 1000, 100, 100, 8.0), (256, 1000, 100, 100, 30.0), (64, 10000, 30, 25, 8.0))
if os.environ.get("VAB_OZAKI_SHAPES") == "c4":
    SHAPES = SHAPES[3:4]
for P, M, N, K, spread in SHAPES:
    r = ctx.ozaki_gemm_probe(P, M, N, K, reps=5, spread=spread)
    r.update({"P": P, "M": M, "N": N, "K": K, "spread_octaves": spread})
    print(json.dumps(r), flush=True)
