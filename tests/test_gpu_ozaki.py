"""The tcgen05 / TMEM / TMA path for fp64 contractions (csrc/ozaki_gemm.cu): exactness of the
Ozaki-split int8 GEMM against an fp64 FMA reference, for the one-tile-per-CTA kernel and for the
persistent, warp-specialised, pipelined one, over shapes with ragged tiles (M, N not multiples of
the 128 x 64 / 128 x 32 tiles, K < 128) and up to 30 octaves of dynamic range inside a row.
Bound: (6 K + 2) 2^-49 of the product of the row maxima (~3e-13 at K = 100); asserted 1e-12 of max|C|."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 128, 64, 32, 0.0), (3, 200, 100, 100, 4.0), (16, 1000, 100, 100, 30.0),
                                   (5, 333, 37, 25, 8.0), (2, 129, 65, 128, 12.0)])
def test_ozaki_tcgen05_gemm_is_exact_to_fp64(shape):
    import torch
    from varanneal_b200 import _lib
    P, M, N, K, spread = shape
    ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
    r = ctx.ozaki_gemm_probe(P, M, N, K, reps=2, spread=spread)
    assert r["max_abs_C"] > 0.0
    assert r["max_rel_err"] <= 1e-12, r
    assert r["v2_max_rel_err"] <= 1e-12, r
    assert r["ms_tcgen05"] > 0.0 and r["v2_ms_tcgen05"] > 0.0
    ctx.close()
