"""Host emulation of the CUDA strip-walk kernels vs the oracle (no GPU needed).

The emulator (tests/emul/) compiles varanneal_b200/csrc/ode_walk.cuh with g++ and runs every
CTA/thread/phase serially; these tests pin the kernels' index arithmetic, halo exchange, segment
ownership and reductions before any GPU time is spent.  Tolerance: 1e-12 relative (fp64, the
only difference from the oracle is summation order).
"""
import numpy as np
import pytest

from emul_wrap import emul_action_grad
from oracle.ode_port import OdeProblem

TOL = 1e-12
NAKL_P = [120.0, 20.0, 0.3, 50.0, -77.0, -54.0, -40.0, 15.0, 0.1, 0.4, -60.0, -15.0, 1.0, 7.0,
          -55.0, 30.0, 1.0, 5.0]
DISCS = ["euler", "trapezoid", "SimpsonHermite", "forwardmap", "rk4"]


def _check(model, D, Nd, nskip, disc, Lidx, P, Pidx, RM, RF0, B=2, tseg=0, stim=False, seed=0,
           active=None):
    rng = np.random.RandomState(seed)
    N = (Nd - 1) * nskip + 1
    Y = rng.randn(Nd, len(Lidx))
    st = rng.randn(N) if stim else None
    if RM == "arr":
        RM = rng.rand(Nd, len(Lidx)) + 0.5
    if isinstance(RF0, str):
        RF0 = (rng.rand(N - 1, D) + 0.5) * 1e-2
    prob = OdeProblem(model, D, Y, Lidx, 0.01, disc, np.array(P), Pidx, RM, nskip=nskip, stim=st)
    XP = rng.randn(B, prob.n) * 2
    if model == "nakl":
        XP[:, :prob.nX] = 0.2 * rng.rand(B, prob.nX) + 0.4
        XP[:, 0:prob.nX:4] = -70 + 20 * rng.randn(B, N)
        XP[:, prob.nX:] = np.array(P)[Pidx] * (1 + 0.01 * rng.randn(B, len(Pidx)))
    else:
        XP[:, prob.nX:] = np.array(P)[Pidx] + 0.1 * rng.randn(B, len(Pidx))
    scale = 1.5 ** 7
    A, me, fe, G = emul_action_grad(prob, XP, scale, RF0, tseg=tseg, active=active)
    for b in range(B):
        if active is not None and not active[b]:
            assert np.all(np.isnan(G[b, :prob.nX]))      # untouched
            continue
        Ar, mer, fer, gr = prob.action_grad(XP[b], RF0 * scale, parts=True)
        assert abs(A[b] - Ar) <= TOL * abs(Ar)
        assert abs(me[b] - mer) <= TOL * max(abs(mer), 1e-300)
        assert abs(fe[b] - fer) <= TOL * abs(fer)
        assert np.max(np.abs(G[b] - gr)) <= TOL * np.max(np.abs(gr))


@pytest.mark.parametrize("disc", DISCS)
@pytest.mark.parametrize("tseg", [0, 8])
def test_l96_d20(disc, tseg):
    _check("lorenz96", 20, 41, 1, disc, [0, 2, 4, 6, 8, 10, 14, 16], [8.17], [0], 4.0, 4e-3, tseg=tseg)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_arrays_nskip_noparams(disc):
    _check("lorenz96", 10, 21, 2, disc, [1, 3, 9], [8.17], [], "arr", "arr", tseg=6)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_odd_D_scalar_strip(disc):
    _check("lorenz96", 7, 21, 3, disc, [0, 6], [8.17], [0], 2.0, 1e-2, tseg=10)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_d100_batch3(disc):
    _check("lorenz96", 100, 51, 1, disc, list(range(0, 100, 3)), [8.17], [0], 2.0, 1e-2, tseg=10, B=3)


@pytest.mark.parametrize("disc", DISCS)
def test_l96_wide_row_256_threads(disc):
    _check("lorenz96", 600, 13, 1, disc, list(range(0, 600, 7)), [8.17], [0], 2.0, 1e-2, tseg=6, B=2)


@pytest.mark.parametrize("disc", DISCS)
def test_l63(disc):
    _check("lorenz63", 3, 31, 1, disc, [0], [10, 28, 8 / 3], [0, 1, 2], 1.0, 0.3, tseg=4)
    _check("lorenz63", 3, 31, 2, disc, [0, 2], [10, 28, 8 / 3], [1], 1.0, 0.3, B=3)


@pytest.mark.parametrize("disc", DISCS[:4])
def test_nakl_stimulus(disc):
    _check("nakl", 4, 41, 1, disc, [0], NAKL_P, list(range(18)), 1.0,
           np.resize([1e-2, 1e2, 1e2, 1e2], (40, 4)), stim=True, tseg=10)
    _check("nakl", 4, 41, 1, disc, [0], NAKL_P, [0, 3, 17], 1.0, 0.5, stim=True)


def test_active_mask_skips_paths():
    _check("lorenz96", 20, 41, 1, "trapezoid", [0, 2], [8.17], [0], 4.0, 4e-3, B=3,
           active=np.array([1, 0, 1]))


@pytest.mark.parametrize("disc", ["trapezoid", "SimpsonHermite"])
def test_many_segments_many_ctas(disc):
    # 301 rows in segments of 16 -> 19 segments/path, 4 paths -> units spread over several CTAs
    _check("lorenz96", 20, 301, 1, disc, [0, 5, 11], [8.17], [0], 4.0, 4e-3, B=4, tseg=16)
