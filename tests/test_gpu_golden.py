"""GPU parity against the committed golden vectors produced by the reference's own code
(tests/golden/make_golden.py): action value, measurement / model error split, and gradient
(complex-step through the reference) for every ODE case -- shipped Lorenz96 data under all four
reference discretisations, merr_nskip = 2 with per-entry RM and per-component RF0, no estimated
parameters, and the NaKL neuron with an injected-current stimulus.  Tolerance 1e-10 relative
(BASELINE.json north_star)."""
import numpy as np
import pytest

import golden_util

pytestmark = pytest.mark.gpu
TOL = 1e-10
ODE_CASES = golden_util.ode_cases()


@pytest.mark.parametrize("c", ODE_CASES, ids=[c["name"] for c in ODE_CASES])
def test_ode_action_grad_vs_reference_golden(c):
    from varanneal_b200 import va_ode
    an = va_ode.Annealer()
    an.set_model(c["model"], c["X0"].shape[1])
    an.set_data(c["Y"], stim=c["stim"], t=c["t"])
    X0 = c["X0"].copy()
    RM = c["RM"]
    RF0 = c["RF0"] if np.isscalar(c["RF0"]) else list(c["RF0"])
    an.anneal_init(X0, c["P0"].copy(), c["alpha"], [c["beta"]], RM, RF0, c["Lidx"], c["Pidx"],
                   dt_model=c["dt_model"], init_to_data=False, disc=c["disc"])
    XP = np.append(c["X0"].ravel(), c["P0"][c["Pidx"]]) if len(c["Pidx"]) else c["X0"].ravel()
    A, g = an.A_gradA_taped(XP)
    assert abs(A - c["A"][0]) <= TOL * abs(c["A"][0])
    assert np.max(np.abs(g - c["grad"])) <= TOL * np.max(np.abs(c["grad"]))
    assert abs(an.me_gaussian(XP[:c["X0"].size]) - c["A"][1]) <= TOL * abs(c["A"][1])
    assert abs(an.fe_gaussian(XP) - c["A"][2]) <= TOL * abs(c["A"][2])
    assert abs(an.A_gaussian(XP) - c["A"][0]) <= TOL * abs(c["A"][0])


RM_CASES = golden_util.rm_matrix_cases()


@pytest.mark.parametrize("c", RM_CASES, ids=[c["name"] for c in RM_CASES])
def test_matrix_rm_vs_reference_golden(c):
    """RM as an (L, L) or (N_data, L, L) matrix (va_ode.py:149-152, 616-617), not symmetric."""
    from varanneal_b200 import va_ode
    an = va_ode.Annealer()
    an.set_model(c["model"], c["X0"].shape[1])
    an.set_data(c["Y"], t=c["t"])
    an.anneal_init(c["X0"].copy(), c["P0"].copy(), c["alpha"], [c["beta"]], c["RM"].copy(), c["RF0"], c["Lidx"],
                   c["Pidx"], dt_model=c["dt_model"], init_to_data=False, disc=c["disc"])
    XP = np.append(c["X0"].ravel(), c["P0"][c["Pidx"]])
    A, g = an.A_gradA_taped(XP)
    assert abs(A - c["A"][0]) <= TOL * abs(c["A"][0])
    assert np.max(np.abs(g - c["grad"])) <= TOL * np.max(np.abs(c["grad"]))
    assert abs(an.me_gaussian(XP[:c["X0"].size]) - c["A"][1]) <= TOL * abs(c["A"][1])
    assert abs(an.fe_gaussian(XP) - c["A"][2]) <= 1e-9 * abs(c["A"][2])     # fe = A - me, six digits smaller than me
    # a batch through the minimiser: every path ends at a stationary point of the oracle action
    import scipy.optimize as opt
    from oracle.ode_port import OdeProblem
    B = 3
    rng = np.random.default_rng(5)
    X0 = c["X0"][None] + 0.2 * rng.standard_normal((B,) + c["X0"].shape)
    beta = [c["beta"], c["beta"] + 2]
    an = va_ode.Annealer()
    an.set_model(c["model"], c["X0"].shape[1])
    an.set_data(c["Y"], t=c["t"])
    an.anneal(X0, np.tile(c["P0"], (B, 1)), c["alpha"], beta, c["RM"].copy(), c["RF0"], c["Lidx"], c["Pidx"],
              dt_model=c["dt_model"], init_to_data=True, disc=c["disc"], opt_args={"gtol": 1e-9, "ftol": 1e-13})
    nd, L = c["Y"].shape[0], len(c["Lidx"])
    RMf = c["RM"] if c["RM"].ndim == 3 else np.resize(c["RM"], (nd, L, L))
    prob = OdeProblem(c["model"], c["X0"].shape[1], c["Y"], c["Lidx"], an.dt_model, c["disc"], c["P0"], c["Pidx"], RMf,
                      nskip=an.merr_nskip)
    rf = c["RF0"] * c["alpha"] ** beta[-1]
    for b in range(B):
        xp = an._est_slice(an.minpaths[b, -1][None])[0]
        A0, g0 = prob.action_grad(xp, rf)
        assert abs(A0 - an.A_array[b, -1]) <= TOL * abs(A0)
        r = opt.minimize(lambda v: prob.action_grad(v, rf), xp, jac=True, method="L-BFGS-B",
                         options=dict(gtol=1e-9, ftol=1e-13, maxiter=200))
        # (flat valleys: what SciPy still gains from the device's stopping point is bounded, not zero)
        assert A0 - r.fun <= 1e-4 * abs(A0), (b, A0, r.fun, r.nit, float(np.max(np.abs(g0))), an.exitflags[b], an.nit_array[b])


RF_CASES = golden_util.rf_matrix_cases()


@pytest.mark.parametrize("c", RF_CASES, ids=[c["name"] for c in RF_CASES])
def test_matrix_rf_vs_reference_golden(c):
    """RF0 as a (D, D) or (N-1, D, D) matrix with SimpsonHermite (va_ode.py:211-218, 629-636)."""
    import scipy.optimize as opt
    from oracle.ode_port import OdeProblem
    from varanneal_b200 import va_ode
    N, D = c["X0"].shape

    def annealer():
        an = va_ode.Annealer()
        an.set_model(c["model"], D)
        an.set_data(c["Y"], t=c["t"])
        return an
    an = annealer()
    an.anneal_init(c["X0"].copy(), c["P0"].copy(), c["alpha"], [c["beta"]], c["RM"], c["RF0"].copy(), c["Lidx"],
                   c["Pidx"], init_to_data=False, disc="SimpsonHermite")
    XP = np.append(c["X0"].ravel(), c["P0"][c["Pidx"]])
    A, g = an.A_gradA_taped(XP)
    assert abs(A - c["A"][0]) <= TOL * abs(c["A"][0])
    assert np.max(np.abs(g - c["grad"])) <= TOL * np.max(np.abs(c["grad"]))
    assert abs(an.fe_gaussian(XP) - c["A"][2]) <= TOL * abs(c["A"][2])
    assert abs(an.me_gaussian(XP[:N * D]) - c["A"][1]) <= TOL * abs(c["A"][1])
    with pytest.raises(ValueError, match="SimpsonHermite"):
        annealer().anneal_init(c["X0"].copy(), c["P0"].copy(), c["alpha"], [c["beta"]], c["RM"], c["RF0"].copy(),
                               c["Lidx"], c["Pidx"], init_to_data=False, disc="trapezoid")
    # a two-rung ladder of a small batch ends at stationary points of the oracle action
    B = 2
    rng = np.random.default_rng(9)
    X0 = c["X0"][None] + 0.2 * rng.standard_normal((B, N, D))
    beta = [c["beta"], c["beta"] + 2]
    an = annealer()
    an.anneal(X0, np.tile(c["P0"], (B, 1)), c["alpha"], beta, c["RM"], c["RF0"].copy(), c["Lidx"], c["Pidx"],
              init_to_data=True, disc="SimpsonHermite", opt_args={"gtol": 1e-9, "ftol": 1e-13})
    prob = OdeProblem(c["model"], D, c["Y"], c["Lidx"], an.dt_model, "SimpsonHermite", c["P0"], c["Pidx"], c["RM"])
    RF = (c["RF0"] if c["RF0"].ndim == 3 else np.resize(c["RF0"], (N - 1, D, D))) * c["alpha"] ** beta[-1]
    for b in range(B):
        xp = an._est_slice(an.minpaths[b, -1][None])[0]
        A0, g0 = prob.action_grad(xp, RF)
        assert abs(A0 - an.A_array[b, -1]) <= TOL * abs(A0)
        r = opt.minimize(lambda v: prob.action_grad(v, RF), xp, jac=True, method="L-BFGS-B",
                         options=dict(gtol=1e-9, ftol=1e-13, maxiter=200))
        assert A0 - r.fun <= 1e-4 * abs(A0), (b, A0, r.fun, r.nit)


def test_kernel_families_agree_bitwise_contract():
    """The TMA stream kernels and the register sweep kernels implement the same arithmetic with
    different data movement; on the shipped Lorenz96 case they must agree to rounding."""
    import os
    import subprocess
    import sys
    code = ("import numpy as np, golden_util\n"
            "from varanneal_b200 import va_ode\n"
            "c=[c for c in golden_util.ode_cases() if c['name']=='l96_shipped_SimpsonHermite_b30'][0]\n"
            "an=va_ode.Annealer(); an.set_model('lorenz96',20); an.set_data(c['Y'],t=c['t'])\n"
            "an.anneal_init(c['X0'].copy(),c['P0'].copy(),c['alpha'],[c['beta']],c['RM'],c['RF0'],c['Lidx'],c['Pidx'],init_to_data=False,disc=c['disc'])\n"
            "A,g=an.A_gradA(np.append(c['X0'].ravel(),c['P0']))\n"
            "print(repr(A)); print(repr(float(np.abs(g).sum())))\n")
    outs = []
    for kern in ("stream", "sweep"):
        env = dict(os.environ, VAB_KERNEL=kern, PYTHONPATH=os.pathsep.join(sys.path))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append([float(v) for v in r.stdout.split()])
    assert abs(outs[0][0] - outs[1][0]) <= 1e-13 * abs(outs[0][0])
    assert abs(outs[0][1] - outs[1][1]) <= 1e-12 * abs(outs[0][1])
