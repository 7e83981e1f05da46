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
