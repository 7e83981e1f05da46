"""Generates the golden fixtures in this directory from the *reference itself*.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

The reference's own NumPy action code (varanneal/va_ode.py, va_nnet.py) is loaded verbatim through
oracle.ref_shim (py2->py3 text substitutions + a stub adolc module, nothing else) and evaluated on
seeded inputs:
  * action values  = reference A_gaussian / me_gaussian / fe_gaussian,
  * gradients      = complex-step differentiation through that same reference code (equal to
                     ADOL-C's reverse mode up to rounding),
  * ladders        = the reference's anneal()/anneal_step() driving SciPy L-BFGS-B, with the
                     gradient supplied by the NumPy adjoint (oracle.ode_port), which the first two
                     items pin.
The inputs are stored next to the outputs so the GPU box (which has no /root/reference) can
replay them.  The only reference *data* copied is the shipped Lorenz96 observation file
(examples/Lorenz96_D20/l96_D20_dt0p025_N161_sm0p5_sec1_mem1.npy, 27 KB) and the first rows of the
tutorial's NaKL data, as test inputs.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle.models_np import MODELS              # noqa: E402
from oracle.ode_port import OdeProblem           # noqa: E402
from oracle import nnet_port                     # noqa: E402

REF = ref_shim.REFERENCE_ROOT
L96_FILE = os.path.join(REF, "examples", "Lorenz96_D20", "l96_D20_dt0p025_N161_sm0p5_sec1_mem1.npy")
NAKL_DIR = os.path.join(REF, "examples", "jupyter-tutorial", "NaKL", "data")

ShimOde, ShimNnet = ref_shim.make_shim_classes()


def ref_annealer(model, D, data_t, Y, stim, X0, P0, alpha, beta, RM, RF0, Lidx, Pidx, dt_model, disc):
    an = ShimOde()
    an.set_model(MODELS[model], D)
    an.set_data(Y, stim=stim, t=data_t)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        an.anneal_init(X0, P0, alpha, beta, RM, RF0, np.array(Lidx), Pidx, dt_model=dt_model,
                       init_to_data=False, disc=disc)
    return an


def ode_cases():
    """(name, kwargs) list; every case is evaluated by the reference."""
    data = np.load(L96_FILE)
    t, Yall = data[:, 0], data[:, 1:]
    cases = []
    Lidx8 = [0, 2, 4, 6, 8, 10, 14, 16]
    for disc in ("euler", "trapezoid", "SimpsonHermite", "forwardmap"):
        for beta in (0, 30):
            cases.append(dict(name="l96_shipped_%s_b%d" % (disc, beta), model="lorenz96", D=20, t=t,
                              Y=Yall[:, Lidx8], stim=None, Lidx=Lidx8, Pidx=[0], P=[8.17], seed=1,
                              alpha=1.5, beta=beta, RM=4.0, RF0=4e-6, dt_model=None, disc=disc, N=161))
    # README variant: 7 observed components, no parameter estimated
    Lidx7 = [0, 2, 4, 8, 10, 14, 16]
    cases.append(dict(name="l96_7obs_noparam_trapezoid", model="lorenz96", D=20, t=t[:41],
                      Y=Yall[:41, Lidx7], stim=None, Lidx=Lidx7, Pidx=[], P=[8.17], seed=2, alpha=1.5,
                      beta=12, RM=4.0, RF0=4e-6, dt_model=None, disc="trapezoid", N=41))
    # model grid twice as fine as the data (merr_nskip = 2), RM per (n, l), RF0 per component
    rng = np.random.RandomState(77)
    for disc in ("trapezoid", "SimpsonHermite", "euler"):
        cases.append(dict(name="l96_nskip2_arrays_%s" % disc, model="lorenz96", D=20, t=t[:21],
                          Y=Yall[:21, Lidx8], stim=None, Lidx=Lidx8, Pidx=[0], P=[7.5], seed=3,
                          alpha=1.5, beta=20, RM=rng.rand(21, 8) + 0.5, RF0=list(1e-5 * (rng.rand(20) + 0.5)),
                          dt_model=0.0125, disc=disc, N=41))
    # NaKL neuron with injected current (tutorial data), vector RF0 as in the tutorial
    V = np.load(os.path.join(NAKL_DIR, "NaKL_Vdata_dt0p02_N6001_sm1p0.npy"))[:101]
    stim = np.load(os.path.join(NAKL_DIR, "NaKL_stim_dt0p02_N6001.npy"))[:101]
    ptrue = np.load(os.path.join(NAKL_DIR, "NaKL_trueparam_dt0p02_N6001.npy"))
    for disc, Pidx in (("trapezoid", list(range(18))), ("SimpsonHermite", [0, 3, 17]), ("euler", list(range(18)))):
        cases.append(dict(name="nakl_%s_%dp" % (disc, len(Pidx)), model="nakl", D=4, t=V[:, 0], Y=V[:, 1:2],
                          stim=stim[:, 1], Lidx=[0], Pidx=Pidx, P=list(ptrue * 1.02), seed=4, alpha=1.1,
                          beta=60, RM=1.0, RF0=[1e-8, 1e-4, 1e-4, 1e-4], dt_model=None, disc=disc, N=101))
    return cases


def make_ode():
    out = {}
    names = []
    for c in ode_cases():
        rng = np.random.RandomState(c["seed"])
        N, D = c["N"], c["D"]
        if c["model"] == "nakl":
            X0 = np.column_stack([-70 + 20 * rng.randn(N), 0.2 * rng.rand(N) + 0.4,
                                  0.2 * rng.rand(N) + 0.4, 0.2 * rng.rand(N) + 0.4])
        else:
            X0 = 20.0 * rng.rand(N, D) - 10.0
        P0 = np.array(c["P"], dtype=np.float64)
        an = ref_annealer(c["model"], D, c["t"], c["Y"], c["stim"], X0.copy(), P0.copy(), c["alpha"],
                          [c["beta"]], c["RM"], c["RF0"], c["Lidx"], c["Pidx"], c["dt_model"], c["disc"])
        XP = np.append(X0.ravel(), P0[c["Pidx"]]) if len(c["Pidx"]) else X0.ravel().copy()
        A = float(an.A(XP))
        me = float(an.me_gaussian(XP[:N * D]))
        fe = float(an.fe_gaussian(XP))
        t0 = time.time()
        g = ref_shim.complex_step_grad(an.A, XP)
        # pin the NumPy port on the way
        prob = OdeProblem(c["model"], D, c["Y"], c["Lidx"], an.dt_model, c["disc"], P0, c["Pidx"],
                          c["RM"] if np.isscalar(c["RM"]) else np.asarray(c["RM"]), nskip=an.merr_nskip,
                          stim=c["stim"])
        rf = an.RF if np.isscalar(an.RF) else np.asarray(an.RF)
        Ap, gp = prob.action_grad(XP, rf)
        print("%-34s A=%.16e port rel %.1e grad rel %.1e (%.1fs)" % (
            c["name"], A, abs(Ap - A) / abs(A), np.max(np.abs(gp - g)) / np.max(np.abs(g)), time.time() - t0))
        n = c["name"]
        names.append(n)
        out[n + "/X0"] = X0
        out[n + "/P0"] = P0
        out[n + "/t"] = np.asarray(c["t"])
        out[n + "/Y"] = np.asarray(c["Y"])
        out[n + "/stim"] = np.zeros(0) if c["stim"] is None else np.asarray(c["stim"])
        out[n + "/Lidx"] = np.asarray(c["Lidx"], dtype=np.int64)
        out[n + "/Pidx"] = np.asarray(c["Pidx"], dtype=np.int64)
        out[n + "/RM"] = np.asarray(c["RM"], dtype=np.float64)
        out[n + "/RF0"] = np.asarray(c["RF0"], dtype=np.float64)
        out[n + "/meta"] = np.array([c["alpha"], c["beta"], -1.0 if c["dt_model"] is None else c["dt_model"]])
        out[n + "/model_disc"] = np.array([c["model"], c["disc"]])
        out[n + "/A"] = np.array([A, me, fe])
        out[n + "/grad"] = g
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ode_action_golden.npz"), **out)


def make_nnet():
    out = {}
    names = []
    for name, structure, M, seed, beta in (("bar_25_30_4_M10", [25, 30, 4], 10, 43759436, 0),
                                           ("twin_6x10_M3", [10] * 6, 3, 89072545, 40),
                                           ("mixed_7_12_5_9_M4", [7, 12, 5, 9], 4, 11, 25)):
        rng = np.random.RandomState(seed)
        structure = np.array(structure)
        NDnet = int(structure.sum())
        data_in = rng.rand(M, structure[0])
        data_out = rng.rand(M, structure[-1])
        NP = int(sum(structure[n] * structure[n + 1] + structure[n + 1] for n in range(len(structure) - 1)))
        X0 = rng.rand(M * NDnet)
        P0 = 0.4 * rng.randn(NP)
        Pidx = np.arange(NP) if name != "mixed_7_12_5_9_M4" else np.arange(0, NP, 2)
        RM, alpha = 3.0, 1.1
        RF0 = 1e-3
        an = ShimNnet()
        an.set_structure(structure)
        an.set_activation(nnet_port.sigmoid)
        an.set_input_data(data_in)
        an.set_output_data(data_out)
        Lidx = [np.arange(structure[0]), np.arange(structure[-1])]
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            an.anneal_init(X0.copy(), P0.copy(), alpha, [beta], RM, RF0, Pidx, Lidx=Lidx, init_to_data=False)
        XP = np.append(X0, P0[Pidx])
        A = float(an.A(XP))
        me = float(an.me_gaussian(XP[:M * NDnet]))
        fe = float(an.fe_gaussian(XP))
        g = ref_shim.complex_step_grad(an.A, XP)
        prob = nnet_port.NnetProblem(structure, data_in, data_out, Lidx, P0, Pidx, RM)
        Ap, gp = prob.action_grad(XP, RF0 * alpha ** beta)
        print("%-34s A=%.16e port rel %.1e grad rel %.1e" % (
            name, A, abs(Ap - A) / abs(A), np.max(np.abs(gp - g)) / np.max(np.abs(g))))
        names.append(name)
        for k, v in (("structure", structure), ("data_in", data_in), ("data_out", data_out), ("X0", X0),
                     ("P0", P0), ("Pidx", Pidx), ("meta", np.array([RM, RF0, alpha, beta])),
                     ("A", np.array([A, me, fe])), ("grad", g)):
            out[name + "/" + k] = np.asarray(v)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "nnet_action_golden.npz"), **out)


def make_ladder():
    """Reference anneal() on the shipped Lorenz96 example (config 1), both the signature-default
    trapezoid rule and the script's SimpsonHermite, with tight optimiser tolerances so that every
    rung is a converged local minimum rather than an early stop."""
    data = np.load(L96_FILE)
    t, Yall = data[:, 0], data[:, 1:]
    Lidx = [0, 2, 4, 6, 8, 10, 14, 16]
    out = {"data": data}
    for disc in ("trapezoid", "SimpsonHermite"):
        np.random.seed(12345)
        X0 = (20.0 * np.random.rand(161 * 20) - 10.0).reshape((161, 20))
        P0 = np.array([4.0 * np.random.rand() + 6.0])
        beta_array = np.arange(0, 60, 3)
        alpha, RM, RF0 = 1.5, 4.0, 4e-6
        opts = {"gtol": 1e-11, "ftol": 1e-15, "maxfun": 1000000, "maxiter": 1000000}
        an = ShimOde()
        an.set_model(MODELS["lorenz96"], 20)
        an.set_data(Yall[:, Lidx], t=t)
        prob = OdeProblem("lorenz96", 20, Yall[:, Lidx], Lidx, 0.025, disc, P0, [0], RM)
        an.grad_fn = lambda XP, an=an, prob=prob: prob.action_grad(XP, an.RF)[1]
        t0 = time.time()
        X0c = X0.copy()
        an.anneal_quiet(X0c, P0.copy(), alpha, beta_array, RM, RF0, np.array(Lidx), [0], dt_model=0.025,
                        init_to_data=True, disc=disc, method="L-BFGS-B", opt_args=opts)
        table = np.column_stack([an.beta_array, an.A_array, an.me_array, an.fe_array,
                                 an.fe_array / (RF0 * alpha ** an.beta_array.astype(float))])
        print("ladder %s: %.1f s; A first/last %.6e %.6e; k last %.6f" % (
            disc, time.time() - t0, table[0, 1], table[-1, 1], an.minpaths[-1, -1]))
        out[disc + "/X0"] = X0
        out[disc + "/P0"] = P0
        out[disc + "/beta"] = beta_array
        out[disc + "/table"] = table
        out[disc + "/params"] = an.minpaths[:, -1]
        out[disc + "/lastpath"] = an.minpaths[-1]
        out[disc + "/meta"] = np.array([alpha, RM, RF0, opts["gtol"], opts["ftol"]])
    np.savez_compressed(os.path.join(HERE, "l96_ladder_golden.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["ode", "nnet", "ladder"]
    if "ode" in what:
        make_ode()
    if "nnet" in what:
        make_nnet()
    if "ladder" in what:
        make_ladder()
