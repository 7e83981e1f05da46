"""Golden fixtures for the matrix form of RM -- (L, L) or (N_data, L, L), va_ode.py:149-152, 612-621 --
generated from the reference itself.  Run in the build container (needs /root/reference):

    python tests/golden/make_rm_matrix_golden.py

Action values are the verbatim reference's (A_gaussian / me_gaussian / fe_gaussian through
oracle.ref_shim), gradients are complex-step differentiation through it.  The matrices are
deliberately *not* symmetric -- the gradient of diff.(RM diff) is (RM + RM') diff -- but their
symmetric part is positive definite, so the action can be minimised.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle.models_np import MODELS              # noqa: E402
from oracle.ode_port import OdeProblem           # noqa: E402

REF = ref_shim.REFERENCE_ROOT
L96_FILE = os.path.join(REF, "examples", "Lorenz96_D20", "l96_D20_dt0p025_N161_sm0p5_sec1_mem1.npy")
ShimOde, _ = ref_shim.make_shim_classes()
LIDX = [0, 2, 4, 6, 8, 10, 14, 16]


def main():
    data = np.load(L96_FILE)
    t, Yall = data[:, 0], data[:, 1:]
    rng = np.random.RandomState(2024)
    L = len(LIDX)
    out, names = {}, []
    for name, disc, nd, dt_model, per_time in (("LL_trapezoid", "trapezoid", 41, None, False),
                                               ("NLL_simpson", "SimpsonHermite", 41, None, True),
                                               ("NLL_nskip2_euler", "euler", 21, 0.0125, True),
                                               ("LL_lorenz63", "trapezoid", 31, None, False)):
        if "lorenz63" in name:
            model, D, Lidx, P0, Pidx = "lorenz63", 3, [0, 2], np.array([10.0, 28.0, 8.0 / 3.0]), [1]
            tt = 0.01 * np.arange(nd)
            Y = rng.randn(nd, 2) * 5.0
        else:
            model, D, Lidx, P0, Pidx = "lorenz96", 20, LIDX, np.array([8.17]), [0]
            tt, Y = t[:nd], Yall[:nd][:, LIDX]
        Lc = len(Lidx)
        def mat():          # positive definite symmetric part (the action stays bounded below) + a skew part
            G, H = rng.randn(Lc, Lc), rng.randn(Lc, Lc)
            return 3.0 * np.eye(Lc) + G.dot(G.T) / Lc + 0.8 * (H - H.T)
        RM = np.array([mat() for _ in range(nd)]) if per_time else mat()
        an = ShimOde()
        an.set_model(MODELS[model], D)
        an.set_data(Y, t=tt)
        nskip = 1 if dt_model is None else int(round((tt[1] - tt[0]) / dt_model))
        N = (nd - 1) * nskip + 1
        X0 = 20.0 * rng.rand(N, D) - 10.0
        alpha, beta, RF0 = 1.5, 14, 4e-6
        with contextlib.redirect_stdout(io.StringIO()):
            an.anneal_init(X0.copy(), P0.copy(), alpha, [beta], RM.copy(), RF0, np.array(Lidx), Pidx,
                           dt_model=dt_model, init_to_data=False, disc=disc)
        XP = np.append(X0.ravel(), P0[Pidx])
        A = float(an.A(XP))
        me = float(an.me_gaussian(XP[:N * D]))
        fe = float(an.fe_gaussian(XP))
        g = ref_shim.complex_step_grad(an.A, XP)
        RMfull = RM if RM.ndim == 3 else np.resize(RM, (nd, Lc, Lc))
        prob = OdeProblem(model, D, Y, Lidx, an.dt_model, disc, P0, Pidx, RMfull, nskip=nskip)
        Ap, mep, fep, gp = prob.action_grad(XP, RF0 * alpha ** beta, parts=True)
        print("%-18s A=%.16e me=%.6e port rel %.1e / %.1e grad rel %.1e" % (
            name, A, me, abs(Ap - A) / abs(A), abs(mep - me) / abs(me), np.max(np.abs(gp - g)) / np.max(np.abs(g))))
        names.append(name)
        for k, v in (("X0", X0), ("P0", P0), ("t", tt), ("Y", Y), ("Lidx", np.asarray(Lidx, dtype=np.int64)),
                     ("Pidx", np.asarray(Pidx, dtype=np.int64)), ("RM", RM),
                     ("meta", np.array([alpha, beta, RF0, -1.0 if dt_model is None else dt_model])),
                     ("model_disc", np.array([model, disc])), ("A", np.array([A, me, fe])), ("grad", g)):
            out[name + "/" + k] = np.asarray(v)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ode_rm_matrix_golden.npz"), **out)


def main_nnet():
    """va_nnet: RM = [RM_in, RM_out] as one (2, L, L) array (va_nnet.py:135-139; reachable in the
    reference only when as many input as output components are measured)."""
    from oracle import nnet_port
    _, ShimNnet = ref_shim.make_shim_classes()
    rng = np.random.RandomState(77)
    out, names = {}, []
    for name, structure, M, Lidx in (("nn_8_12_8_full", [8, 12, 8], 7, [np.arange(8), np.arange(8)]),
                                     ("nn_7_10_9_partial", [7, 10, 9], 5, [np.array([0, 2, 5]), np.array([1, 4, 8])])):
        structure = np.array(structure)
        NDnet = int(structure.sum())
        L = len(Lidx[0])
        data_in, data_out = rng.rand(M, L), rng.rand(M, L)
        NP = int(sum(structure[n] * structure[n + 1] + structure[n + 1] for n in range(len(structure) - 1)))
        X0 = rng.rand(M * NDnet)
        P0 = 0.4 * rng.randn(NP)
        Pidx = np.arange(NP)

        def mat():
            G, H = rng.randn(L, L), rng.randn(L, L)
            return 2.0 * np.eye(L) + G.dot(G.T) / L + 0.5 * (H - H.T)
        RM = np.array([mat(), mat()])
        RF0, alpha, beta = 1e-3, 1.1, 30
        an = ShimNnet()
        an.set_structure(structure)
        an.set_activation(nnet_port.sigmoid)
        an.set_input_data(data_in)
        an.set_output_data(data_out)
        with contextlib.redirect_stdout(io.StringIO()):
            an.anneal_init(X0.copy(), P0.copy(), alpha, [beta], RM.copy(), RF0, Pidx, Lidx=Lidx, init_to_data=False)
        XP = np.append(X0, P0[Pidx])
        A = float(an.A(XP))
        me = float(an.me_gaussian(XP[:M * NDnet]))
        fe = float(an.fe_gaussian(XP))
        g = ref_shim.complex_step_grad(an.A, XP)
        prob = nnet_port.NnetProblem(structure, data_in, data_out, Lidx, P0, Pidx, RM)
        Ap, gp = prob.action_grad(XP, RF0 * alpha ** beta)
        print("%-20s A=%.16e me=%.6e port rel %.1e grad rel %.1e" % (
            name, A, me, abs(Ap - A) / abs(A), np.max(np.abs(gp - g)) / np.max(np.abs(g))))
        names.append(name)
        for k, v in (("structure", structure), ("data_in", data_in), ("data_out", data_out), ("X0", X0), ("P0", P0),
                     ("Pidx", Pidx), ("Lin", Lidx[0]), ("Lout", Lidx[1]), ("RM", RM),
                     ("meta", np.array([RF0, alpha, beta])), ("A", np.array([A, me, fe])), ("grad", g)):
            out[name + "/" + k] = np.asarray(v)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "nnet_rm_matrix_golden.npz"), **out)


def main_rf():
    """Matrix RF0 -- (D, D) or (N_model-1, D, D), va_ode.py:211-223, 629-636 -- with SimpsonHermite,
    the only discretisation for which the reference's branch runs (:222 contracts the whole residual
    array for the others).  RF = RF0 * alpha**beta scales the matrices (va_ode.py:650)."""
    data = np.load(L96_FILE)
    t, Yall = data[:, 0], data[:, 1:]
    rng = np.random.RandomState(909)
    out, names = {}, []
    for name, model, per_time in (("rf_DD_l96", "lorenz96", False), ("rf_NDD_l96", "lorenz96", True),
                                  ("rf_NDD_l63", "lorenz63", True)):
        nd = 41 if model == "lorenz96" else 31
        if model == "lorenz63":
            D, Lidx, P0, Pidx = 3, [0, 2], np.array([10.0, 28.0, 8.0 / 3.0]), [1]
            tt = 0.01 * np.arange(nd)
            Y = rng.randn(nd, 2) * 5.0
        else:
            D, Lidx, P0, Pidx = 20, LIDX, np.array([8.17]), [0]
            tt, Y = t[:nd], Yall[:nd][:, LIDX]

        def mat():
            G, H = rng.randn(D, D), rng.randn(D, D)
            return 1e-3 * (2.0 * np.eye(D) + G.dot(G.T) / D + 0.6 * (H - H.T))
        RF0 = np.array([mat() for _ in range(nd - 1)]) if per_time else mat()
        an = ShimOde()
        an.set_model(MODELS[model], D)
        an.set_data(Y, t=tt)
        X0 = 20.0 * rng.rand(nd, D) - 10.0
        alpha, beta, RM = 1.5, 9, 4.0
        with contextlib.redirect_stdout(io.StringIO()):
            an.anneal_init(X0.copy(), P0.copy(), alpha, [beta], RM, RF0.copy(), np.array(Lidx), Pidx,
                           dt_model=None, init_to_data=False, disc="SimpsonHermite")
        XP = np.append(X0.ravel(), P0[Pidx])
        A = float(an.A(XP))
        me = float(an.me_gaussian(XP[:nd * D]))
        fe = float(an.fe_gaussian(XP))
        g = ref_shim.complex_step_grad(an.A, XP)
        RFfull = RF0 if RF0.ndim == 3 else np.resize(RF0, (nd - 1, D, D))
        prob = OdeProblem(model, D, Y, Lidx, an.dt_model, "SimpsonHermite", P0, Pidx, RM)
        Ap, mep, fep, gp = prob.action_grad(XP, RFfull * alpha ** beta, parts=True)
        print("%-12s A=%.16e fe=%.6e port rel %.1e / %.1e grad rel %.1e" % (
            name, A, fe, abs(Ap - A) / abs(A), abs(fep - fe) / abs(fe), np.max(np.abs(gp - g)) / np.max(np.abs(g))))
        names.append(name)
        for k, v in (("X0", X0), ("P0", P0), ("t", tt), ("Y", Y), ("Lidx", np.asarray(Lidx, dtype=np.int64)),
                     ("Pidx", np.asarray(Pidx, dtype=np.int64)), ("RF0", RF0), ("meta", np.array([alpha, beta, RM])),
                     ("model", np.array([model])), ("A", np.array([A, me, fe])), ("grad", g)):
            out[name + "/" + k] = np.asarray(v)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ode_rf_matrix_golden.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["ode", "nnet", "rf"]
    if "rf" in what:
        main_rf()
    if "ode" in what:
        main()
    if "nnet" in what:
        main_nnet()
