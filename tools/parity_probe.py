"""Runs the golden ladders on the GPU and prints the per-rung parity / restart-acceptance tables
(tests/ladder_parity.py).  Diagnostic companion of tests/test_gpu_ladder_configs.py:
    python tools/parity_probe.py [c1 c2 nakl nnet] > gpurun_out/parity.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ladder_parity as lp    # noqa: E402

what = sys.argv[1:] or ["c1", "c2", "nakl", "nnet"]
np.set_printoptions(linewidth=200)
for w in what:
    runs = {"c1": [lambda: lp.run_c1("trapezoid"), lambda: lp.run_c1("SimpsonHermite")],
            "c2": [lp.run_c2_slice],
            "nakl": [lambda: lp.run_nakl("trapezoid"), lambda: lp.run_nakl("SimpsonHermite")],
            "nnet": [lp.run_nnet]}[w]
    for fn in runs:
        t0 = time.time()
        try:
            an, s, z = fn()
        except Exception as exc:
            print("FAILED %s: %r" % (w, exc), flush=True)
            continue
        print(lp.fmt(s))
        print("  wall %.1f s (incl. %d SciPy restarts); device nfev %d nit %d; exitflags %s; graph cycles %d"
              % (time.time() - t0, len(s["beta"]), int(an.nfev_array.sum()), int(an.nit_array.sum()),
                 np.unique(an.exitflags).tolist(), an._ctx.graph_launches))
        if "counts" in z.files:
            print("  SciPy reference nfev %d nit %d" % (z["counts"][:, 1].sum(), z["counts"][:, 0].sum()))
        for k in ("nactive_dev", "nactive_ref", "inside"):
            if k in s:
                print("  %s: %s" % (k, np.asarray(s[k]).T.tolist() if k != "inside" else s[k]))
        sys.stdout.flush()
