"""Generates tests/golden/datagen_golden.npz from the reference's own data-generator *scripts*.

Run in the build container (needs /root/reference):  python tests/golden/make_datagen_golden.py

The scripts (Python 2, top-level code that writes files) are read as text, given the syntactic
py2->py3 substitutions (xrange, exec statements, integer division of two ints) and -- where the
shipped sizes would write tens of thousands of files -- a smaller example count, and executed in a
scratch directory.  Small extracts of what they write are stored, with the sizes used, so that
tests/test_datagen.py can pin varanneal_b200.datagen on any machine.
"""
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VARANNEAL_REFERENCE", "/root/reference")


def py3(src):
    src = re.sub(r'exec ("[^"]*")%\(([^)]*)\)', r'exec(\1%(\2))', src)
    src = src.replace("xrange", "range")
    src = src.replace("Ndiag/2", "Ndiag//2").replace("dim/2", "dim//2")
    return src


def run_script(path, cwd, subs=()):
    src = py3(open(path).read())
    for a, b in subs:
        assert a in src, a
        src = src.replace(a, b)
    old = os.getcwd()
    os.chdir(cwd)
    try:
        exec(compile(src, path, "exec"), {"__name__": "__main__"})
    finally:
        os.chdir(old)


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- nnet_twin: teacher parameters, then input/output pairs (25 examples instead of 10000)
        d = os.path.join(REF, "examples", "nnet_twin", "data")
        run_script(os.path.join(d, "gen_params.py"), tmp)
        W1 = np.load(os.path.join(tmp, "params", "W_1.npy"))
        W3 = np.load(os.path.join(tmp, "params", "W_3.npy"))
        out["twin/W1_layers_0_50_98"] = W1[[0, 50, 98]]
        out["twin/W3_layer_7"] = W3[7]
        out["twin/b1_absmax"] = np.abs(np.load(os.path.join(tmp, "params", "b_1.npy"))).max()
        run_script(os.path.join(d, "gen_io_pairs.py"), tmp, subs=[("Nexamples = 10000", "Nexamples = 25")])
        for i in (1, 2):
            io = np.array([np.load(os.path.join(tmp, "training", "param%d" % i, "noisyio_sm0p005_%d.npy" % (j + 1)))
                           for j in range(25)])
            out["twin/io_param%d" % i] = io                                   # (25, 2, 10)
        out["twin/truestates_p2_ex25_layers_0_1_99"] = np.array(
            np.load(os.path.join(tmp, "training", "param2", "truestates_25.npy")))[[0, 1, 99]]
    with tempfile.TemporaryDirectory() as tmp:
        # ---- bar images: centred (as shipped, 1000 sets) and all positions (3 sets)
        p = os.path.join(REF, "examples", "nnet_barimages", "data", "bardata_gen.py")
        run_script(p, tmp)
        data, lab = np.load(os.path.join(tmp, "training_data.npy")), np.load(os.path.join(tmp, "training_label.npy"))
        out["bars/centered_shape"] = np.array(data.shape)
        out["bars/centered_head"] = data[:8]
        out["bars/centered_tail"] = data[-4:]
        out["bars/centered_labels_head"] = lab[:8]
        out["bars/centered_sum"] = data.sum()
        run_script(p, tmp, subs=[('imagetype = "centered"  #', 'imagetype = "allpositions"  #'),
                                 ("Nsets = 1000", "Nsets = 3")])
        data, lab = np.load(os.path.join(tmp, "training_data.npy")), np.load(os.path.join(tmp, "training_label.npy"))
        out["bars/allpos_data"] = data
        out["bars/allpos_labels"] = lab
    np.savez_compressed(os.path.join(HERE, "datagen_golden.npz"), **out)
    print("wrote datagen_golden.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
