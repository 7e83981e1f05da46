"""varanneal_b200.datagen against the reference's own generator scripts.

tests/golden/datagen_golden.npz holds extracts of what the unmodified scripts
(examples/nnet_twin/data/gen_params.py, gen_io_pairs.py, examples/nnet_barimages/data/
bardata_gen.py) write, produced by tests/golden/make_datagen_golden.py; the functions reproduce
them bit for bit (same seeds, same random-number call order).  CPU only."""
import os

import numpy as np
import pytest

from varanneal_b200 import datagen

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datagen_golden.npz"))


def test_nnet_twin_params_bitwise():
    sets = datagen.nnet_twin_params([10] * 100, seed=17439860, nsets=3, n_unused=1)
    W1, b1 = sets[0]
    assert len(W1) == 99 and W1[0].shape == (10, 10)
    assert np.array_equal(np.array([W1[0], W1[50], W1[98]]), G["twin/W1_layers_0_50_98"])
    assert np.array_equal(sets[2][0][7], G["twin/W3_layer_7"])
    assert all(np.all(b == 0.0) for b in b1) and G["twin/b1_absmax"] == 0.0
    # rectangular layers: W_n is (d_{n+1}, d_n) scaled by its fan-in
    W, b = datagen.nnet_twin_params([4, 7, 3], seed=1)[0]
    assert W[0].shape == (7, 4) and W[1].shape == (3, 7) and np.max(np.abs(W[0])) <= 0.25


def test_nnet_twin_io_bitwise():
    sets = datagen.nnet_twin_params([10] * 100, seed=17439860, nsets=2, n_unused=1)
    rng = np.random.RandomState(43650832)           # one stream over both teachers, as the script
    for i, (W, b) in enumerate(sets):
        nin, nout, states = datagen.nnet_twin_io(W, b, 25, sigma=0.005, rng=rng)
        io = G["twin/io_param%d" % (i + 1)]
        assert np.array_equal(nin, io[:, 0]) and np.array_equal(nout, io[:, 1])
        assert nout.min() >= 0.0001 and nout.max() <= 0.9999
    ts = G["twin/truestates_p2_ex25_layers_0_1_99"]
    assert np.array_equal(np.array([states[24][0], states[24][1], states[24][99]]), ts)
    assert abs(np.std(states[0][0]) - 1.0) < 1e-12 and abs(np.mean(states[0][0])) < 1e-12


def test_bar_images_bitwise():
    data, lab = datagen.bar_images(5, 1000, "centered", seed=85964309)
    assert list(data.shape) == list(G["bars/centered_shape"]) and lab.dtype == np.int8
    assert np.array_equal(data[:8], G["bars/centered_head"]) and np.array_equal(data[-4:], G["bars/centered_tail"])
    assert np.array_equal(lab[:8], G["bars/centered_labels_head"]) and data.sum() == G["bars/centered_sum"]
    data, lab = datagen.bar_images(5, 3, "allpositions", seed=85964309)
    assert np.array_equal(data, G["bars/allpos_data"]) and np.array_equal(lab, G["bars/allpos_labels"])
    with pytest.raises(ValueError):
        datagen.bar_images(4, 1)


def test_lorenz96_twin_layout_and_bench_recipe():
    """The twin-data recipe bench.py uses: deterministic, on the attractor, noise
    of the requested size, and laid out for set_data (time in column 0)."""
    t, truth, Y = datagen.lorenz96_twin(D=20, N=161, dt=0.025, k=8.17, sigma=0.5, Lidx=[0, 2, 4], seed=100)
    t2, truth2, Y2 = datagen.lorenz96_twin(D=20, N=161, dt=0.025, k=8.17, sigma=0.5, Lidx=[0, 2, 4], seed=100)
    assert np.array_equal(truth, truth2) and np.array_equal(Y, Y2)
    assert truth.shape == (161, 20) and Y.shape == (161, 3) and np.allclose(np.diff(t), 0.025)
    assert 2.0 < truth.std() < 6.0 and abs(truth.mean() - 2.3) < 1.5          # Lorenz96 climatology at k ~ 8
    assert 0.35 < (Y - truth[:, [0, 2, 4]]).std() < 0.65
    # one RK4 step of the returned trajectory reproduces the next row
    x, k, dt = truth[10], 8.17, 0.025
    f = datagen.lorenz96
    k1 = f(x, k); k2 = f(x + 0.5 * dt * k1, k); k3 = f(x + 0.5 * dt * k2, k); k4 = f(x + dt * k3, k)
    assert np.allclose(x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4), truth[11], rtol=0, atol=1e-13)
    from varanneal_b200 import va_ode
    an = va_ode.Annealer()
    an.set_data(np.column_stack([t, Y]))
    assert an.N_data == 161 and an.Y.shape == (161, 3)
