"""Synthetic-data generators of the reference's examples, as functions.

The reference ships them as stand-alone Python 2 scripts that write files; here they return arrays
(host NumPy -- this is input preparation, not the hot path) with the *same random-number call
order*, so a given seed reproduces the reference's files bit for bit
(tests/test_datagen.py pins that against the scripts themselves / golden extracts):

  * ``nnet_twin_params``  examples/nnet_twin/data/gen_params.py:11-52
  * ``nnet_twin_io``      examples/nnet_twin/data/gen_io_pairs.py:8-64
  * ``bar_images``        examples/nnet_barimages/data/bardata_gen.py:21-140
  * ``lorenz96_twin``     the twin-experiment recipe of the Lorenz96 example (data file
                          l96_D20_dt0p025_N161_sm0p5_sec1_mem1.npy: RK4 trajectory + N(0, sm^2)
                          noise, time in column 0); the reference ships the file, not the script,
                          so this one follows SURVEY.md 8(d) and is not pinned.
"""
import numpy as np


# ------------------------------------------------------------------ neural-network twin experiment
def nnet_twin_params(structure, seed=17439860, nsets=1, n_unused=0):
    """Teacher networks: W_n ~ U(-1, 1) / fan_in, b_n = 0 (gen_params.py:30-46).  Returns a list of
    ``nsets`` pairs (W, b), W a list of (d_{n+1}, d_n) arrays; the sets are drawn one after the
    other from one RandomState(seed) stream, like the script's loop over Nparamsets.

    ``n_unused``: the script draws N weight matrices for its N-layer network (first, N-2 hidden,
    last) although N layers have N-1 connections and gen_io_pairs.py:51 only uses the first N-1;
    pass ``n_unused=1`` to draw (and drop) that extra matrix after each set, which is what keeps
    the random stream aligned with the shipped parameter files from the second set on."""
    structure = [int(d) for d in structure]
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(nsets):
        W, b = [], []
        for n in range(len(structure) - 1):
            d_in, d_out = structure[n], structure[n + 1]
            W.append((2.0 * rng.rand(d_out, d_in) - 1.0) / float(d_in))
            b.append(np.zeros(d_out))
        for _ in range(n_unused):
            rng.rand(structure[-1], structure[-2])
        out.append((W, b))
    return out


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def nnet_twin_io(W, b, M, sigma=0.005, seed=43650832, rng=None):
    """Input/output pairs of a teacher network (gen_io_pairs.py:45-64): inputs ~ N(0, 1)
    standardised per example, outputs through the sigmoid chain, Gaussian noise of standard
    deviation ``sigma`` on both, the output clipped to [1e-4, 0.9999].  Returns
    (noisy_in (M, d_0), noisy_out (M, d_last), truestates list of M lists of layer states).
    Pass ``rng`` to continue one stream over several teachers, as the script does."""
    rng = np.random.RandomState(seed) if rng is None else rng
    d_in, d_out = W[0].shape[1], W[-1].shape[0]
    nin = np.zeros((M, d_in))
    nout = np.zeros((M, d_out))
    states = []
    for j in range(M):
        yin = rng.randn(d_in)
        yin = (yin - np.average(yin)) / np.std(yin)
        y = [yin]
        for n in range(len(W)):
            y.append(_sigmoid(np.dot(W[n], y[n]) + b[n]))
        states.append(y)
        if sigma > 0:
            nin[j] = y[0] + sigma * rng.randn(d_in)
            nout[j] = np.clip(y[-1] + sigma * rng.randn(d_out), 0.0001, 0.9999)
        else:
            nin[j], nout[j] = y[0], y[-1]
    return nin, nout, states


# ------------------------------------------------------------------ bar images
def _bar_templates(dim):
    h = [np.zeros((dim, dim), dtype=int) for _ in range(dim)]
    v = [np.zeros((dim, dim), dtype=int) for _ in range(dim)]
    for i in range(dim):
        v[i][:, i] = 1
        h[i][i, :] = 1
    ndiag = 2 * dim - 3
    d1, d2 = [], []
    for i in range(ndiag):
        ones = np.ones(i + 2) if i <= ndiag // 2 else np.ones(ndiag - i + 1)
        d = np.diag(ones, -(ndiag // 2) + i)
        d1.append(d)
        d2.append(np.fliplr(d))
    return h, v, d1, d2, ndiag


def bar_images(dim=5, Nsets=1000, imagetype="centered", seed=85964309):
    """Noisy images of horizontal / vertical / diagonal bars with one-hot labels
    (bardata_gen.py:59-140).  ``centered``: 4 images per set (h, v, d1, d2 through the centre,
    labels [1,0,0,0], [0,0,1,0], [0,1,0,0], [0,0,0,1]); ``allpositions``: every bar position.
    Pixel noise 0.35 (0.5 - U(0, 1)).  Returns (data (M, dim*dim) float64, labels (M, 4) int8)."""
    if dim < 3 or dim % 2 == 0:
        raise ValueError("image dimension must be odd and >= 3")
    rng = np.random.RandomState(seed)
    size = (dim, dim)
    h, v, d1, d2, ndiag = _bar_templates(dim)
    noisy = lambda img: (img + 0.35 * (0.5 - rng.random_sample(size))).flatten()  # noqa: E731
    LH, LD1, LV, LD2 = [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]
    if imagetype == "centered":
        data = np.zeros((4 * Nsets, dim * dim))
        labels = np.zeros((4 * Nsets, 4), dtype=np.int8)
        hv, dg = dim // 2, ndiag // 2
        for m in range(Nsets):
            data[4 * m] = noisy(h[hv]); labels[4 * m] = LH
            data[4 * m + 1] = noisy(v[hv]); labels[4 * m + 1] = LV
            data[4 * m + 2] = noisy(d1[dg]); labels[4 * m + 2] = LD1
            data[4 * m + 3] = noisy(d2[dg]); labels[4 * m + 3] = LD2
        return data, labels
    if imagetype == "allpositions":
        M = Nsets * (2 * dim + 2 * ndiag)
        data = np.zeros((M, dim * dim))
        labels = np.zeros((M, 4), dtype=np.int8)
        idx = 0
        for m in range(Nsets):
            for group, lab in ((h, LH), (v, LV), (d1, LD1), (d2, LD2)):
                for img in group:
                    data[idx] = noisy(img)
                    labels[idx] = lab
                    idx += 1
        return data, labels
    raise ValueError("imagetype must be 'centered' or 'allpositions'")


# ------------------------------------------------------------------ Lorenz96 twin experiment
def lorenz96(x, k):
    return np.roll(x, 1, -1) * (np.roll(x, -1, -1) - np.roll(x, 2, -1)) - x + k


def lorenz96_twin(D=20, N=161, dt=0.025, k=8.17, sigma=0.5, Lidx=None, seed=100, transient=1000):
    """RK4 trajectory of Lorenz96 from x_i = k + N(0, 1) after ``transient`` discarded steps, and
    noisy observations of the components ``Lidx`` (all by default).  Returns (t (N,), truth (N, D),
    Y (N, L)); ``np.column_stack([t, Y])`` is the layout of the shipped data file / set_data."""
    rng = np.random.RandomState(seed)
    x = k + rng.randn(D)
    rows = []
    for n in range(transient + N):
        k1 = lorenz96(x, k)
        k2 = lorenz96(x + 0.5 * dt * k1, k)
        k3 = lorenz96(x + 0.5 * dt * k2, k)
        k4 = lorenz96(x + dt * k3, k)
        x = x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        if n >= transient:
            rows.append(x.copy())
    truth = np.array(rows)
    Lidx = list(range(D)) if Lidx is None else list(Lidx)
    Y = truth[:, Lidx] + sigma * np.random.RandomState(seed + 1).randn(N, len(Lidx))
    return dt * np.arange(N), truth, Y
