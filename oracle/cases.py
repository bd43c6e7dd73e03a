"""Seeded inputs shared by the golden generator, the oracle tests and the GPU
parity tests (TEST INFRASTRUCTURE).  Nothing here reads /root/reference."""
import numpy as np
import scipy.sparse as sp


def kat_dense():
    """SURVEY 8c dense known-answer input (also the stale doctest of nmf.py:117-126)."""
    return np.array([[1, 1], [2, 1], [3, 1.2], [4, 1], [5, .8], [6, 1]], dtype=np.float64)


def kat_csr():
    """SURVEY 8c CSR known-answer input."""
    return sp.csr_matrix(np.array([[0, 2, 0, 1], [1, 0, 0, 3], [0, 0, 4, 0],
                                   [2, 1, 0, 0], [0, 0, 1, 1]], dtype=np.float64))


def cfg1_X():
    """BASELINE.json configs[0] / SURVEY 8d cfg1: 500 x 200 uniform, k=10, seed(1)."""
    return np.abs(np.random.RandomState(0).random_sample((500, 200)))


def ragged_dense_X():
    """Shapes that are multiples of nothing: 301 x 157, k=13."""
    rs = np.random.RandomState(21)
    W = rs.gamma(0.7, 1.0, size=(301, 13))
    H = rs.dirichlet(0.3 * np.ones(157), 13)
    return W.dot(H) * 40 + 0.05 * rs.random_sample((301, 157))


def zeros_dense_X():
    """Dense input with ~60 % exact zeros (histogram-like)."""
    rs = np.random.RandomState(33)
    X = rs.poisson(0.6, size=(200, 96)).astype(np.float64)
    X[:, 5] = 0.0          # an all-zero feature
    X[17, :] = 0.0         # an all-zero sample
    return X


def sparse_mid_X():
    """CSR 300 x 400 at 5 % density with count-like values, one empty row."""
    rs = np.random.RandomState(44)
    X = sp.random(300, 400, density=0.05, random_state=rs, format='csr')
    X.data = np.ceil(5 * X.data)
    X = X.tolil()
    X[123, :] = 0
    X = X.tocsr()
    X.eliminate_zeros()
    X.sort_indices()
    return X


def sub_dictionary(k, f):
    """A dictionary whose rows do NOT sum to one (what reconstruct_internal passes,
    learner.py:74)."""
    rs = np.random.RandomState(99)
    H = rs.random_sample((k, f)) + .01
    return 0.5 * H / H.sum(axis=1, keepdims=True)


def learner_small():
    """cfg2 in miniature: dense 'motion' histograms (rows sum to 1) + CSR 'sound'
    counts; coefficients = 1/mean(row sum) as in experiment.py:70-72."""
    rs = np.random.RandomState(0)
    n = 200
    motion = rs.dirichlet(0.1 * np.ones(60), n)
    sound = sp.random(n, 500, density=0.04, random_state=rs, format='csr')
    sound.data = np.ceil(5 * sound.data)
    coefs = [1. / np.mean(motion.sum(axis=1)),
             1. / np.mean(np.asarray(sound.sum(axis=1)))]
    return motion, sound, coefs


def cfg2_inputs(n=1000, f_sound=110000, density=0.01):
    """motion 450 dense Dirichlet rows + HAC-like CSR sound."""
    rs = np.random.RandomState(0)
    motion = rs.dirichlet(0.1 * np.ones(450), n)
    sound = sp.random(n, f_sound, density=density, random_state=rs, format='csr')
    sound.data = np.ceil(5 * sound.data)
    coefs = [1. / np.mean(motion.sum(axis=1)),
             1. / np.mean(np.asarray(sound.sum(axis=1)))]
    return motion, sound, coefs


def rise_case():
    """A fit whose objective RISES between the first and the second evaluation in float64 (found by a seeded search,
    seed 51): zero-heavy 9 x 4 data and an initial dictionary with a huge dynamic range, k = 2.  The reference's
    update (new W with the stale ratio, nmf.py:349) is not monotone by construction; with tol = 0 it breaks here
    (nmf.py:215) with one recorded error.  Returns (X, H0)."""
    rs = np.random.RandomState(51)
    n = rs.randint(3, 10)
    f = rs.randint(3, 10)
    k = rs.randint(1, min(n, f))
    X = (rs.random_sample((n, f)) < 0.4) * rs.random_sample((n, f)) * 10
    H = np.exp(rs.normal(0, 3, size=(k, f)))
    return X, H / (1e-16 + H.sum(axis=1, keepdims=True))


def rel_fro(a, b):
    """Norm-relative error used by every parity test (tiny entries
    differ wildly between paths, so never compare element-relative)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def store_dictionaries():
    """Two small row-normalised dictionaries (4 x 9) for the logger-store fixtures (oracle/make_store_golden.py)."""
    rs = np.random.RandomState(42)
    out = []
    for _ in range(2):
        d = rs.random_sample((4, 9)) + .01
        out.append(d / d.sum(axis=1, keepdims=True))
    return out
