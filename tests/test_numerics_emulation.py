"""CPU: the arithmetic of the one-pass `tf32r` mode emulated in numpy (operands rounded to nearest TF32, products
summed exactly) against the float64 oracle.  It pins, without a GPU, the two design decisions DESIGN.md section 2
argues for: the centered ratio with its exact remainders, and the clamp that keeps the re-assembled sums
non-negative where the ratio vanishes."""
import numpy as np

from oracle import cases, klnmf_oracle as O


def rn_tf32(a):
    """float32 -> nearest value with a 10-bit mantissa (cvt.rna.tf32.f32; ties away from zero)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def tf32r_iteration(X, W, H, clamp=True):
    """One fit iteration as libklnmf's tf32r mode computes it: S, G, N from the rounded halves of W, H, Q - 1; the
    FP32 state itself is kept (hi + lo)."""
    eps = O.EPS
    S = rn_tf32(W).dot(rn_tf32(H))
    objective = float(np.sum(X * np.log((X + eps) / (S + eps)) - X + S))
    Qc = rn_tf32((X + eps) / (S + eps) - 1.0)                        # centered, rounded before it is stored
    G = Qc.dot(rn_tf32(H).T) + H.sum(axis=1)                          # + rowsum(H): exact remainder
    if clamp:
        G = np.maximum(G, 0.0)
    Wn = (W * G).astype(np.float32).astype(np.float64)
    N = rn_tf32(Wn).T.dot(Qc) + Wn.sum(axis=0)[:, None]               # + colsum(W'): exact remainder
    if clamp:
        N = np.maximum(N, 0.0)
    Hn = H * N
    Hn = Hn / (O.NORM_EPS + Hn.sum(axis=1, keepdims=True))
    return Wn, Hn.astype(np.float32).astype(np.float64), objective


def test_rn_tf32_keeps_ten_mantissa_bits():
    x = np.float32(1.0) + np.float32(2.0 ** -11)          # half an ulp of TF32 above 1: ties go away from zero
    assert rn_tf32([x])[0] == 1.0 + 2.0 ** -10
    assert rn_tf32([np.float32(1.0) + np.float32(2.0 ** -12)])[0] == 1.0
    v = np.random.RandomState(0).random_sample(1000).astype(np.float32)
    assert np.max(np.abs(rn_tf32(v) - v) / v) <= 2.0 ** -11


def test_tf32r_emulation_tracks_the_oracle():
    rs = np.random.RandomState(3)
    X = rs.gamma(0.5, 1.0, size=(300, 256))
    np.random.seed(5)
    H = O.init_dictionary(64, 256)
    W = np.asarray(X.dot(H.T))
    Wr, Hr = W.copy(), H.copy()
    for _ in range(10):
        e_ref = O.error(X, Wr, Hr)
        Wr, Hr = O.update(X, Wr, Hr, fit=True)
        W, H, e = tf32r_iteration(X, W, H)
        assert abs(e - e_ref) < 1e-5 * abs(e_ref)
    # the stated tolerance of the mode is 5e-4 (tests/test_gpu_parity.py); 1.1e-4 on this 300 x 256, k = 64 case, and
    # 1e-5..4e-5 at the benchmark shapes, where more terms average the operand rounding
    assert cases.rel_fro(W, Wr) < 3e-4 and cases.rel_fro(H, Hr) < 3e-4
    assert (W >= 0).all() and (H >= 0).all()


def test_clamp_restores_non_negativity_on_all_zero_samples():
    # a sample with no observation: its ratio is eps / (s + eps) ~ 0, the centered ratio -1, and
    # (Q-1).H^T + rowsum(H) = sum of the rounding residues of H -- either sign without the clamp
    X = cases.zeros_dense_X().copy()
    X[::7] = 0.0
    np.random.seed(7)
    H = O.init_dictionary(8, X.shape[1])
    W = np.asarray(X.dot(H.T)) + 1e-3                      # keep the empty samples' coefficients alive for one step
    Wn_plain, _, _ = tf32r_iteration(X, W, H, clamp=False)
    Wn, Hn, _ = tf32r_iteration(X, W, H, clamp=True)
    assert (Wn >= 0).all() and (Hn >= 0).all()
    assert Wn_plain.min() < 0 or np.allclose(Wn_plain, Wn)   # the hazard is real on this data, or absent altogether
    empty = np.arange(X.shape[0])[::7]
    assert np.abs(Wn[empty]).max() < 1e-6                  # exact arithmetic gives ~1e-8 W there; the clamp gives 0


def _rz32(x):
    """float64 -> float32 rounded toward zero (the accumulator update of the tensor core)."""
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y


def _accumulate_truncating(products):
    """Sum along axis 1 the way one tcgen05 pass does: 8 products per step added exactly, the FP32 accumulator truncated."""
    acc = np.zeros(products.shape[0], dtype=np.float32)
    for j in range(0, products.shape[1], 8):
        acc = _rz32(acc.astype(np.float64) + products[:, j:j + 8].sum(axis=1))
    return acc.astype(np.float64)


def test_truncating_accumulation_drifts_with_the_length_and_centering_removes_it():
    # a contraction of non-negative terms loses ~2.5e-9 per term of relative accuracy to the
    # truncating FP32 accumulator (2e-5 at f = 8192 per pass, three passes in tf32x3: the 8e-5 measured on W); the
    # centered ratio sums terms of both signs and adds the exact remainder rowsum(H), and the drift is gone
    rs = np.random.RandomState(0)
    drift = {}
    for f in (1024, 8192):
        q = rs.gamma(1.0, 1.0, size=(200, f))
        h = rs.random_sample((200, f)) / f
        exact = (q * h).sum(axis=1)
        plain = _accumulate_truncating(q * h)
        centered = _accumulate_truncating((q - 1.0) * h) + h.sum(axis=1)
        drift[f] = np.mean((plain - exact) / exact)
        assert abs(np.mean((centered - exact) / exact)) < 1e-7
        assert np.sqrt(np.mean(((centered - exact) / exact) ** 2)) < 1e-6
    assert -4e-6 < drift[1024] < -1.5e-6 and -3e-5 < drift[8192] < -1.2e-5      # always low, proportional to f
    assert 6.0 < drift[8192] / drift[1024] < 10.0
