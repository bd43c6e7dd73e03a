"""-m gpu: the dense contraction engines in isolation (through the C ABI's diagnostic entry
point) against float64 numpy, for every operand-major combination the iteration uses and
for ragged shapes.  Tolerances: TF32 keeps 10 mantissa bits per operand (products ~1e-3
relative before averaging), split-TF32 is FP32-grade, DMMA is float64."""
import numpy as np
import pytest

from multimodal_b200 import _native

pytestmark = pytest.mark.gpu

TOL = {"tf32": 2e-3, "tf32x3": 3e-6, "fp64": 1e-13}
# the tensor core truncates when it adds into the FP32 accumulator: measured bias -1.1e-8 * K
# for split-TF32 (three MMAs per 8 of K); see DESIGN.md "Accuracy"
ACC_BIAS_PER_K = {"tf32": 0.0, "tf32x3": 1.4e-8, "fp64": 0.0}

SHAPES = [
    (128, 256, 32),      # exactly one tile, one K block
    (128, 256, 512),     # one tile, 16 K blocks (pipeline wraps)
    (300, 520, 200),     # ragged in M, N and K
    (97, 40, 1000),      # narrow N (BN=128 path), long K
    (1024, 1024, 96),    # many tiles -> persistent loop + TMEM double buffering
    (5, 3, 7),           # tiny
]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("mode", ["fp64", "tf32", "tf32x3"])
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_contract_matches_numpy(mode, a_trans, b_trans, M, N, K):
    rs = np.random.RandomState(M * 7 + N * 3 + K)
    A = rs.random_sample((M, K)) + 0.1
    B = rs.random_sample((K, N)) + 0.1
    ref = A.dot(B)
    A_in = np.ascontiguousarray(A.T) if a_trans else A
    B_in = np.ascontiguousarray(B.T) if b_trans else B
    out = _native.contract(A_in, B_in, mode, a_trans=a_trans, b_trans=b_trans)
    assert out.shape == ref.shape
    assert np.isfinite(out).all()
    assert rel(out, ref) < TOL[mode] + ACC_BIAS_PER_K[mode] * K, (mode, a_trans, b_trans, M, N, K, rel(out, ref))


@pytest.mark.parametrize("mode", ["tf32", "tf32x3"])
def test_contract_is_exact_on_tf32_representable_inputs(mode):
    # small integers are exact in TF32, products/sums exact in fp32: any layout or descriptor
    # mistake shows up as a wrong integer, not as rounding noise
    rs = np.random.RandomState(3)
    A = rs.randint(0, 8, size=(256, 160)).astype(np.float64)
    B = rs.randint(0, 8, size=(160, 384)).astype(np.float64)
    for a_t in (False, True):
        for b_t in (False, True):
            out = _native.contract(np.ascontiguousarray(A.T) if a_t else A,
                                   np.ascontiguousarray(B.T) if b_t else B, mode, a_trans=a_t, b_trans=b_t)
            assert np.array_equal(out, A.dot(B)), (mode, a_t, b_t)


def test_engine_names_are_native():
    for mode, name in [("tf32", "tcgen05_tf32"), ("tf32x3", "tcgen05_tf32x3"), ("fp64", "dmma_f64")]:
        with _native.Engine(4, 4, 2, mode=mode) as e:
            assert e.engine_name == name
