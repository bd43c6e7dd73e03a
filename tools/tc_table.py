"""Bring-up helper: relative error table of the contraction engine over modes, operand majors
and shapes (complete output, no truncation)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _native

def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)

shapes = [(128, 256, 32), (128, 256, 64), (128, 256, 512), (300, 520, 200), (97, 40, 1000), (1024, 1024, 96), (5, 3, 7),
          (256, 256, 2048), (256, 256, 8192)]
modes = sys.argv[1:] or ["tf32", "tf32x3"]
for mode in modes:
    for (a_t, b_t) in [(False, True), (False, False), (True, True), (True, False)]:
        for (M, N, K) in shapes:
            rs = np.random.RandomState(M + N + K)
            A = rs.random_sample((M, K)) + 0.1
            B = rs.random_sample((K, N)) + 0.1
            ref = A.dot(B)
            try:
                out = _native.contract(np.ascontiguousarray(A.T) if a_t else A, np.ascontiguousarray(B.T) if b_t else B,
                                       mode, a_trans=a_t, b_trans=b_t)
                bias = float(np.mean((out - ref) / ref))
                print("%-7s A_%s B_%s %5dx%5dx%5d rel=%.3e bias=%+.3e" % (mode, "MN" if a_t else "K ", "K " if b_t else "MN", M, N, K, rel(out, ref), bias), flush=True)
            except Exception as e:
                print("%-7s A_%s B_%s %5dx%5dx%5d EXC %s" % (mode, "MN" if a_t else "K ", "K " if b_t else "MN", M, N, K, e), flush=True)
                sys.exit(1)
