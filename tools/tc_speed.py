"""Device-time table of the contraction engine over operand majors (TFLOP/s)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _native

shapes = [(8192, 8192, 8192), (65536, 8192, 512), (65536, 512, 8192), (512, 8192, 65536)]
for mode in (sys.argv[1:] or ["tf32"]):
    for (M, N, K) in shapes:
        for (a_t, b_t) in [(False, True), (False, False), (True, True), (True, False)]:
            ms = _native.contract_bench(M, N, K, mode, a_trans=a_t, b_trans=b_t, iters=5)
            print("%-7s A_%s B_%s %6dx%6dx%6d %8.3f ms %7.1f TFLOP/s" % (
                mode, "MN" if a_t else "K ", "K " if b_t else "MN", M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
