"""Bring-up helper: probes the tcgen05 engine with structured inputs and prints which
(row, col, k) the hardware actually read, per operand-major combination."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _native


def probe(mode, a_t, b_t, M, N, K):
    k_idx = np.arange(K, dtype=np.float64)
    # out[i,j] = sum_k A[i,k] B[k,j]
    tests = {
        "ones":  (np.ones((M, K)), np.ones((K, N))),                                   # expect K
        "rowid": (np.tile(np.arange(M)[:, None] % 64, (1, K)) / K, np.ones((K, N))),   # expect i%64
        "colid": (np.ones((M, K)) / K, np.tile(np.arange(N)[None, :] % 64, (K, 1))),   # expect j%64
        "kselA": (np.eye(M, K), np.tile(k_idx[:, None] % 64, (1, N))),                 # expect i%64 (i<K)
    }
    for name, (A, B) in tests.items():
        ref = A.dot(B)
        Ai = np.ascontiguousarray(A.T) if a_t else A
        Bi = np.ascontiguousarray(B.T) if b_t else B
        try:
            out = _native.contract(Ai, Bi, mode, a_trans=a_t, b_trans=b_t)
        except Exception as e:
            print("  %-6s EXC %s" % (name, e)); return False
        bad = np.abs(out - ref) > 1e-2 * max(1.0, np.abs(ref).max())
        print("  %-6s a_t=%d b_t=%d %dx%dx%d  bad=%d/%d maxerr=%.3g" % (name, a_t, b_t, M, N, K, bad.sum(), bad.size,
                                                                          np.abs(out - ref).max()))
        if bad.any():
            ii, jj = np.nonzero(bad)
            print("    bad rows: %s ... cols: %s ..." % (sorted(set(ii.tolist()))[:12], sorted(set(jj.tolist()))[:12]))
            for t in range(min(4, len(ii))):
                print("    out[%d,%d]=%.4g ref=%.4g" % (ii[t], jj[t], out[ii[t], jj[t]], ref[ii[t], jj[t]]))
            print("    out[0,:8]=%s" % np.array2string(out[0, :8], precision=3))
            print("    out[:8,0]=%s" % np.array2string(out[:8, 0], precision=3))
    return True


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    for (a_t, b_t) in [(False, True), (False, False), (True, False), (True, True)]:
        print("== mode %s a_trans=%s b_trans=%s" % (mode, a_t, b_t))
        for (M, N, K) in [(128, 256, 32), (128, 128, 64), (256, 512, 96)]:
            if not probe(mode, a_t, b_t, M, N, K):
                sys.exit(1)
