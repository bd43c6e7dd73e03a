"""Bring-up helper: sweep MN-major descriptor candidates (env overrides) on exact integer inputs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _native

cands = [  # (KLNMF_TC_MN "lt,lbo,sbo,kadv", KLNMF_TC_MN_TMA enum)  TMA: 3=128B 4=128B_ATOM_32B
    ("1,4096,512,1024", "4"),
    ("1,512,4096,1024", "4"),
    ("1,4096,1024,1024", "4"),
    ("2,4096,1024,1024", "3"),
    ("2,4096,1024,1024", "4"),
    ("1,4096,512,1024", "3"),
]
rs = np.random.RandomState(1)
for (M, N, K) in [(128, 256, 32), (128, 256, 96), (256, 512, 160)]:
    A = rs.randint(0, 8, size=(M, K)).astype(np.float64)
    B = rs.randint(0, 8, size=(K, N)).astype(np.float64)
    ref = A.dot(B)
    for mn, tma in cands:
        os.environ["KLNMF_TC_MN"] = mn
        os.environ["KLNMF_TC_MN_TMA"] = tma
        res = []
        for (a_t, b_t) in [(False, False), (True, True), (True, False)]:
            try:
                out = _native.contract(np.ascontiguousarray(A.T) if a_t else A, np.ascontiguousarray(B.T) if b_t else B,
                                       "tf32", a_trans=a_t, b_trans=b_t)
                res.append("%s%s:%d" % ("MN" if a_t else "K", "K" if b_t else "MN", int((out != ref).sum())))
            except Exception as e:
                res.append("EXC %s" % e)
                print(M, N, K, mn, tma, res, flush=True)
                sys.exit(1)
        print("%dx%dx%d MN=%s TMA=%s wrong entries: %s" % (M, N, K, mn, tma, " ".join(res)), flush=True)
