"""2+ rank check on real GPUs: the sharded fit (NCCL all-reduce of the numerator) against the
float64 oracle on the full matrix.  Launched with torch.distributed.run."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from multimodal_b200 import distributed as D
from oracle import cases, klnmf_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rs = np.random.RandomState(3)
X = rs.gamma(0.5, 1.0, size=(1500, 700))
k = 48
np.random.seed(4)
H0 = O.init_dictionary(k, X.shape[1])
b = D.shard_bounds(X.shape[0], world)
for mode, tol in (("fp64", 1e-9), ("tf32x3", 2e-5), ("tf32r", 3e-4), ("tf32", 3e-4)):
    sh = D.ShardedNMF(k, max_iter=10, tol=0, mode=mode, device=local)
    W, errs = sh.fit_transform(X[b[rank]:b[rank + 1]], X.shape[0], H0=H0, fit=True, return_errors=True)
    np.random.seed(4)
    Wr, Hr, er, _ = O.fit_transform(X, k=k, max_iter=10, tol=0)
    eW = cases.rel_fro(W, Wr[b[rank]:b[rank + 1]]); eH = cases.rel_fro(sh.components_, Hr)
    eK = abs(errs[-1] - er[-1]) / abs(er[-1])
    print("rank %d mode %-6s relerr W %.2e H %.2e KL %.2e %s" % (rank, mode, eW, eH, eK, "ok" if max(eW, eH) < tol else "FAIL"), flush=True)
    assert max(eW, eH) < tol and len(errs) == 10
# sparse shard (CSR rows) through the same all-reduce
import scipy.sparse as sp
Xs = sp.random(1200, 900, density=0.05, random_state=np.random.RandomState(8), format="csr")
Xs.data = np.ceil(5 * Xs.data)
np.random.seed(6)
H0s = O.init_dictionary(24, Xs.shape[1])
bs = D.shard_bounds(Xs.shape[0], world)
for mode, tol in (("fp64", 1e-9), ("tf32", 2e-5)):
    sh = D.ShardedNMF(24, max_iter=10, tol=0, mode=mode, device=local)
    W, errs = sh.fit_transform(Xs[bs[rank]:bs[rank + 1]], Xs.shape[0], H0=H0s, fit=True, return_errors=True)
    np.random.seed(6)
    Wr, Hr, er, _ = O.fit_transform(Xs.copy(), k=24, max_iter=10, tol=0)
    eW = cases.rel_fro(W, Wr[bs[rank]:bs[rank + 1]]); eH = cases.rel_fro(sh.components_, Hr)
    print("rank %d sparse mode %-6s relerr W %.2e H %.2e %s" % (rank, mode, eW, eH, "ok" if max(eW, eH) < tol else "FAIL"), flush=True)
    assert max(eW, eH) < tol and len(errs) == 10
dist.barrier()
dist.destroy_process_group()
