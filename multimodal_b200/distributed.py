"""Sample-sharded KL-NMF over the GPUs of one box.

One process per GPU.  Rows of X and W are split into contiguous blocks; the k x f
dictionary is replicated.  Per fit iteration the library all-reduces (NCCL, on the
context's stream) the k x f numerator W'^T.Q and the objective partials; transform needs no
communication at all.  torch.distributed is only the bootstrap that carries the
128-byte NCCL unique id and the host-drawn H0 to every rank.
"""
import numpy as np

from . import _native


def shard_bounds(n, world):
    """Contiguous row blocks: rank r owns rows [b[r], b[r+1])."""
    base, rem = divmod(int(n), int(world))
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_object(obj, src=0):
    dist = _dist()
    box = [obj if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_unique_id():
    """Rank 0 asks NCCL for a unique id; everybody gets the same 128 bytes."""
    dist = _dist()
    uid = _native.nccl_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_object(uid, 0)
    assert isinstance(uid, (bytes, bytearray)) and len(uid) == 128
    return bytes(uid)


def draw_shared_dictionary(k, f):
    """H0 exactly as nmf.py:150-151 on rank 0's global numpy RNG, then broadcast."""
    dist = _dist()
    H0 = None
    if dist.get_rank() == 0:
        H0 = np.abs(np.random.random((k, f))) + .01
        H0 = H0 / (1.e-16 + H0.sum(axis=1, keepdims=True))
    return broadcast_object(H0, 0)


class ShardedNMF(object):
    """fit / transform of a row shard; every rank calls the same methods in the same order."""

    def __init__(self, n_components, max_iter=200, tol=1e-6, mode=None, device=0):
        self.n_components, self.max_iter, self.tol = n_components, max_iter, tol
        self.mode, self.device = mode, device
        self.components_ = None

    def _engine(self, X_local, n_global):
        import scipy.sparse as sp
        dist = _dist()
        n, f = X_local.shape
        eng = _native.Engine(n, f, self.n_components, mode=self.mode, device=self.device)
        if sp.issparse(X_local):
            eng.set_csr(X_local)
        else:
            eng.set_dense(X_local)
        if dist.get_world_size() > 1:
            _native.nccl_load()
            eng.comm_init(broadcast_unique_id(), dist.get_rank(), dist.get_world_size())
        return eng

    def fit_transform(self, X_local, n_global, H0=None, fit=True, return_errors=False):
        f = X_local.shape[1]
        if H0 is None:
            H0 = self.components_ if not fit else draw_shared_dictionary(self.n_components, f)
        eng = self._engine(X_local, n_global)
        try:
            eng.set_dictionary(H0)
            eng.init_coefficients()
            errors, n_iter = eng.run(self.max_iter, self.tol * n_global * f, fit)
            W = eng.get_coefficients()
            if fit:
                self.components_ = eng.get_dictionary()
        finally:
            eng.close()
        return (W, list(errors)) if return_errors else W
