"""Sample-sharded KL-NMF over the GPUs of one box.

Rows of X and W are split into contiguous blocks, one per GPU; the k x f dictionary is replicated.  Per fit iteration
the library all-reduces (NCCL, over NVLink) the k x f numerator W'^T.Q and the objective partials; transform needs no
communication at all.  Two ways to drive the shards:

* `DeviceGroup` -- ONE process, one host thread and one engine per GPU.  This is what `KLdivNMF(device=[0, 1, ...])`
  and `MultimodalLearner(..., device=[...])` use, so `learner.train` (learner.py:31-41) shards over the box without
  `torchrun`; every thread uploads its own row block, so the host->device copies run in parallel over the GPUs' own
  PCIe links.
* `ShardedNMF` -- one process per GPU under `torchrun`; torch.distributed is only the bootstrap that carries the
  128-byte NCCL unique id and the host-drawn H0 to every rank.

NCCL communicators are cached per process (`ncclCommInitRank` costs 0.2-1 s; an engine is created per call).
"""
import threading

import numpy as np

from . import _native


def shard_bounds(n, world, weights=None):
    """Contiguous row blocks: rank r owns rows [b[r], b[r+1]).  `weights` (one positive number per rank, e.g. measured
    samples per second) sizes the blocks in proportion -- the GPUs of one box differ by several per cent under their
    power caps, and every fit iteration ends in an all-reduce that waits for the slowest shard."""
    n, world = int(n), int(world)
    if weights is None:
        base, rem = divmod(n, world)
        b = [0]
        for r in range(world):
            b.append(b[-1] + base + (1 if r < rem else 0))
        return b
    w = np.asarray(weights, dtype=np.float64)
    assert w.shape == (world,) and (w > 0).all(), weights
    edges = np.floor(np.cumsum(w) / w.sum() * n + 0.5).astype(np.int64)
    edges[-1] = n
    return [0] + [int(e) for e in np.maximum.accumulate(edges)]


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


# ---- communicator caches -------------------------------------------------------------------------------------
def _hybrid_for_all_shards(X, bounds, mode, min_cols=1024):
    """A mixed (dense + CSR) stack on several GPUs: its shards must ALL keep the dense modalities dense (hybrid stack,
    klnmf_set_stacked_blocks_host) or all build the CSR stack -- their numerators are summed element by element.  The
    library decides per context (dense width, arithmetic mode); here the same rule is applied once for all shards.
    None: X is not a mixed stack."""
    import os
    import scipy.sparse as sp
    from .lib.array_utils import MixedBlocks
    if not isinstance(X, MixedBlocks):
        return None
    env = os.environ.get("KLNMF_HYBRID")
    if env is not None:
        min_cols = 0 if int(env) == 0 else 1
    fd = sum(b.shape[1] for b in X.blocks if not sp.issparse(b))
    return bool(min_cols > 0 and fd >= min_cols and _native.resolve_mode(mode) != _native.resolve_mode("tf32x3"))


_RANK_COMM = {}      # torchrun: (device, rank, world) -> _native.Comm
_GROUP_COMMS = {}    # one process: tuple(devices) -> [_native.Comm per device]
_CACHE_LOCK = threading.Lock()


def rank_comm(device):
    """The communicator of this torchrun rank, created (collectively) on first use and kept for the process."""
    dist = _dist()
    key = (int(device), dist.get_rank(), dist.get_world_size())
    comm = _RANK_COMM.get(key)
    if comm is None:
        _native.nccl_load()
        comm = _native.Comm(device, broadcast_unique_id(), key[1], key[2])
        _RANK_COMM[key] = comm
    return comm


def _in_threads(fns):
    """Run the callables concurrently (one per device); re-raise the first exception."""
    errs = [None] * len(fns)
    outs = [None] * len(fns)

    def wrap(i):
        try:
            outs[i] = fns[i]()
        except BaseException as e:          # noqa: B902 -- re-raised below
            errs[i] = e
    ts = [threading.Thread(target=wrap, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    real = [e for e in errs if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]                      # the cause, not the BrokenBarrierError it gave the other threads
    for e in errs:
        if e is not None:
            raise e
    return outs


def group_comms(devices):
    """One communicator per device of this process (ncclCommInitRank from one thread per device), cached."""
    key = tuple(int(d) for d in devices)
    with _CACHE_LOCK:
        comms = _GROUP_COMMS.get(key)
        if comms is None:
            uid = _native.nccl_unique_id()
            comms = _in_threads([(lambda r=r, d=d: _native.Comm(d, uid, r, len(key))) for r, d in enumerate(key)])
            _GROUP_COMMS[key] = comms
    return comms


def _rows(X, r0, r1):
    """Row block of a dense array (a view), a CSR matrix (a copy of the block) or a StackedBlocks stack."""
    from .lib.array_utils import StackedBlocks, MixedBlocks
    if isinstance(X, StackedBlocks):
        return StackedBlocks([b[r0:r1] for b in X.blocks], X.coefs)
    if isinstance(X, MixedBlocks):
        return MixedBlocks([b[r0:r1] for b in X.blocks], X.coefs)
    return X[r0:r1]


class DeviceGroup(object):
    """fit / transform over several GPUs of one box from ONE process: rows sharded, one engine and one host thread
    per GPU (ctypes releases the GIL inside the library), NCCL all-reduce of the numerator between the engines."""

    def __init__(self, devices, mode=None):
        self.devices = [int(d) for d in devices]
        assert len(self.devices) >= 1 and len(set(self.devices)) == len(self.devices), devices
        self.mode = mode

    def run(self, X, k, H_init, H_loop, max_iter, tol_abs, fit, want_coefficients=True, set_data=None):
        """The body of KLdivNMF.fit_transform (nmf.py:193-222) on row shards.  Returns
        (W or None, H after the run, errors, n_iter, (negative, non_finite))."""
        n = X.shape[0]
        world = len(self.devices)
        bounds = shard_bounds(n, world)
        comms = group_comms(self.devices) if world > 1 else [None]
        W = np.empty((n, k), dtype=np.float64) if want_coefficients else None
        flags = [None] * world
        results = [None] * world
        gate = threading.Barrier(world)
        hybrid = _hybrid_for_all_shards(X, bounds, self.mode)

        def shard(r):
            r0, r1 = bounds[r], bounds[r + 1]
            eng = _native.Engine(r1 - r0, X.shape[1], k, mode=self.mode, device=self.devices[r])
            try:
                if world > 1 and hybrid is not None:
                    eng.set_hybrid_min_cols(1 if hybrid else 0)   # every shard takes the same form of a mixed stack
                set_data(eng, _rows(X, r0, r1))
                flags[r] = eng.check_input()
                gate.wait()                                    # every shard validated before anybody iterates
                if any(fl[0] or fl[1] for fl in flags):
                    return
                if comms[r] is not None:
                    eng.comm_attach(comms[r])
                eng.set_dictionary(H_init)
                eng.init_coefficients()                        # W0 = X.H0^T   (nmf.py:156), rows are independent
                if H_loop is not None:
                    eng.set_dictionary(H_loop)
                errors, n_iter = eng.run(max_iter, tol_abs, fit)
                if W is not None and r1 > r0:
                    eng.get_coefficients_into(W[r0:r1])
                results[r] = (errors, n_iter, eng.get_dictionary() if (fit and r == 0) else None)
            except BaseException:
                gate.abort()
                raise
            finally:
                eng.close()

        _in_threads([(lambda r=r: shard(r)) for r in range(world)])
        neg = any(fl[0] for fl in flags)
        bad = any(fl[1] for fl in flags)
        if neg or bad:
            return None, None, None, None, (neg, bad)
        errors, n_iter, H = results[0]
        return W, H, errors, n_iter, (False, False)


class ShardedNMF(object):
    """fit / transform of a row shard under torchrun; every rank calls the same methods in the same order."""

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
            eng.comm_attach(rank_comm(self.device))
        return eng

    def fit_transform(self, X_local, n_global, H0=None, fit=True, return_errors=False, want_coefficients=True):
        f = X_local.shape[1]
        if H0 is None:
            H0 = self.components_ if not fit else draw_shared_dictionary(self.n_components, f)
        eng = self._engine(X_local, n_global)
        try:
            eng.set_dictionary(H0)
            eng.init_coefficients()
            errors, n_iter = eng.run(self.max_iter, self.tol * n_global * f, fit)
            W = eng.get_coefficients() if want_coefficients else None
            if fit:
                self.components_ = eng.get_dictionary()
        finally:
            eng.close()
        return (W, list(errors)) if return_errors else W

    def fit(self, X_local, n_global, H0=None):
        """Like KLdivNMF.fit (nmf.py:259-273): the dictionary is learnt and kept (`components_`, identical on every
        rank), the coefficients of the shard are thrown away -- and therefore never leave the device."""
        self.fit_transform(X_local, n_global, H0=H0, fit=True, want_coefficients=False)
        return self
