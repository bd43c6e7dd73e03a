"""ctypes binding of libklnmf.so (include/klnmf.h) -- the only way into the CUDA engine.

There is no CPU implementation behind this module: if the shared library is missing,
or no B200 is visible, every entry point raises.  numpy <-> device copies are done by
the library's own `*_host` entry points; torch is only used by callers that already
hold device tensors (bench) or need torch.distributed for the NCCL bootstrap.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libklnmf.so")

MODE_TF32, MODE_TF32X3, MODE_FP64 = 0, 1, 2
MODE_TF32R = 3
MODES = {"tf32": MODE_TF32, "tf32x3": MODE_TF32X3, "fp64": MODE_FP64, "tf32r": MODE_TF32R}
# The one default of the library, the tests and bench.py: one tcgen05 pass on round-to-nearest TF32 copies of W, H and
# the centered ratio, FP32 state, cancellation-free FP64-accumulated objective.
DEFAULT_MODE = "tf32r"
ABI_VERSION = 3
F32, F64 = 0, 1

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_vp = ctypes.c_void_p
_c_dbl = ctypes.c_double

class Block(ctypes.Structure):
    """klnmf_block (include/klnmf.h): one modality of a stack, dense or CSR, with its coefficient."""
    _fields_ = [("kind", _c_int), ("dtype", _c_int), ("cols", _c_i64), ("scale", _c_dbl), ("product_f32", _c_int),
                ("dense", _c_vp), ("ld", _c_i64), ("indptr", _c_vp), ("indices", _c_vp), ("values", _c_vp), ("nnz", _c_i64)]


# name -> (restype, argtypes); must list every symbol include/klnmf.h declares
PROTOTYPES = {
    "klnmf_abi_version": (_c_int, []),
    "klnmf_last_error": (ctypes.c_char_p, []),
    "klnmf_device_count": (_c_int, []),
    "klnmf_create": (_c_int, [ctypes.POINTER(_c_vp), _c_int, _c_i64, _c_i64, _c_i64, _c_int]),
    "klnmf_destroy": (_c_int, [_c_vp]),
    "klnmf_set_stream": (_c_int, [_c_vp, _c_vp]),
    "klnmf_set_scratch_limit": (_c_int, [_c_vp, _c_i64]),
    "klnmf_set_dense_host": (_c_int, [_c_vp, _c_vp, _c_int, _c_i64]),
    "klnmf_set_dense_device": (_c_int, [_c_vp, _c_vp, _c_int, _c_i64]),
    "klnmf_set_dense_blocks_host": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "klnmf_set_stacked_blocks_host": (_c_int, [_c_vp, _c_int, ctypes.POINTER(Block)]),
    "klnmf_set_hybrid_min_cols": (_c_int, [_c_vp, _c_i64]),
    "klnmf_is_hybrid": (_c_int, [_c_vp]),
    "klnmf_set_csr_host": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_i64]),
    "klnmf_set_csr_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_i64]),
    "klnmf_create_column_view": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_i64, ctypes.POINTER(_c_vp)]),
    "klnmf_check_input": (_c_int, [_c_vp, ctypes.POINTER(ctypes.c_int32)]),
    "klnmf_set_dictionary_host": (_c_int, [_c_vp, _c_vp, _c_i64]),
    "klnmf_get_dictionary_host": (_c_int, [_c_vp, _c_vp, _c_i64]),
    "klnmf_init_coefficients": (_c_int, [_c_vp]),
    "klnmf_set_coefficients_host": (_c_int, [_c_vp, _c_vp, _c_i64]),
    "klnmf_get_coefficients_host": (_c_int, [_c_vp, _c_vp, _c_i64]),
    "klnmf_coefficients_device": (_c_int, [_c_vp, ctypes.POINTER(_c_vp), ctypes.POINTER(_c_i64),
                                           ctypes.POINTER(_c_int)]),
    "klnmf_dictionary_device": (_c_int, [_c_vp, ctypes.POINTER(_c_vp), ctypes.POINTER(_c_i64),
                                         ctypes.POINTER(_c_int)]),
    "klnmf_run": (_c_int, [_c_vp, _c_int, _c_dbl, _c_int, _c_vp, ctypes.POINTER(_c_int),
                           ctypes.POINTER(_c_int)]),
    "klnmf_run_resume": (_c_int, [_c_vp, _c_int, _c_dbl, _c_int, _c_dbl, _c_vp, ctypes.POINTER(_c_int),
                                  ctypes.POINTER(_c_int)]),
    "klnmf_error": (_c_int, [_c_vp, ctypes.POINTER(_c_dbl)]),
    "klnmf_dictionary_step": (_c_int, [_c_vp]),
    "klnmf_ratio_host": (_c_int, [_c_vp, _c_vp, _c_int, _c_i64]),
    "klnmf_sddmm_host": (_c_int, [_c_vp, _c_vp]),
    "klnmf_reconstruct_host": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64]),
    "klnmf_nccl_load": (_c_int, [ctypes.c_char_p]),
    "klnmf_nccl_unique_id": (_c_int, [_c_vp]),
    "klnmf_comm_init": (_c_int, [_c_vp, _c_vp, _c_int, _c_int]),
    "klnmf_comm_create": (_c_int, [ctypes.POINTER(_c_vp), _c_int, _c_vp, _c_int, _c_int]),
    "klnmf_comm_attach": (_c_int, [_c_vp, _c_vp, _c_int, _c_int]),
    "klnmf_comm_destroy": (_c_int, [_c_vp]),
    "klnmf_fill_dense_synthetic": (_c_int, [_c_vp, ctypes.c_uint64]),
    "klnmf_fill_csr_synthetic": (_c_int, [_c_vp, _c_i64, ctypes.c_uint64]),
    "klnmf_get_dense_host": (_c_int, [_c_vp, _c_vp, _c_int, _c_i64]),
    "klnmf_counters": (_c_int, [_c_vp, ctypes.POINTER(_c_i64)]),
    "klnmf_last_run_profile": (_c_int, [_c_vp, ctypes.POINTER(_c_dbl), ctypes.POINTER(_c_i64)]),
    "klnmf_contract_host": (_c_int, [_c_int, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_int, _c_vp, _c_int, _c_vp]),
    "klnmf_contract_bench": (_c_int, [_c_int, _c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_int,
                                      ctypes.POINTER(_c_dbl)]),
    "klnmf_l2_read_bench": (_c_int, [_c_int, _c_i64, _c_int, ctypes.POINTER(_c_dbl)]),
    "klnmf_engine_name": (ctypes.c_char_p, [_c_vp]),
    "klnmf_pairwise_host": (_c_int, [_c_int, _c_int, _c_i64, _c_i64, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64,
                                     _c_vp, _c_vp]),
}

_lib = None


class KlnmfError(RuntimeError):
    pass


def load():
    """dlopen libklnmf.so.  With nvcc on the machine the library is (re)built first whenever it is missing or older
    than any source or header (build.is_stale()); without nvcc (the GPU box) the shipped binary is used as it is, and
    the ABI version below catches a header / binary mismatch."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.have_nvcc():
        _build.build()
    elif not os.path.exists(LIB_PATH):
        raise KlnmfError("libklnmf.so is missing and there is no nvcc to build it: " + LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.klnmf_abi_version() != ABI_VERSION:
        raise KlnmfError("libklnmf ABI mismatch: the binary says %d, this binding %d -- rebuild with "
                         "`python -m multimodal_b200.build --force`" % (lib.klnmf_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def device_count():
    return load().klnmf_device_count()


def _check(rc):
    if rc != 0:
        msg = load().klnmf_last_error().decode("utf-8", "replace")
        if rc == -3:
            raise KlnmfError("libklnmf needs a B200 (sm_100a) device and has no CPU path: " + msg)
        if rc == -1:
            raise ValueError(msg)
        if rc == -6:
            raise MemoryError(msg)
        raise KlnmfError("libklnmf error %d: %s" % (rc, msg))


def resolve_mode(mode):
    if mode is None:
        mode = os.environ.get("KLNMF_MODE", DEFAULT_MODE)
    if isinstance(mode, str):
        if mode not in MODES:
            raise ValueError("unknown arithmetic mode %r (expected one of %s)" % (mode, sorted(MODES)))
        return MODES[mode]
    return int(mode)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_c_vp)


class Engine(object):
    """One klnmf context: a shard of samples on one GPU."""

    def __init__(self, n, f, k, mode=None, device=0, scratch_limit=None):
        self.lib = load()
        self.n, self.f, self.k = int(n), int(f), int(k)
        self.mode = resolve_mode(mode)
        h = _c_vp()
        _check(self.lib.klnmf_create(ctypes.byref(h), int(device), self.n, self.f, self.k, self.mode))
        self.h = h
        self._keep = []          # host/device buffers the context borrows
        if scratch_limit:
            _check(self.lib.klnmf_set_scratch_limit(self.h, int(scratch_limit)))

    # -- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.klnmf_destroy(self.h)
            self.h = None
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def engine_name(self):
        return self.lib.klnmf_engine_name(self.h).decode()

    # -- data -------------------------------------------------------------------------
    def set_dense(self, X):
        """X: C-contiguous float32/float64 ndarray (n, f) on the host."""
        if X.dtype == np.float32:
            dt = F32
        else:
            X = np.asarray(X, dtype=np.float64)
            dt = F64
        if not X.flags.c_contiguous:
            X = np.ascontiguousarray(X)
        assert X.shape == (self.n, self.f)
        _check(self.lib.klnmf_set_dense_host(self.h, _ptr(X), dt, X.strides[0] // X.itemsize if self.n > 1 else self.f))

    def set_dense_blocks(self, blocks, scales):
        """safe_hstack([c * m]) of learner.stack_data (learner.py:53-56) formed on the device: `blocks` are host
        ndarrays (n, f_b), `scales` the modality coefficients; no stacked host copy is made."""
        mats, f32 = [], []
        for m, c in zip(blocks, scales):
            m = np.asarray(m)
            # numpy's promotion of `c * m` decides in which precision the reference forms the product
            f32.append(1 if (c * m[:0, :0]).dtype == np.float32 else 0)
            if m.dtype not in (np.float32, np.float64):
                m = m.astype(np.float64)
            assert m.ndim == 2 and m.shape[0] == self.n, (m.shape, self.n)
            if m.strides[1] != m.itemsize or m.strides[0] < m.shape[1] * m.itemsize:
                m = np.ascontiguousarray(m)          # column slices of one array (row pitch > width) go as they are
            mats.append(m)
        nb = len(mats)
        assert sum(m.shape[1] for m in mats) == self.f
        ptrs = (ctypes.c_void_p * nb)(*[m.ctypes.data for m in mats])
        dts = (ctypes.c_int * nb)(*[F32 if m.dtype == np.float32 else F64 for m in mats])
        lds = (ctypes.c_int64 * nb)(*[(m.strides[0] // m.itemsize) if self.n > 1 else m.shape[1] for m in mats])
        cols = (ctypes.c_int64 * nb)(*[m.shape[1] for m in mats])
        scl = (ctypes.c_double * nb)(*[float(s) for s in scales])
        pf = (ctypes.c_int * nb)(*f32)
        _check(self.lib.klnmf_set_dense_blocks_host(self.h, nb, ptrs, dts, lds, cols, scl, pf))

    def set_hybrid_min_cols(self, cols):
        """Dense columns from which a mixed stack keeps its dense blocks dense (klnmf_set_hybrid_min_cols; default 1024,
        0 = never).  Call before set_stacked_blocks."""
        _check(self.lib.klnmf_set_hybrid_min_cols(self.h, int(cols)))

    def is_hybrid(self):
        return bool(self.lib.klnmf_is_hybrid(self.h))

    def set_stacked_blocks(self, blocks, scales):
        """safe_hstack([c * m]) of learner.stack_data (learner.py:53-56) when a modality is sparse -- the reference then
        makes the whole stack sparse on the host (array_utils.py:5-9).  `blocks` are host ndarrays and canonical scipy CSR
        matrices (n, f_b); they are uploaded as they are and the scaled, stacked CSR matrix is built on the device
        (klnmf_set_stacked_blocks_host) -- or, when the dense blocks are wide enough (set_hybrid_min_cols), only the CSR
        blocks are stacked and the dense ones stay dense (hybrid stack: tcgen05 contractions next to the sparse passes)."""
        import scipy.sparse as sp
        arr = (Block * len(blocks))()
        keep = []
        for i, (m, c) in enumerate(zip(blocks, scales)):
            b = arr[i]
            b.scale = float(c)
            b.cols = int(m.shape[1])
            assert m.shape[0] == self.n, (m.shape, self.n)
            if sp.issparse(m):
                m = m.tocsr()
                # numpy's promotion of `c * m` decides in which precision the reference forms the product
                b.product_f32 = 1 if (c * m.data[:0]).dtype == np.float32 else 0
                indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
                indices = np.ascontiguousarray(m.indices, dtype=np.int32)
                data = np.ascontiguousarray(m.data) if m.data.dtype == np.float32 else np.ascontiguousarray(m.data, dtype=np.float64)
                keep += [indptr, indices, data]
                b.kind, b.dtype, b.nnz = 1, (F32 if data.dtype == np.float32 else F64), int(m.nnz)
                b.indptr, b.indices, b.values = indptr.ctypes.data, indices.ctypes.data, data.ctypes.data
            else:
                m = np.asarray(m)
                b.product_f32 = 1 if (c * m[:0, :0]).dtype == np.float32 else 0
                if m.dtype not in (np.float32, np.float64):
                    m = m.astype(np.float64)
                if m.strides[1] != m.itemsize or m.strides[0] < m.shape[1] * m.itemsize:
                    m = np.ascontiguousarray(m)
                keep.append(m)
                b.kind, b.dtype = 0, (F32 if m.dtype == np.float32 else F64)
                b.dense, b.ld = m.ctypes.data, ((m.strides[0] // m.itemsize) if self.n > 1 else m.shape[1])
        assert sum(int(b.cols) for b in arr) == self.f
        _check(self.lib.klnmf_set_stacked_blocks_host(self.h, len(blocks), arr))
        del keep

    def set_dense_device(self, ptr, dtype, ld, keepalive=None):
        self._keep.append(keepalive)
        _check(self.lib.klnmf_set_dense_device(self.h, _c_vp(ptr), dtype, int(ld)))

    def set_csr(self, X):
        """X: scipy CSR (n, f), canonical (no duplicates), explicit zeros removed."""
        indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(X.indices, dtype=np.int32)
        if X.data.dtype == np.float32:
            data, dt = np.ascontiguousarray(X.data), F32
        else:
            data, dt = np.ascontiguousarray(X.data, dtype=np.float64), F64
        assert X.shape == (self.n, self.f)
        _check(self.lib.klnmf_set_csr_host(self.h, _ptr(indptr), _ptr(indices), _ptr(data), dt, int(X.nnz)))

    def column_view(self, ranges, k):
        """A new Engine over the columns [start, stop) of every (start, stop) in `ranges`, concatenated in that order
        and gathered on the device from this engine's dense data (klnmf_create_column_view)."""
        starts = (ctypes.c_int64 * len(ranges))(*[int(a) for a, _ in ranges])
        widths = (ctypes.c_int64 * len(ranges))(*[int(b) - int(a) for a, b in ranges])
        h = _c_vp()
        _check(self.lib.klnmf_create_column_view(self.h, len(ranges), starts, widths, int(k), ctypes.byref(h)))
        child = Engine.__new__(Engine)
        child.lib, child.n, child.f, child.k = self.lib, self.n, sum(int(b) - int(a) for a, b in ranges), int(k)
        child.mode, child.h, child._keep = self.mode, h, []
        return child

    def check_input(self):
        out = (ctypes.c_int32 * 2)()
        _check(self.lib.klnmf_check_input(self.h, out))
        return bool(out[0]), bool(out[1])

    def fill_dense_synthetic(self, seed):
        _check(self.lib.klnmf_fill_dense_synthetic(self.h, int(seed)))

    def fill_csr_synthetic(self, nnz_per_row, seed):
        _check(self.lib.klnmf_fill_csr_synthetic(self.h, int(nnz_per_row), int(seed)))

    def get_dense(self, out):
        dt = F32 if out.dtype == np.float32 else F64
        _check(self.lib.klnmf_get_dense_host(self.h, _ptr(out), dt, out.strides[0] // out.itemsize))
        return out

    # -- state ---------------------------------------------------------------------------
    def set_dictionary(self, H):
        H = _as_f64(H)
        assert H.shape == (self.k, self.f), (H.shape, (self.k, self.f))
        _check(self.lib.klnmf_set_dictionary_host(self.h, _ptr(H), self.f))

    def get_dictionary(self):
        H = np.empty((self.k, self.f), dtype=np.float64)
        _check(self.lib.klnmf_get_dictionary_host(self.h, _ptr(H), self.f))
        return H

    def init_coefficients(self):
        _check(self.lib.klnmf_init_coefficients(self.h))

    def set_coefficients(self, W):
        W = _as_f64(W)
        assert W.shape == (self.n, self.k), (W.shape, (self.n, self.k))
        _check(self.lib.klnmf_set_coefficients_host(self.h, _ptr(W), self.k))

    def get_coefficients(self):
        W = np.empty((self.n, self.k), dtype=np.float64)
        _check(self.lib.klnmf_get_coefficients_host(self.h, _ptr(W), self.k))
        return W

    def get_coefficients_into(self, out):
        """The coefficients written into a caller's C-contiguous float64 (n, k) array (e.g. a row block of W)."""
        assert out.dtype == np.float64 and out.shape == (self.n, self.k) and out.flags.c_contiguous
        _check(self.lib.klnmf_get_coefficients_host(self.h, _ptr(out), self.k))
        return out

    # -- compute ----------------------------------------------------------------------------
    def run(self, max_iter, tol_abs, fit, prev_objective=float("inf")):
        """Returns (errors ndarray, n_iter) with the reference's meaning of both.  `prev_objective`: the last objective
        recorded by an earlier call on this engine when a fit is run in several pieces (klnmf_run_resume)."""
        errs = np.empty(max(int(max_iter), 1), dtype=np.float64)
        ne, ni = _c_int(0), _c_int(0)
        _check(self.lib.klnmf_run_resume(self.h, int(max_iter), float(tol_abs), 1 if fit else 0, float(prev_objective),
                                         _ptr(errs), ctypes.byref(ne), ctypes.byref(ni)))
        return errs[:ne.value].copy(), ni.value

    def error(self):
        out = _c_dbl(0.0)
        _check(self.lib.klnmf_error(self.h, ctypes.byref(out)))
        return out.value

    def dictionary_step(self):
        _check(self.lib.klnmf_dictionary_step(self.h))

    def ratio(self, nnz=None):
        if nnz is None:
            out = np.empty((self.n, self.f), dtype=np.float64)
            _check(self.lib.klnmf_ratio_host(self.h, _ptr(out), F64, self.f))
        else:
            out = np.empty(int(nnz), dtype=np.float64)
            _check(self.lib.klnmf_ratio_host(self.h, _ptr(out), F64, int(nnz)))
        return out

    def sddmm(self, nnz):
        out = np.empty(int(nnz), dtype=np.float64)
        _check(self.lib.klnmf_sddmm_host(self.h, _ptr(out)))
        return out

    def reconstruct(self, H_dest):
        H_dest = _as_f64(H_dest)
        assert H_dest.shape[0] == self.k
        fd = H_dest.shape[1]
        out = np.empty((self.n, fd), dtype=np.float64)
        _check(self.lib.klnmf_reconstruct_host(self.h, _ptr(H_dest), fd, fd, _ptr(out), fd))
        return out

    # -- multi GPU -----------------------------------------------------------------------------
    def comm_init(self, uid, rank, world):
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(uid))
        _check(self.lib.klnmf_comm_init(self.h, ctypes.cast(buf, _c_vp), int(rank), int(world)))

    def comm_attach(self, comm):
        """Lend a cached communicator (see `Comm`) to this context."""
        _check(self.lib.klnmf_comm_attach(self.h, comm.h, comm.rank, comm.world))
        self._keep.append(comm)

    # -- accounting ------------------------------------------------------------------------------
    def counters(self):
        out = (_c_i64 * 4)()
        _check(self.lib.klnmf_counters(self.h, out))
        return {"launches": out[0], "nccl_calls": out[1], "h2d_bytes": out[2], "d2h_bytes": out[3]}

    def last_run_profile(self):
        ms = (_c_dbl * 6)()
        cnt = (_c_i64 * 5)()
        _check(self.lib.klnmf_last_run_profile(self.h, ms, cnt))
        names = ["ratio", "coefficient", "numerator", "dictionary", "allreduce", "total"]
        return ({n: ms[i] for i, n in enumerate(names)}, {n: cnt[i] for i, n in enumerate(names[:5])})


class Comm(object):
    """An NCCL communicator that outlives the engines it is lent to (klnmf_comm_create): creating one is collective
    and costs 0.2-1 s, an Engine is created per fit / transform call."""

    def __init__(self, device, uid, rank, world):
        self.lib = load()
        nccl_load()
        self.device, self.rank, self.world = int(device), int(rank), int(world)
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(uid))
        h = _c_vp()
        _check(self.lib.klnmf_comm_create(ctypes.byref(h), self.device, ctypes.cast(buf, _c_vp), self.rank, self.world))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.klnmf_comm_destroy(self.h)
            self.h = None


def nccl_unique_id(libnccl_path=None):
    lib = load()
    if libnccl_path is None:
        libnccl_path = find_libnccl()
    _check(lib.klnmf_nccl_load(libnccl_path.encode() if libnccl_path else None))
    buf = (ctypes.c_char * 128)()
    _check(lib.klnmf_nccl_unique_id(ctypes.cast(buf, _c_vp)))
    return bytes(buf)


def nccl_load(libnccl_path=None):
    lib = load()
    if libnccl_path is None:
        libnccl_path = find_libnccl()
    _check(lib.klnmf_nccl_load(libnccl_path.encode() if libnccl_path else None))


def find_libnccl():
    """The NCCL that ships with torch (nvidia-nccl-cu12 wheel), else the system one."""
    try:
        import nvidia.nccl
        for base in list(nvidia.nccl.__path__):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def contract(A, B, mode, a_trans=False, b_trans=False, device=0):
    """Diagnostic: op(A).op(B) through the dense engine of `mode` (see klnmf_contract_host)."""
    lib = load()
    A, B = _as_f64(A), _as_f64(B)
    M, K = (A.shape[1], A.shape[0]) if a_trans else A.shape
    N = B.shape[0] if b_trans else B.shape[1]
    assert (B.shape[1] if b_trans else B.shape[0]) == K
    out = np.empty((M, N), dtype=np.float64)
    _check(lib.klnmf_contract_host(int(device), resolve_mode(mode), M, N, K, _ptr(A), 1 if a_trans else 0,
                                   _ptr(B), 1 if b_trans else 0, _ptr(out)))
    return out


def contract_bench(M, N, K, mode, a_trans=False, b_trans=False, iters=10, device=0):
    """Diagnostic: average device ms of one M x N x K contraction (see klnmf_contract_bench)."""
    lib = load()
    ms = _c_dbl(0.0)
    _check(lib.klnmf_contract_bench(int(device), resolve_mode(mode), int(M), int(N), int(K), 1 if a_trans else 0,
                                    1 if b_trans else 0, int(iters), ctypes.byref(ms)))
    return ms.value


def l2_read_bandwidth(bytes=48 << 20, iters=200, device=0):
    """Diagnostic: sustained L2 -> SM read bandwidth in GB/s (klnmf_l2_read_bench)."""
    lib = load()
    out = _c_dbl(0.0)
    _check(lib.klnmf_l2_read_bench(int(device), int(bytes), int(iters), ctypes.byref(out)))
    return out.value


MEASURES = {"kl_div": 0, "rev_kl_div": 1, "sym_kl_div": 2, "frobenius": 3, "cosine_diff": 4}


def pairwise(A, B, measure, want_dists=True, want_argmin=False, device=0):
    """Distances of every row of A (n_test, d) to every row of B (n_ex, d) under `measure` (a key of MEASURES)
    on the device (csrc/evaluation.cu); returns (dists or None, argmin or None)."""
    lib = load()
    A, B = _as_f64(np.atleast_2d(A)), _as_f64(np.atleast_2d(B))
    if A.shape[1] != B.shape[1]:
        raise ValueError("operands could not be broadcast together with shapes %s %s" % (A.shape, B.shape))
    nt, ne, d = A.shape[0], B.shape[0], A.shape[1]
    D = np.empty((nt, ne), dtype=np.float64) if want_dists else None
    idx = np.empty((nt,), dtype=np.int32) if want_argmin else None
    if want_argmin and ne == 0:
        raise ValueError("attempt to get argmin of an empty sequence")
    _check(lib.klnmf_pairwise_host(int(device), MEASURES[measure], nt, ne, d, _ptr(A), max(d, 1), _ptr(B), max(d, 1),
                                   _ptr(D) if D is not None else None, max(ne, 1),
                                   _ptr(idx) if idx is not None else None, None))
    return D, idx
