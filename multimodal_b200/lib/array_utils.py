"""Array helpers on the path (reference: multimodal/lib/array_utils.py:5-22)."""
import numpy as np
import scipy.sparse as sp


def safe_hstack(blocks):
    """Modality concatenation; any sparse block makes the whole stack sparse
    (reference array_utils.py:5-9)."""
    if any([sp.issparse(b) for b in blocks]):
        return sp.hstack(blocks)
    else:
        return np.hstack(blocks)


class StackedBlocks(object):
    """safe_hstack([c * m for m, c in zip(blocks, coefs)]) of dense modalities, NOT yet formed: the estimator uploads the
    blocks one by one and scales them on the device (klnmf_set_dense_blocks_host), so that the stacked matrix -- 131 GB
    for three modalities of 4 million samples -- never exists on the host.  Anything else that wants the matrix gets it
    from `toarray()` / `np.asarray(...)`, which is exactly what the reference's stack_data returns (learner.py:53-56)."""

    def __init__(self, blocks, coefs):
        self.blocks = [np.asarray(b) for b in blocks]
        self.coefs = list(coefs)            # kept as given: numpy's promotion rules see np.float32 / Python floats
        assert len(self.blocks) == len(self.coefs) and len(self.blocks) > 0
        n = self.blocks[0].shape[0]
        assert all(b.ndim == 2 and b.shape[0] == n for b in self.blocks)
        self.shape = (n, sum(b.shape[1] for b in self.blocks))
        self.ndim = 2
        self.dtype = np.result_type(np.float64, *[b.dtype for b in self.blocks]) \
            if any(b.dtype != np.float32 for b in self.blocks) else np.dtype(np.float32)

    def toarray(self):
        return np.hstack([c * b for b, c in zip(self.blocks, self.coefs)])

    def __array__(self, dtype=None, copy=None):
        a = self.toarray()
        return a if dtype is None else a.astype(dtype)


class MixedBlocks(object):
    """safe_hstack([c * m ...]) of modalities of which at least one is sparse, NOT yet formed.  The reference makes the
    whole stack sparse on the host (scipy.sparse.hstack, array_utils.py:5-9: the dense modalities are sparsified, every
    block scaled, everything concatenated and converted to CSR); the estimator instead uploads the blocks as they are
    and builds the scaled, stacked CSR matrix on the device (klnmf_set_stacked_blocks_host).  `tocsr()` is exactly what
    the reference's stack_data returns, for anything else that wants the matrix."""

    def __init__(self, blocks, coefs):
        self.blocks = [b if sp.issparse(b) else np.asarray(b) for b in blocks]
        self.coefs = list(coefs)
        assert len(self.blocks) == len(self.coefs) and any(sp.issparse(b) for b in self.blocks)
        n = self.blocks[0].shape[0]
        assert all(b.ndim == 2 and b.shape[0] == n for b in self.blocks)
        self.shape = (n, sum(b.shape[1] for b in self.blocks))
        self.ndim = 2

    def tocsr(self):
        return sp.hstack([c * b for b, c in zip(self.blocks, self.coefs)]).tocsr()

    def toarray(self):
        return self.tocsr().toarray()

    def canonical(self):
        """The same stack with every sparse block as canonical CSR (sorted indices, duplicates summed), on copies: the
        reference's eliminate_zeros() (nmf.py:66) touches the temporary stack, never the caller's modality matrices."""
        out = []
        for b in self.blocks:
            if sp.issparse(b):
                b = b.tocsr()
                if not b.has_canonical_format:
                    b = b.copy()
                    b.sum_duplicates()
            out.append(b)
        return MixedBlocks(out, self.coefs)


def normalize_sum(a, axis=0, eps=1.e-16):
    """a / (eps + sum(a, axis)) (reference array_utils.py:19-22)."""
    if axis >= len(a.shape):
        raise ValueError
    return a / (eps + np.expand_dims(np.sum(a, axis=axis), axis))
