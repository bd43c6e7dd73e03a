# encoding: utf-8
"""One KL-NMF dictionary over several modalities, B200-native.

Drop-in for `multimodal/learner.py` of omangin/multimodal: `fit_coefficients` and `MultimodalLearner` keep the
reference's names, argument meaning, assertions and state (`mod`, `dim`, `coef`, `k`, `dico`, `nmf_train`), so callers
(`experiment.py:154-180, 274, 365`, the samples) work unchanged.  The arithmetic is not here: fitting, inferring
coefficients and the reconstruction product run on the GPU through `KLdivNMF` / libklnmf, and a stack of dense
modalities is scaled and concatenated on the device (`StackedBlocks`), never on the host.

Column layout of the dictionary (learner.py:43-65): modality i owns columns [sum(dim[:i]), sum(dim[:i]) + dim[i]).
"""
import numpy as np

from . import _native
import scipy.sparse as sp

from .lib.array_utils import MixedBlocks, StackedBlocks, safe_hstack
from .lib.nmf import KLdivNMF as NMF


def fit_coefficients(data_obs, dictionary, iter_nmf=100, verbose=False, mode=None, device=0):
    """Coefficients of `data_obs` on a FIXED dictionary (learner.py:11-15).  tol=0 means: exactly `iter_nmf`
    transform iterations unless the objective rises; `scale_W` is passed and ignored, as in the reference."""
    estimator = NMF(n_components=dictionary.shape[0], max_iter=iter_nmf, tol=0, mode=mode, device=device)
    estimator.components_ = dictionary
    return estimator.transform(data_obs, scale_W=True)


class MultimodalLearner(object):
    """Joint dictionary of several modalities (learner.py:18-94)."""

    def __init__(self, modalities, dimensions, coefficients, k,
                 sparseness=None, sp_coef=.1, mode=None, device=0):
        self.mod = modalities            # names
        self.dim = dimensions            # feature counts
        self.coef = coefficients         # weights that balance the modalities in the stack
        self.k = k
        self.sparseness = sparseness     # 'data' / 'components' / None in the reference; only None is implemented there
        self.sp_coef = sp_coef
        self.dico = None                 # k x sum(dim) once trained (or assigned by the caller)
        self.mode = mode                 # arithmetic mode of libklnmf (not in the reference)
        self.device = device             # CUDA ordinal, or a list of ordinals: the samples are sharded over them

    # ---- layout ---------------------------------------------------------------------------------------------
    def get_index(self, modality):
        return self.mod.index(modality)

    def get_axis_range(self, modality):
        """(start, stop) of the modality's columns in the stacked data / dictionary (learner.py:58-62)."""
        position = self.get_index(modality)
        first = int(np.sum(self.dim[:position], dtype=np.int64)) if position else 0
        return (first, first + self.dim[position])

    def get_dico(self, modality=None):
        if modality is None:
            return self.dico
        first, last = self.get_axis_range(modality)
        return self.dico[:, first:last]

    def get_stacked_dicos(self, modalities):
        return safe_hstack([self.get_dico(modality=name) for name in modalities])

    def stack_data(self, modalities, data_matrices):
        """coef-weighted concatenation in the order of `modalities` (learner.py:53-56).  A sparse block makes the
        whole stack sparse (array_utils.py:5-9).  Several blocks stay apart until the device has them: dense ones are
        scaled and concatenated there (StackedBlocks), a mix of dense and sparse ones is turned into the scaled, stacked
        CSR matrix there (MixedBlocks) -- no sparsification, scaling or stacking pass on the host."""
        weights = [self.coef[self.get_index(name)] for name in modalities]
        dense = [isinstance(block, np.ndarray) and block.ndim == 2 for block in data_matrices]
        if all(dense) and len(data_matrices) > 1:
            return StackedBlocks(data_matrices, weights)
        if len(data_matrices) > 1 and all(d or sp.issparse(block) for d, block in zip(dense, data_matrices)):
            return MixedBlocks(data_matrices, weights)
        return safe_hstack([w * block for block, w in zip(data_matrices, weights)])

    # ---- training (learner.py:31-41) --------------------------------------------------------------------------
    def train(self, data_matrices, iterations):
        n_samples = data_matrices[0].shape[0]
        for block, width in zip(data_matrices, self.dim):
            assert(block.shape == (n_samples, width))
        stacked = self.stack_data(self.mod, data_matrices)
        if self.sparseness is not None:
            raise NotImplemented          # noqa: F901  (the reference raises exactly this object)
        self.nmf_train = NMF(n_components=self.k, max_iter=iterations, tol=0,
                             mode=self.mode, device=self.device)
        self.nmf_train.fit(stacked, scale_W=True)
        self.dico = self.nmf_train.components_

    # ---- inference (learner.py:67-94) -------------------------------------------------------------------------
    def reconstruct_internal_multi(self, orig_mods, test_data, iterations):
        """Internal coefficients of samples observed through a subset of the modalities."""
        for name, block in zip(orig_mods, test_data):
            assert(block.shape[1] == self.dim[self.get_index(name)])
        return fit_coefficients(self.stack_data(orig_mods, test_data), self.get_stacked_dicos(orig_mods),
                                iter_nmf=iterations, mode=self.mode, device=self.device)

    def reconstruct_internal_batch(self, test_data, subsets, iterations):
        """Internal coefficients of the SAME samples observed through several subsets of the modalities -- what the
        evaluation of a dictionary asks for (experiment.py:350-369: `_get_all_internals` calls
        reconstruct_internal_multi once per subset, six for three modalities, each call re-stacking and re-uploading
        its data).  `test_data` holds one matrix per modality of the learner, in the learner's order; `subsets` is a
        list of lists of modality names.  Returns {tuple(subset): internal}, each entry equal to
        reconstruct_internal_multi(subset, [data of the subset], iterations).

        Dense modalities: the scaled stack of ALL modalities is uploaded and validated once; every subset is a column
        view gathered on the device (klnmf_create_column_view) with its own sub-dictionary.  A sparse modality makes the
        reference's stack sparse (array_utils.py:5-9): those calls go one by one, as before."""
        assert len(test_data) == len(self.mod)
        n_samples = test_data[0].shape[0]
        for block, width in zip(test_data, self.dim):
            assert(block.shape == (n_samples, width))
        subsets = [list(s) for s in subsets]
        if not all(isinstance(block, np.ndarray) and block.ndim == 2 for block in test_data):
            return dict((tuple(s), self.reconstruct_internal_multi(
                s, [test_data[self.get_index(name)] for name in s], iterations)) for s in subsets)
        device = self.device[0] if isinstance(self.device, (list, tuple)) else self.device
        out = {}
        with _native.Engine(n_samples, int(sum(self.dim)), self.k, mode=self.mode, device=device) as parent:
            parent.set_dense_blocks(test_data, self.coef)
            negative, non_finite = parent.check_input()
            if non_finite:
                raise ValueError("array contains NaN or infinity")
            if negative:
                raise ValueError("Negative values in data passed to NMF.fit")
            for s in subsets:
                dictionary = np.ascontiguousarray(self.get_stacked_dicos(s), dtype=np.float64)
                with parent.column_view([self.get_axis_range(name) for name in s], dictionary.shape[0]) as view:
                    view.set_dictionary(dictionary)
                    view.init_coefficients()                      # W0 = X.H^T (nmf.py:156)
                    view.run(iterations, 0.0, False)              # fit_coefficients: tol = 0, transform (learner.py:11-15)
                    out[tuple(s)] = view.get_coefficients()
        return out

    def reconstruct_internal(self, orig_mod, test_data, iterations):
        return self.reconstruct_internal_multi([orig_mod], [test_data], iterations)

    def _dot(self, internal, dico):
        """internal.dot(dico) on the device (learner.py:81, 84)."""
        device = self.device[0] if isinstance(self.device, (list, tuple)) else self.device
        return _native.contract(np.asarray(internal, dtype=np.float64), np.asarray(dico, dtype=np.float64),
                                self.mode, device=device)

    def reconstruct_modalities(self, dest_mods, internal):
        return self._dot(internal, self.get_stacked_dicos(dest_mods))

    def reconstruct_modality(self, dest_mod, internal):
        return self._dot(internal, self.get_dico(dest_mod))

    def modalities_to_modalities(self, orig_mods, dest_mods, test_data, iterations):
        hidden = self.reconstruct_internal_multi(orig_mods, test_data, iterations)
        return self.reconstruct_modalities(dest_mods, hidden)

    def modality_to_modality(self, orig_mod, dest_mod, test_data, iterations):
        return self.modalities_to_modalities([orig_mod], [dest_mod], [test_data], iterations)
