# encoding: utf-8
"""Learning from multiple modalities with one KL-NMF dictionary, B200-native.

Drop-in for `multimodal/learner.py` of omangin/multimodal (94 lines there): the same
functions, class, method names, argument meaning, assertions and state (`dico`,
`nmf_train`); the NMF fit / transform and the reconstruction product run on the GPU
through `multimodal_b200.lib.nmf.KLdivNMF` and libklnmf.
"""
import numpy as np

from . import _native
from .lib.nmf import KLdivNMF as NMF
from .lib.array_utils import safe_hstack, StackedBlocks


def fit_coefficients(data_obs, dictionary, iter_nmf=100, verbose=False, mode=None, device=0):
    """Coefficients of `data_obs` on a fixed dictionary (learner.py:11-15): tol=0, so exactly
    `iter_nmf` transform iterations unless the objective rises."""
    nmf_obs = NMF(n_components=dictionary.shape[0], max_iter=iter_nmf, tol=0, mode=mode, device=device)
    nmf_obs.components_ = dictionary
    coefficients = nmf_obs.transform(data_obs, scale_W=True)
    return coefficients


class MultimodalLearner(object):

    def __init__(self, modalities, dimensions, coefficients, k,
                 sparseness=None, sp_coef=.1, mode=None, device=0):
        self.mod = modalities  # Names of the modalities
        self.dim = dimensions  # Dimensions of modalities
        self.coef = coefficients  # Coefficients used to compensate between modalities
        self.k = k
        self.sparseness = sparseness  # data, components, None
        self.sp_coef = sp_coef
        self.dico = None  # None means not trained yet
        self.mode = mode
        self.device = device

    def train(self, data_matrices, iterations):
        """learner.py:31-41."""
        n_samples = data_matrices[0].shape[0]
        for m, d in zip(data_matrices, self.dim):
            assert(m.shape == (n_samples, d))
        Vtrain = self.stack_data(self.mod, data_matrices)
        if self.sparseness is not None:
            raise NotImplemented          # noqa: F901  (the reference raises exactly this)
        self.nmf_train = NMF(n_components=self.k, max_iter=iterations, tol=0,
                             mode=self.mode, device=self.device)
        self.nmf_train.fit(Vtrain, scale_W=True)
        self.dico = self.nmf_train.components_

    def get_dico(self, modality=None):
        if modality is None:
            return self.dico
        else:
            start, stop = self.get_axis_range(modality)
            return self.dico[:, start:stop]

    def get_stacked_dicos(self, modalities):
        return safe_hstack([self.get_dico(modality=m) for m in modalities])

    def stack_data(self, modalities, data_matrices):
        """Scaled concatenation (learner.py:53-56); one sparse block makes the stack sparse."""
        coefs = [self.coef[self.get_index(mod)] for mod in modalities]
        if len(data_matrices) > 1 and all(isinstance(m, np.ndarray) and m.ndim == 2 for m in data_matrices):
            # all modalities dense: the scaled concatenation is formed on the device, block by block
            return StackedBlocks(data_matrices, coefs)
        return safe_hstack([c * m
                            for m, c in zip(data_matrices, coefs)])

    def get_axis_range(self, modality):
        idx = self.get_index(modality)
        start = sum(self.dim[:idx])
        stop = start + self.dim[idx]
        return (start, stop)

    def get_index(self, modality):
        return self.mod.index(modality)

    def reconstruct_internal(self, orig_mod, test_data, iterations):
        return self.reconstruct_internal_multi([orig_mod], [test_data],
                                               iterations)

    def reconstruct_internal_multi(self, orig_mods, test_data, iterations):
        """learner.py:71-78."""
        for mod, data in zip(orig_mods, test_data):
            assert(data.shape[1] == self.dim[self.get_index(mod)])
        stacked_dico = self.get_stacked_dicos(orig_mods)
        stacked_data = self.stack_data(orig_mods, test_data)
        internal = fit_coefficients(stacked_data, stacked_dico,
                                    iter_nmf=iterations, mode=self.mode, device=self.device)
        return internal

    def _dot(self, internal, dico):
        """internal.dot(dico) on the device (learner.py:81,84)."""
        return _native.contract(np.asarray(internal, dtype=np.float64), np.asarray(dico, dtype=np.float64),
                                self.mode, device=self.device)

    def reconstruct_modality(self, dest_mod, internal):
        return self._dot(internal, self.get_dico(dest_mod))

    def reconstruct_modalities(self, dest_mods, internal):
        return self._dot(internal, self.get_stacked_dicos(dest_mods))

    def modality_to_modality(self, orig_mod, dest_mod, test_data, iterations):
        return self.modalities_to_modalities([orig_mod], [dest_mod],
                                             [test_data], iterations)

    def modalities_to_modalities(self, orig_mods, dest_mods, test_data,
                                 iterations):
        internal = self.reconstruct_internal_multi(orig_mods, test_data,
                                                   iterations)
        return self.reconstruct_modalities(dest_mods, internal)
