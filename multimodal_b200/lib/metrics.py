"""Host-side objective helper with the reference's name and meaning
(reference: multimodal/lib/metrics.py:15-20).  It is a small-array utility for callers and
tests; the estimator's own objective (`KLdivNMF.error`) is computed on the GPU."""
import numpy as np

EPSILON = 1.e-8


def generalized_KL(x, y, eps=EPSILON, axis=None):
    return (np.multiply(x, np.log(np.divide(x + eps, y + eps))) - x + y).sum(axis=axis)
