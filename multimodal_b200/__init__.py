"""multimodal_b200 -- B200-native drop-in for ONE hot path of omangin/multimodal:
the KL-divergence NMF multiplicative-update loop (`multimodal/lib/nmf.py`) as driven by
`multimodal/learner.py`.  Same class / method names and semantics; the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/klnmf.h (no CPU fallback).

    from multimodal_b200.lib.nmf import KLdivNMF
    from multimodal_b200.learner import MultimodalLearner, fit_coefficients
"""
__version__ = "0.1.0"
