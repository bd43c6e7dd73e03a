"""cProfile of MultimodalLearner.train / modality_to_modality at cfg2 (1000 x (450 dense + 110 000 CSR), k = 50, 50 iterations)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
from multimodal_b200.lib.nmf import KLdivNMF
from multimodal_b200.learner import MultimodalLearner
from oracle import cases
mot, snd, coefs = cases.cfg2_inputs()
lr = MultimodalLearner(['motion', 'sound'], [450, 110000], coefs, 50)
np.random.seed(3); lr.train([mot, snd], 50)     # warm
for rep in range(2):
    t0=time.perf_counter(); np.random.seed(3); lr.train([mot, snd], 50); t1=time.perf_counter()
    out = lr.modality_to_modality('motion', 'sound', mot[:100], 50); t2=time.perf_counter()
    print("train 50 it: %.1f ms   motion->sound (100 samples, 50 it): %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3))
pr = cProfile.Profile(); pr.enable(); np.random.seed(3); lr.train([mot, snd], 50); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
