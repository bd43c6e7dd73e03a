"""BASELINE.json configs[0] and configs[1] at their real sizes on the GPU: iterations/s of the public API next to
the float64 oracle port on the host cores (both are launch-latency bound problems)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200.lib.nmf import KLdivNMF
from multimodal_b200.learner import MultimodalLearner
from oracle import cases, klnmf_oracle as O


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t)


X = cases.cfg1_X()
for mode in ("tf32", "tf32x3", "fp64"):
    def run():
        np.random.seed(1)
        KLdivNMF(n_components=10, max_iter=100, tol=0, mode=mode).fit_transform(X)
    print("cfg1 500x200 k=10 100 it  %-7s %8.1f it/s (whole fit_transform call incl. transfers)" % (mode, 100 / timed(run)))
def cpu():
    np.random.seed(1)
    O.fit_transform(X, k=10, max_iter=100, tol=0)
print("cfg1 oracle port (float64 numpy, host)  %8.1f it/s" % (100 / timed(cpu)))

motion, sound, coefs = cases.cfg2_inputs()
for mode in ("tf32", "tf32x3", "fp64"):
    def run():
        np.random.seed(3)
        lr = MultimodalLearner(['motion', 'sound'], [450, 110000], coefs, 50, mode=mode)
        lr.train([motion, sound.copy()], 50)
        return lr
    t = timed(run, 2)
    print("cfg2 learner.train 1000x110450 CSR k=50 50 it  %-7s %8.2f it/s" % (mode, 50 / t))
lr = run()
t = timed(lambda: lr.modality_to_modality('motion', 'sound', motion[:100], 50), 2)
print("cfg2 modality_to_modality(motion->sound, 100 samples, 50 it)  %.3f s" % t)
ref = O.Learner(['motion', 'sound'], [450, 110000], coefs, 50)
t0 = time.perf_counter(); np.random.seed(3); ref.train([motion, sound.copy()], 3); t = time.perf_counter() - t0
print("cfg2 oracle port train (3 it)  %8.2f it/s" % (3 / t))
