#!/usr/bin/env python
"""Host->device upload rates of the public data entry points (pinned host memory): one contiguous array
(klnmf_set_dense_host) against the learner's three modality blocks (klnmf_set_dense_blocks_host: column slices of one
array, and three separate arrays)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_b200 import _native  # noqa: E402

n, f, k = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 8192, 64
X = torch.empty((n, f), dtype=torch.float32, pin_memory=True).numpy()
X[:] = 0.5
dims = [4096, 3072, 1024]
offs = [0, 4096, 7168, 8192]
seps = [torch.empty((n, d), dtype=torch.float32, pin_memory=True).numpy() for d in dims]
for s in seps:
    s[:] = 0.5
pageable = np.full((n, f), 0.5, dtype=np.float32)
gb = X.nbytes / 1e9
for rep in range(2):
    for name, fn in [
        ("set_dense contiguous pinned", lambda e: e.set_dense(X)),
        ("set_dense_blocks column slices", lambda e: e.set_dense_blocks([X[:, offs[i]:offs[i + 1]] for i in range(3)], [1.0, 0.5, 2.0])),
        ("set_dense_blocks separate arrays", lambda e: e.set_dense_blocks(seps, [1.0, 0.5, 2.0])),
        ("set_dense pageable", lambda e: e.set_dense(pageable)),
    ]:
        with _native.Engine(n, f, k, mode="tf32") as e:
            t0 = time.perf_counter()
            fn(e)
            e.check_input()                  # synchronises the stream
            dt = time.perf_counter() - t0
        print("rep %d %-36s %.2f GB in %.3f s = %.1f GB/s" % (rep, name, gb, dt, gb / dt), flush=True)
