#!/bin/bash
# Install the UNMODIFIED reference (omangin/multimodal) into baseline/_ref for bench.py's `--impl reference` arm.
# Run once in the build container (the GPU box has no /root/reference; baseline/_ref is git-ignored but travels
# with the gpurun snapshot).  The source tree is read-only, so the build runs on a copy under /tmp; --no-deps
# because setup.py also lists matplotlib / librosa (plotting and feature extraction, not on the NMF path), which
# the image lacks.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${KLNMF_REFERENCE:-/root/reference}
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/src"
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/src"
rm -rf "$TMP"
python - <<PY
import sys, numpy as np
np.Inf = np.inf            # numpy >= 2: nmf.py:206 still says np.Inf
sys.path.insert(0, "$ROOT/baseline/_ref")
from multimodal.lib.nmf import KLdivNMF
from multimodal.learner import MultimodalLearner
import multimodal, os
print("reference installed:", os.path.dirname(multimodal.__file__))
PY
