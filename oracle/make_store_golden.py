"""Write tests/golden/logger_store.{json,npz} with the REAL reference logger (test infrastructure; build container
only):  python oracle/make_store_golden.py   # needs /root/reference

The file holds two experiment runs stored the way multimodal/experiment.py:160-170 does (train / test index lists and
the trained dictionary) plus one global array and one global scalar; tests/test_store.py reads it back with
multimodal_b200.store and checks that the files this package writes are the same."""
import os
import sys

import numpy as np

REF = os.environ.get("KLNMF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multimodal.lib.logger import Logger      # noqa: E402  (the reference)
from oracle import cases                      # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden", "logger_store")
    lg = Logger(filename=out)
    lg.store_global('sample-pairing', np.arange(12).reshape(6, 2))
    lg.store_global('k', 4)
    for run, d in enumerate(cases.store_dictionaries()):
        lg.new_run()
        lg.store('train', [0, 1, 2, 3 + run])
        lg.store('test', [4, 5])
        lg.store('dictionary', d)
    lg.save()
    back = Logger.load(out)
    assert np.array_equal(back.get_last_value('dictionary'), cases.store_dictionaries()[-1])
    print("wrote", out + ".json", out + ".npz")


if __name__ == "__main__":
    main()
