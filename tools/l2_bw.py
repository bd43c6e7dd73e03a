#!/usr/bin/env python
"""L2 -> SM read bandwidth by footprint (klnmf_l2_read_bench: eight 512-byte pieces in flight per warp, the access shape of
the sparse path's gathers) -- the denominator of bench.py's `roofline.l2_roofline`.  B200: 19.4 TB/s from 8 to 64 MB."""
import sys; sys.path.insert(0,'.')
from multimodal_b200 import _native
for mb in (8, 16, 24, 32, 48, 64, 96):
    print(mb, "MB", round(_native.l2_read_bandwidth(bytes=mb<<20, iters=200)), "GB/s")
