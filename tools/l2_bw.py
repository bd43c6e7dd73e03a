import sys; sys.path.insert(0,'.')
from multimodal_b200 import _native
for mb in (8, 16, 24, 32, 48, 64, 96):
    print(mb, "MB", round(_native.l2_read_bandwidth(bytes=mb<<20, iters=200)), "GB/s")
