import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth
L, R = synth.stereo_pair(0)
ex = api.ORBextractor(max_images=2)
pl, pr = api.PinnedArray(L.shape, np.uint8), api.PinnedArray(R.shape, np.uint8)
pl.array[:], pr.array[:] = L, R
out = ex.alloc_stereo_out(1, pinned=True)
for _ in range(20): ex.stereo_frames(pl.array[None], pr.array[None], out)
n = 400
t0 = time.perf_counter()
for _ in range(n): ex.stereo_frames(pl.array[None], pr.array[None], out)
t2 = time.perf_counter()
print(f"stereo_frames(1 pair): {1e6*(t2-t0)/n:.1f} us/call")
