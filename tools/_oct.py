import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ex = api.ORBextractor(max_images=2 * F)
Ls, Rs = zip(*[synth.stereo_pair(i) for i in range(F)])
pl, pr = api.PinnedArray((F,) + Ls[0].shape, np.uint8), api.PinnedArray((F,) + Ls[0].shape, np.uint8)
pl.array[:], pr.array[:] = np.stack(Ls), np.stack(Rs)
out = ex.alloc_stereo_out(F, pinned=True)
for _ in range(3): ex.stereo_frames(pl.array, pr.array, out)
