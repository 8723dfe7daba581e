import os, sys, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth
F=128
base=[synth.stereo_pair(s) for s in range(8)]
L=np.stack([base[i%8][0] for i in range(F)]); R=np.stack([base[i%8][1] for i in range(F)])
ex=api.ORBextractor(max_images=2*F)
pl,pr=api.PinnedArray(L.shape,np.uint8),api.PinnedArray(R.shape,np.uint8)
pl.array[:],pr.array[:]=L,R
out=ex.alloc_stereo_out(F,pinned=True)
for i in range(4):
    t0=time.perf_counter(); ex.stereo_frames(pl.array,pr.array,out); print("call ms", 1e3*(time.perf_counter()-t0), file=sys.stderr)
