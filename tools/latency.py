#!/usr/bin/env python3
"""Latency of the reference-shaped calls (one image / one stereo frame per call, host buffers, synchronous):
what ORB_SLAM2::ORBextractor::extract and the keyframe path cost when the caller does not batch, and how the
kernel time of such a call splits over the stages (CUDA events)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth
L, R = synth.stereo_pair(0)
ex = api.ORBextractor(max_images=2)
pl, pr = api.PinnedArray(L.shape, np.uint8), api.PinnedArray(R.shape, np.uint8)
pl.array[:], pr.array[:] = L, R
out = ex.alloc_stereo_out(1, pinned=True)
for _ in range(20):
    ex.extract(pl.array); ex.stereo_frames(pl.array[None], pr.array[None], out)
n = 300
t0 = time.perf_counter()
for _ in range(n): ex.extract(pl.array)
t1 = time.perf_counter()
for _ in range(n): ex.stereo_frames(pl.array[None], pr.array[None], out)
t2 = time.perf_counter()
print(f"extract(image): {1e6*(t1-t0)/n:.0f} us/call   stereo_frames(1 pair): {1e6*(t2-t1)/n:.0f} us/call")
ex.set_profiling(True)
for _ in range(50): ex.stereo_frames(pl.array[None], pr.array[None], out)
ms, calls = ex.stage_ms()
print({k: round(1e3*v/calls,1) for k,v in ms.items()}, "us per call; sum", round(1e3*sum(ms.values())/calls,1))
# the matcher calls of the reference's tracking / mapping threads: one StereoMatch per keyframe, one ProjectionMatch per
# frame against ~1500 local-map points (host arrays in, host arrays out, synchronous)
m = api.Matcher(0)
kl, dl = ex.extract(pl.array)
kr, dr = ex.extract(pr.array)
xy = np.stack([kl["x"], kl["y"]], 1)
xw, mpd = synth.projection_scene(xy, dl, 1500, seed=3)
cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, (0, 0, 0, 0), synth.KITTI_W, synth.KITTI_H)
for _ in range(20):
    m.StereoMatch(kl, dl, kr, dr); m.ProjectionMatch(xw, mpd, None, np.eye(3, 4), cam, kl, dl, 50.0)
t0 = time.perf_counter()
for _ in range(n): m.StereoMatch(kl, dl, kr, dr)
t1 = time.perf_counter()
for _ in range(n): m.ProjectionMatch(xw, mpd, None, np.eye(3, 4), cam, kl, dl, 50.0)
t2 = time.perf_counter()
f = api.Frame(m, kl, dl, cam)
for _ in range(20): f.ProjectionMatch(xw, mpd, None, np.eye(3, 4), 50.0)
t3 = time.perf_counter()
for _ in range(n): f.ProjectionMatch(xw, mpd, None, np.eye(3, 4), 50.0)
t4 = time.perf_counter()
print(f"StereoMatch(2000 x 2000): {1e6*(t1-t0)/n:.0f} us/call   ProjectionMatch(1500 points, frame uploaded per call): "
      f"{1e6*(t2-t1)/n:.0f} us/call   against a resident frame: {1e6*(t4-t3)/n:.0f} us/call")

# the same through the C++ adapter (include/sfe_adapter.hpp), no Python in the loop
import subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp")], stdout=subprocess.DEVNULL)
L.tofile("/tmp/lat_l.raw"); R.tofile("/tmp/lat_r.raw")
print("C++ adapter:", subprocess.run([os.path.join(root, "tests", "cpp", "adapter_latency"), "/tmp/lat_l.raw", "/tmp/lat_r.raw",
                                      str(L.shape[1]), str(L.shape[0]), "500"], capture_output=True, text=True).stdout.strip())
