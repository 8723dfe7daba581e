import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from slam_toolkit_b200 import api, synth
F=128
base=[synth.stereo_pair(s) for s in range(8)]
L=np.stack([base[i%8][0] for i in range(F)]); R=np.stack([base[i%8][1] for i in range(F)])
hs=[api.ORBextractor(max_images=2*F, device=local) for _ in range(2)]
pins=[]
for h in hs:
    pl,pr=api.PinnedArray(L.shape,np.uint8),api.PinnedArray(R.shape,np.uint8)
    pl.array[:],pr.array[:]=L,R
    pins.append((pl,pr,h.alloc_stereo_out(F,pinned=True)))
def steps(t,c):
    pl,pr,o=pins[t]
    for _ in range(c): hs[t].stereo_frames(pl.array,pr.array,o)
def timed(n,K=30):
    ths=[threading.Thread(target=steps,args=(t,K//n)) for t in range(n)]
    t0=time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    return F*K/(time.perf_counter()-t0)
for t in range(2): steps(t,3)
mode=sys.argv[1]
for rep in range(3):
    if mode=="barrier": dist.barrier()
    a=timed(1)
    if mode=="barrier": dist.barrier()
    b=timed(2)
    print(f"rank {rank} rep {rep} mode {mode}: 1 thread {a:.0f}  2 threads {b:.0f}", flush=True)
dist.barrier()
dist.destroy_process_group()
