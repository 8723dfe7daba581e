import torch, time
n=120*1024*1024
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(36*1024*1024,dtype=torch.uint8).pin_memory(); d2=torch.empty(36*1024*1024,dtype=torch.uint8,device='cuda')
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    return 20*n/dt/1e9
run(False); print("H2D alone GB/s", run(False)); print("H2D with concurrent D2H GB/s", run(True))
