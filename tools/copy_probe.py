#!/usr/bin/env python3
"""Copy-only probe of the end-to-end path's ceiling: every rank moves exactly what one bench step moves (F stereo frames of
pinned 1241x376 images host -> device in sub-batches, the step's results device -> host), H2D and D2H on two streams at
once, all ranks at the same time, no kernels.  What comes out is the host-memory / PCIe ceiling that `e2e` cannot beat.

    python tools/copy_probe.py                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/copy_probe.py

Variants: default pinned input pages vs write-combined ones (cudaHostAllocWriteCombined), H2D alone, D2H alone, both.
Rank 0 prints one JSON line per variant: per-rank GB/s (min / max) and the aggregate, plus the frames/s the aggregate
H2D rate corresponds to.  torch.distributed (gloo) is used for the rendezvous and barriers only."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W, H, F = 1241, 376, 128
CHUNKS = 8                       # a host call uploads its batch in sub-batches of 16 stereo frames
D2H_BYTES = 35_432_448           # what one 128-frame step returns (bench.py e2e.d2h_bytes_per_step)


def cudart():
    import torch  # noqa: F401  (loads the CUDA runtime the wheel ships)
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    import glob
    import torch as t
    for p in glob.glob(os.path.join(os.path.dirname(t.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        return C.CDLL(p)
    raise OSError("libcudart not found")


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    rt = cudart()
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaStreamCreateWithFlags.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
    rt.cudaStreamSynchronize.argtypes = [C.c_void_p]

    def ck(e):
        if e != 0:
            raise RuntimeError(f"CUDA error {e}")
    ck(rt.cudaSetDevice(local))
    h2d_bytes = 2 * F * W * H

    def host(nbytes, flags):
        p = C.c_void_p()
        ck(rt.cudaHostAlloc(C.byref(p), nbytes, flags))
        C.memset(p, 1, nbytes)       # touch the pages (write only: fine for write-combined memory too)
        return p

    def dev(nbytes):
        p = C.c_void_p()
        ck(rt.cudaMalloc(C.byref(p), nbytes))
        return p
    PORTABLE, WC = 1, 4
    h_in = {"pinned": host(h2d_bytes, PORTABLE), "write_combined": host(h2d_bytes, PORTABLE | WC)}
    h_out = host(D2H_BYTES, PORTABLE)
    d_in, d_out = dev(h2d_bytes), dev(D2H_BYTES)
    s_in, s_out = C.c_void_p(), C.c_void_p()
    ck(rt.cudaStreamCreateWithFlags(C.byref(s_in), 1))
    ck(rt.cudaStreamCreateWithFlags(C.byref(s_out), 1))

    def barrier():
        if dist is not None:
            dist.barrier()

    def run(kind, up, down, seconds=2.0):
        src = h_in[kind]
        chunk = h2d_bytes // CHUNKS
        steps = 0

        def step():
            if up:
                for c in range(CHUNKS):
                    ck(rt.cudaMemcpyAsync(C.c_void_p(d_in.value + c * chunk), C.c_void_p(src.value + c * chunk), chunk, 1, s_in))
            if down:
                for c in range(CHUNKS):
                    n = D2H_BYTES // CHUNKS
                    ck(rt.cudaMemcpyAsync(C.c_void_p(h_out.value + c * n), C.c_void_p(d_out.value + c * n), n, 2, s_out))
        for _ in range(3):
            step()
        rt.cudaStreamSynchronize(s_in); rt.cudaStreamSynchronize(s_out)
        barrier()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                step()
            rt.cudaStreamSynchronize(s_in); rt.cudaStreamSynchronize(s_out)
            steps += 4
        dt = time.perf_counter() - t0
        return steps, dt

    results = []
    for kind in ("pinned", "write_combined"):
        for name, up, down in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
            if kind == "write_combined" and name == "d2h_only":
                continue
            steps, dt = run(kind, up, down)
            up_gbs = steps * h2d_bytes / dt / 1e9 if up else 0.0
            dn_gbs = steps * D2H_BYTES / dt / 1e9 if down else 0.0
            rates = [(up_gbs, dn_gbs)]
            if dist is not None:
                allr = [None] * world
                dist.all_gather_object(allr, (up_gbs, dn_gbs))
                rates = allr
            if rank == 0:
                ups, dns = [r[0] for r in rates], [r[1] for r in rates]
                results.append({"input_pages": kind, "variant": name, "ranks": world, "h2d_gbs_total": sum(ups), "d2h_gbs_total": sum(dns),
                                "h2d_gbs_per_rank_min_max": [min(ups), max(ups)], "d2h_gbs_per_rank_min_max": [min(dns), max(dns)],
                                "frames_per_s_at_this_h2d_rate": sum(ups) * 1e9 / (2 * W * H) if up else None})
                print(json.dumps(results[-1]), flush=True)
            barrier()
    if rank == 0:
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            print(topo, flush=True)
            print(subprocess.run(["lscpu"], capture_output=True, text=True, timeout=20).stdout, flush=True)
        except Exception:
            pass
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
