#!/usr/bin/env python3
"""bench.py -- ORB extract+match throughput on synthetic KITTI-shaped stereo frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of F stereo frames (1241x376, 8 levels, scale 1.2, 2000 features) through
the hot path: extract(left) + extract(right) + StereoMatch, i.e. the keyframe path of the
reference pipeline (src/pipeline.cpp:243-249), BASELINE.json configs[1].
  value : whole-job stereo frames/s with inputs resident in HBM (sfe_stereo_frames_dev), timed
          with CUDA events on the extractor's own stream, max over ranks.
  e2e   : the same metric through the host entry point (sfe_stereo_frames): pinned host images in,
          host keypoints/descriptors/stereo indices out, copies inside the timed region.
  roofline     : the dominant kernel's algorithmic bytes / its CUDA-event time, vs the measured HBM peak.
  cpu_baseline : the CPU oracle (port of the reference) on a bounded sample, host cores stated.
--impl reference times that CPU oracle alone (the reference itself cannot be built here:
OpenCV 3.4 C++/Eigen/g2o/FLANN are absent), all host threads, same metric/config.
Frames are sharded across GPUs with no data-path collective (weak scaling: F frames per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1241, 376
N_DISTINCT = 8           # distinct synthetic frames; batches are rotations of them
PYR_PIXELS = 1_444_097   # sum of level sizes
B_IMG = 466_616 + PYR_PIXELS + 2000 * 60          # algorithmic bytes per image extraction
B_FRAME = 2 * B_IMG + 4 * 2000                     # per stereo frame


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=64, help="stereo frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=0, help="stereo frames in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_frames(n):
    from slam_toolkit_b200 import synth
    base = [synth.stereo_pair(s) for s in range(min(n, N_DISTINCT))]
    L = np.stack([base[i % len(base)][0] for i in range(n)])
    R = np.stack([base[i % len(base)][1] for i in range(n)])
    return L, R


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_sample(frames, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    oracle_c.build()
    L, R = make_frames(frames)
    t0 = time.perf_counter()
    matches, kps = oracle_c.stereo_frames(L, R, threads)
    dt = time.perf_counter() - t0
    return frames / dt, dt, matches, kps


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference path, all host threads, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = args.cpu_frames or max(cores * 2, 16)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(max(cores, 4), cores)
    times = []
    tot_matches = 0
    for _ in range(args.steps):
        fps, dt, matches, _ = cpu_sample(frames, cores)
        times.append(dt)
        tot_matches += matches
    total = sum(times)
    value = frames * args.steps / total
    line = {"impl": "reference", "metric": "orb_extract_match_stereo_frames_per_s", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "kitti_stereo_frontend: extract L + extract R + StereoMatch, 1241x376, 8 levels, 1.2, 2000 feats",
                       "frames_per_step": frames},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{frames} synthetic stereo frames per step, {args.steps} steps, one frame per thread"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "matches_per_s": tot_matches / total, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from slam_toolkit_b200 import api
    F, K, Wm = args.frames, args.steps, max(args.warmup, 3)
    dev = local
    ex = api.ORBextractor(2000, 1.2, 8, 20, 7, device=dev, max_images=2 * F)
    cap = ex.cap
    # ---- inputs: NB rotated batches resident in HBM (NB*F*2*466 KB > L2 so no step re-reads a cached batch)
    L, R = make_frames(F)
    per_batch = 2 * F * W * H
    NB = max(2, int(np.ceil(2.2 * 126e6 / per_batch)))
    d_left, d_right = [], []
    for b in range(NB):
        sh = (b * 3) % F
        d_left.append(api.DeviceBuffer(F * W * H, dev).upload(np.roll(L, sh, axis=0)))
        d_right.append(api.DeviceBuffer(F * W * H, dev).upload(np.roll(R, sh, axis=0)))
    spec = {"kps_l": 28 * cap * F, "desc_l": 32 * cap * F, "n_l": 4 * F, "kps_r": 28 * cap * F, "desc_r": 32 * cap * F,
            "n_r": 4 * F, "stereo_idx": 4 * cap * F, "stereo_dist": 4 * cap * F}
    d_out = {k: api.DeviceBuffer(v, dev) for k, v in spec.items()}
    ptrs = {k: b.ptr for k, b in d_out.items()}

    def barrier():
        if dist is not None:
            dist.barrier()

    def step_resident(i):
        ex.stereo_frames_dev(d_left[i % NB].ptr, d_right[i % NB].ptr, F, W, H, ptrs)

    for i in range(Wm):
        step_resident(i)
    ex.set_profiling(True)
    l0 = ex.launches()
    sampler = ClockSampler(dev)
    sampler.start()
    ev0, ev1 = api.Event(dev), api.Event(dev)
    barrier()
    ev0.record(ex)
    for i in range(K):
        step_resident(Wm + i)
    ev1.record(ex)
    ms = ev0.elapsed_ms(ev1)
    barrier()
    launches = ex.launches() - l0
    stage_ms, calls = ex.stage_ms()
    ex.set_profiling(False)
    n_match = int((d_out["stereo_idx"].download((F, cap), np.int32) >= 0).sum())
    n_kps = int(d_out["n_l"].download((F,), np.int32).sum() + d_out["n_r"].download((F,), np.int32).sum())

    # ---- e2e: pinned host images -> host results through the public host entry point
    pin_l, pin_r = api.PinnedArray((F, H, W), np.uint8), api.PinnedArray((F, H, W), np.uint8)
    pin_l.array[:], pin_r.array[:] = L, R
    out = ex.alloc_stereo_out(F, pinned=True)
    for i in range(Wm):
        ex.stereo_frames(pin_l.array, pin_r.array, out)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        ex.stereo_frames(pin_l.array, pin_r.array, out)
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = 2 * F * W * H
    d2h = sum(v.nbytes for v in out.values())

    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t.tolist()
        e2e_s = e2e_ms / 1e3
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = world * F * K / (ms / 1e3)
    e2e = world * F * K / e2e_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernel = stage with the largest event time; algorithmic bytes per launch in DESIGN.md §5
    alg_bytes = {"pyramid": 2 * F * (466_616 + PYR_PIXELS) / 7.0,  # per launch: 7 launches per step
                 "fast_cells": 2 * F * (PYR_PIXELS + 11_800 * 4), "quadtree": 2 * F * (11_800 * 4 + 2000 * 4),
                 "blur": 2 * F * 2 * PYR_PIXELS, "orient_describe": 2 * F * (2000 * (749 + 512) + 2000 * 60),
                 "stereo_match": F * (2 * 2000 * 40 + 2000 * 8)}
    launches_per_step = {"pyramid": 7, "fast_cells": 1, "quadtree": 1, "blur": 1, "orient_describe": 1, "stereo_match": 1}
    top = max(stage_ms, key=lambda k: stage_ms[k])
    top_ms = stage_ms[top] / max(calls, 1) / launches_per_step[top]
    achieved = alg_bytes[top] / (top_ms / 1e3) / 1e9 if top_ms > 0 else 0.0
    line = {"metric": "orb_extract_match_stereo_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "kitti_stereo_frontend: extract L + extract R + StereoMatch, 1241x376, 8 levels, 1.2, 2000 feats",
                       "frames_per_step_per_gpu": F, "l2": f"inputs rotate over {NB} resident batches ({NB * per_batch / 1e6:.0f} MB > 126 MB L2)",
                       "sharding": "frames partitioned across GPUs, no collective"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                         "whole_step_achieved": value / world * B_FRAME / 1e9, "whole_step_frac": value / world * B_FRAME / 1e9 / hbm_peak,
                         "note": "extraction is integer/shared-memory bound, not HBM bound (SURVEY.md §8d)"},
            "stage_ms_per_step": {k: v / max(calls, 1) for k, v in stage_ms.items()},
            "keypoints_per_frame": n_kps / F, "matches_per_s": value * n_match / F}
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        frames = args.cpu_frames or max(2 * cores, 16)
        fps_all, dt_all, _, _ = cpu_sample(frames, cores)
        fps_1, dt_1, _, _ = cpu_sample(max(4, min(8, frames)), 1)
        line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"{frames} synthetic stereo frames, one frame per thread ({dt_all:.1f} s)",
                                "single_thread_value": fps_1}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
