#!/usr/bin/env python3
"""bench.py -- ORB extract+match throughput on synthetic KITTI-shaped stereo frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of F consecutive stereo frames (1241x376, 8 levels, scale 1.2, 2000 features) through
the hot path, BASELINE.json configs[1] as SURVEY.md §8d defines it: extract(left) + extract(right) + StereoMatch
(the keyframe path of the reference pipeline, src/pipeline.cpp:243-249) + the tracker's ProjectionMatch of the
previous frame's stereo-triangulated keypoints into the frame (StereoFrame::GetDepth, identity motion prior, r = 50).
  value : whole-job stereo frames/s with inputs resident in HBM (sfe_stereo_sequence_dev), timed
          with CUDA events on the extractor's own stream, max over ranks.
  e2e   : the same metric through the host entry point (sfe_stereo_sequence): pinned host images in,
          host keypoints/descriptors/stereo indices/track indices out, copies inside the timed region.
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
SEQ_LEN, SEQ_STEP = 16, 4  # a batch is made of 16-frame sequences: the camera slides 4 px per frame over one synthetic scene
PYR_PIXELS = 1_444_097   # sum of level sizes
B_IMG = 466_616 + PYR_PIXELS + 2000 * 60          # algorithmic bytes per image extraction
B_FRAME = 2 * B_IMG + 4 * 2000                     # per stereo frame
WORKLOAD = ("kitti_stereo_frontend: extract L + extract R + StereoMatch + ProjectionMatch(previous frame's stereo points, r=50), "
            "1241x376, 8 levels, 1.2, 2000 feats")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames", type=int, default=128, help="stereo frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=0, help="stereo frames in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads (one handle each) issuing the e2e calls")
    ap.add_argument("--knn-rows", type=int, default=10_000_000, help="rows of the descriptor map of the Hamming leg (0 = skip)")
    ap.add_argument("--proj-points", type=int, default=500_000, help="map points of the ProjectionMatch leg (0 = skip)")
    ap.add_argument("--clock-period", type=float, default=0.05, help="seconds between NVML clock samples (0 = no sampling)")
    return ap.parse_args()


def make_frames(n):
    """n consecutive stereo frames: 16-frame sequences over scenes 0, 1, 2, ... (synth.stereo_sequence)."""
    from slam_toolkit_b200 import synth
    Ls, Rs = [], []
    for s in range((n + SEQ_LEN - 1) // SEQ_LEN):
        l, r = synth.stereo_sequence(s, min(SEQ_LEN, n - s * SEQ_LEN), SEQ_STEP)
        Ls.append(l)
        Rs.append(r)
    return np.concatenate(Ls), np.concatenate(Rs)


def track_params():
    from slam_toolkit_b200 import api, synth
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, (0, 0, 0, 0), W, H)
    return api.TrackParams.make(cam, synth.KITTI_BASELINE, None, 50.0)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md): NVML every 50 ms,
    `nvidia-smi --query-gpu` every 200 ms when the NVML binding is unavailable.  Rank 0 alone samples, for every GPU of
    the job: NVML queries from several processes at once serialise against the other ranks' CUDA calls in the driver
    (measured: 8 ranks polling at 50 Hz cut the end-to-end figure fivefold)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, devices, period=0.05):
        super().__init__(daemon=True)
        self.devices, self.sm, self.reasons, self.sm_max, self.stop_flag, self.source = list(devices), [], set(), None, False, None
        self.period, self.active = period, False   # samples are kept only while a timed region is running
        self.once = threading.Event()              # one extra sample on request (before / after the end-to-end region)
        self.once_done = threading.Event()

    def _nvml_loop(self):
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        handles = []
        for d in self.devices:
            idx = d
            if vis:
                try:
                    idx = int(vis.split(",")[d])
                except (ValueError, IndexError):
                    pass
            handles.append(N.nvmlDeviceGetHandleByIndex(idx))
        self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(handles[0], N.NVML_CLOCK_SM))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        last = 0.0
        while not self.stop_flag:
            due = self.active and time.perf_counter() - last >= self.period
            if (due or self.once.is_set()) and self.period > 0:
                last = time.perf_counter()
                for h in handles:
                    self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                    r = int(get_reasons(h))
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
                if self.once.is_set():
                    self.once.clear()
                    self.once_done.set()
            time.sleep(0.005)

    def _smi_loop(self):
        self.source = "nvidia-smi"
        while not self.stop_flag:
            if not self.active or self.period <= 0:
                time.sleep(0.01)
                continue
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.devices[0])],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [c.strip() for c in out.split(",")]
                if len(r) >= 9:
                    self.sm.append(float(r[1]))
                    self.sm_max = float(r[2])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def sample_once(self):
        """One sample now (blocking); used around regions where periodic NVML polling would perturb the measurement."""
        if not self.devices or self.period <= 0 or self.source != "nvml":
            return
        self.once_done.clear()
        self.once.set()
        self.once_done.wait(timeout=2)

    def run(self):
        if not self.devices:
            return
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no clock samples"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": self.source}


def cpu_sample(frames, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    oracle_c.build()
    from slam_toolkit_b200 import synth
    L, R = make_frames(frames)
    cam = oracle_c.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0, 0, 0, 0], W, H)
    t0 = time.perf_counter()
    matches, kps, tracked = oracle_c.stereo_sequence(L, R, threads, cam, synth.KITTI_BASELINE, 50.0)
    dt = time.perf_counter() - t0
    return frames / dt, dt, matches + tracked, kps


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference path (oracle/orb_oracle.c; the reference itself cannot be
    built here), all host threads, one frame per thread, rank 0 only.  One step = a bounded sample of the workload:
    its size is chosen from a warm-up measurement so that the K steps take about a minute."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    est_fps, _, _, _ = cpu_sample(max(cores, 4), cores)          # also the warm-up (builds + pages in the oracle)
    for _ in range(max(min(args.warmup, 2) - 1, 0)):
        est_fps, _, _, _ = cpu_sample(max(cores, 4), cores)
    frames = args.cpu_frames or int(min(max(60.0 * est_fps / max(args.steps, 1), cores), 4 * cores))
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
            "config": {"workload": WORKLOAD, "frames_per_step": frames},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{frames} synthetic stereo frames per step, {args.steps} steps, one frame per thread"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "matches_per_s": tot_matches / total, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def hamming_leg(args, dev, rank, world, dist):
    """BASELINE config 4 beside the headline: brute-force Hamming top-2 of Q queries against a 10 M-row descriptor map
    sharded by rows over the ranks (all-gather of the per-rank top-2 keys + merge kernel).  Q = 2000 is `popc`-bound
    (8 popc per pair), Q = 1 streams the map once: that one is the GB/s-vs-HBM figure."""
    from slam_toolkit_b200 import api, sharding, synth
    rows = args.knn_rows
    a, b = sharding.block(rows, world, rank)
    rng = np.random.default_rng(1234 + rank)
    local = rng.integers(0, 256, (b - a, 32), dtype=np.uint8)
    m = api.Matcher(dev)
    sd = sharding.ShardedDatabase(m, local, rows)
    queries, _ = synth.knn_queries(local[:100_000], 2000, seed=5678)   # the same on every rank only at world == 1: fine for timing
    if dist is not None:
        import torch
        qt = torch.from_numpy(queries).cuda(dev)
        dist.broadcast(qt, 0)
        queries = qt.cpu().numpy()
    sd.knn2(queries)
    reps = 5
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = sd.knn2(queries)
    dt = (time.perf_counter() - t0) / reps
    # streaming pass: one query, resident, kernel time by CUDA events on the matcher's stream
    dq, keys = api.DeviceBuffer(32, dev).upload(queries[:1]), api.DeviceBuffer(16, dev)
    m.knn2_dev(sd.db, dq.ptr, 1, keys.ptr)
    e0, e1 = api.Event(dev), api.Event(dev)
    m.set_async(True)
    e0.record(m)
    for _ in range(50):
        m.knn2_dev(sd.db, dq.ptr, 1, keys.ptr)
    e1.record(m)
    m.wait()
    m.set_async(False)
    ms1 = e0.elapsed_ms(e1) / 50
    if dist is not None:
        import torch
        t = torch.tensor([dt, ms1], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, ms1 = t.tolist()
    gbs1 = world * (b - a) * 32 / (ms1 / 1e3) / 1e9
    return {"map_rows": rows, "sharding": f"{world} row shard(s), all-gather of 2 keys per query + merge",
            "q2000_ms_per_batch": dt * 1e3, "q2000_pair_distances_per_s": 2000 * rows / dt,
            "q2000_algorithmic_gbs": (32 * rows + 48 * 2000) / dt / 1e9,
            "q1_stream_ms": ms1, "q1_stream_gbs": gbs1, "accepted_ratio_test": float((2 * out[:, 1] < out[:, 3]).mean())}


def projection_leg(args, dev, rank, world, dist, kps, desc):
    """BASELINE config 5 beside the headline: ProjectionMatch (r = 50 px, identity pose) of a local map of N points
    against one frame's keypoints; map points sharded over the ranks (all-gather of the per-keypoint keys + merge).
    Algorithmic bytes per call: N * (24 + 32) + M * (8 + 32) + 8 * M."""
    from slam_toolkit_b200 import api, sharding, synth
    n, m_kps = args.proj_points, len(kps)
    xy = np.stack([kps["x"], kps["y"]], axis=1).astype(np.float64)
    xw, mpd = synth.projection_scene(xy, desc, n, seed=99)
    a, b = sharding.block(n, world, rank)
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, (0, 0, 0, 0), W, H)
    m = api.Matcher(dev)
    d_xw = api.DeviceBuffer((b - a) * 24, dev).upload(np.ascontiguousarray(xw[a:b]))
    d_mpd = api.DeviceBuffer((b - a) * 32, dev).upload(np.ascontiguousarray(mpd[a:b]))
    d_kps = api.DeviceBuffer(m_kps * 28, dev).upload(np.ascontiguousarray(kps))
    d_kd = api.DeviceBuffer(m_kps * 32, dev).upload(np.ascontiguousarray(desc))
    d_q, d_d = api.DeviceBuffer(m_kps * 4, dev), api.DeviceBuffer(m_kps * 4, dev)
    Tcw = np.eye(4)

    def call():
        m.projection_match_dev(d_xw.ptr, d_mpd.ptr, None, b - a, Tcw, cam, d_kps.ptr, d_kd.ptr, m_kps, 50.0, d_q.ptr, d_d.ptr)
    call()
    e0, e1 = api.Event(dev), api.Event(dev)
    reps = 20
    m.set_async(True)
    e0.record(m)
    for _ in range(reps):
        call()
    e1.record(m)
    m.wait()
    m.set_async(False)
    ms = e0.elapsed_ms(e1) / reps
    matched = int((d_q.download((m_kps,), np.int32) >= 0).sum())
    sharded_ms = None
    if dist is not None:
        import torch
        lm = sharding.ShardedLocalMap(m, xw[a:b], mpd[a:b], n)
        lm.projection_match(Tcw, cam, kps, desc, 50.0)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            lm.projection_match(Tcw, cam, kps, desc, 50.0)
        t = torch.tensor([ms, (time.perf_counter() - t0) / 5 * 1e3], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, sharded_ms = t.tolist()
    alg = n * 56 + m_kps * 48
    return {"map_points": n, "frame_keypoints": m_kps, "radius_px": 50.0, "ms_per_call_shard_kernels": ms,
            "map_points_per_s": n / (ms / 1e3), "algorithmic_gbs": alg / (ms / 1e3) / 1e9,
            "keypoints_matched_this_shard": matched, "sharded_call_ms_with_allgather": sharded_ms,
            "sharding": f"{world} map-point shard(s)"}


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
    PITCH = api.image_pitch(W)           # resident images are pitched (cudaMallocPitch style): TMA needs 16-B multiples
    per_batch = 2 * F * PITCH * H

    def pitched(a):
        out = np.zeros((a.shape[0], H, PITCH), np.uint8)
        out[:, :, :W] = a
        return out
    NB = max(2, int(np.ceil(2.2 * 126e6 / per_batch)))
    d_left, d_right = [], []
    for b in range(NB):
        sh = (b * SEQ_LEN) % F       # rotated by whole sequences: consecutive frames stay consecutive
        d_left.append(api.DeviceBuffer(F * PITCH * H, dev).upload(pitched(np.roll(L, sh, axis=0))))
        d_right.append(api.DeviceBuffer(F * PITCH * H, dev).upload(pitched(np.roll(R, sh, axis=0))))
    spec = {"kps_l": 28 * cap * F, "desc_l": 32 * cap * F, "n_l": 4 * F, "kps_r": 28 * cap * F, "desc_r": 32 * cap * F,
            "n_r": 4 * F, "stereo_idx": 4 * cap * F, "stereo_dist": 4 * cap * F, "track_idx": 4 * cap * F,
            "track_dist": 4 * cap * F}
    tp = track_params()
    d_out = {k: api.DeviceBuffer(v, dev) for k, v in spec.items()}
    ptrs = {k: b.ptr for k, b in d_out.items()}

    def barrier():
        if dist is not None:
            dist.barrier()

    def step_resident(i):
        ex.stereo_sequence_dev(d_left[i % NB].ptr, d_right[i % NB].ptr, F, W, H, ptrs, tp, pitch=PITCH)

    sampler = ClockSampler(range(world) if rank == 0 else [], args.clock_period if rank == 0 else 0)
    sampler.start()                      # NVML initialises here, outside the timed regions
    for i in range(Wm):
        step_resident(i)
    # per-stage CUDA-event times: a few synchronous profiled steps outside the timed region
    ex.set_profiling(True)
    for i in range(min(K, 10)):
        step_resident(i)
    stage_ms, calls = ex.stage_ms()
    ex.set_profiling(False)
    # timed region: K batches queued back to back on the extractor's stream (asynchronous _dev calls), one wait
    ex.set_async(True)
    for i in range(2):
        step_resident(i)
    ex.wait()
    l0 = ex.launches()
    ev0, ev1 = api.Event(dev), api.Event(dev)
    barrier()
    sampler.active = True
    ev0.record(ex)
    for i in range(K):
        step_resident(Wm + i)
    ev1.record(ex)
    ex.wait()
    sampler.active = False
    ms = ev0.elapsed_ms(ev1)
    barrier()
    launches = ex.launches() - l0
    # BASELINE config 3 beside the headline: extraction only (sfe_extract_batch_dev) of the same resident images, 2F
    # independent frames per step (two calls of F), sharded over the ranks like the headline
    def step_extract(i):
        for d in (d_left[i % NB], d_right[i % NB]):
            api._check(api.lib().sfe_extract_batch_dev(ex.h, api._p(d.ptr), PITCH * H, F, W, H, PITCH, api._p(ptrs["kps_l"]),
                                                       api._p(ptrs["desc_l"]), cap, api._p(ptrs["n_l"])))
    KX = max(K // 4, 5)
    for i in range(3):
        step_extract(i)
    ex.wait()
    ev2, ev3 = api.Event(dev), api.Event(dev)
    barrier()
    ev2.record(ex)
    for i in range(KX):
        step_extract(3 + i)
    ev3.record(ex)
    ex.wait()
    ms_x = ev2.elapsed_ms(ev3) / KX
    barrier()
    ex.set_async(False)
    n_stereo = int((d_out["stereo_idx"].download((F, cap), np.int32) >= 0).sum())
    n_track = int((d_out["track_idx"].download((F, cap), np.int32) >= 0).sum())
    n_match = n_stereo + n_track
    n_kps = int(d_out["n_l"].download((F,), np.int32).sum() + d_out["n_r"].download((F,), np.int32).sum())

    # ---- e2e: pinned host images -> host results through the public host entry point (sfe_stereo_frames), synchronous
    # calls.  T host threads, one extractor handle each (handles are per-thread objects, include/sfe.h), take the K
    # steps in turn, so one call's pipeline fill / drain overlaps the other's steady state; T = 1 is reported beside it.
    T = max(1, args.e2e_threads)
    handles = [ex] + [api.ORBextractor(2000, 1.2, 8, 20, 7, device=dev, max_images=2 * F) for _ in range(T - 1)]
    pins = []
    for h in handles:
        pl, pr = api.PinnedArray((F, H, W), np.uint8), api.PinnedArray((F, H, W), np.uint8)
        pl.array[:], pr.array[:] = L, R
        pins.append((pl, pr, h.alloc_stereo_out(F, pinned=True, track=True)))
    out = pins[0][2]

    def e2e_steps(t, count):
        pl, pr, o = pins[t]
        for _ in range(count):
            handles[t].stereo_sequence(pl.array, pr.array, tp, o)

    def timed(nthreads):
        shares = [K // nthreads + (1 if t < K % nthreads else 0) for t in range(nthreads)]
        ths = [threading.Thread(target=e2e_steps, args=(t, shares[t])) for t in range(nthreads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return time.perf_counter() - t0

    for t in range(T):
        e2e_steps(t, Wm)
    barrier()
    # (periodic NVML polling during this region stalls the CUDA calls of the OTHER ranks inside the driver -- measured
    # 73 k -> 26 k frames/s at 2 GPUs -- so the clocks are sampled immediately before and after it instead)
    sampler.sample_once()
    e2e_s = timed(T)
    sampler.sample_once()
    barrier()
    e2e_single_s = timed(1) if T > 1 else e2e_s
    barrier()
    if os.environ.get("SFE_BENCH_DEBUG"):
        print(f"[rank {rank}] e2e {T} threads {F * K / e2e_s:.0f} frames/s, 1 thread {F * K / e2e_single_s:.0f}; again: "
              f"{F * K / timed(T):.0f} / {F * K / timed(1):.0f}", file=sys.stderr, flush=True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    hamming = hamming_leg(args, dev, rank, world, dist) if args.knn_rows > 0 else None
    projection = None
    if args.proj_points > 0:
        nl0 = int(out["n_l"][0])
        projection = projection_leg(args, dev, rank, world, dist, out["kps_l"][0, :nl0].copy(), out["desc_l"][0, :nl0].copy())
    h2d = 2 * F * W * H
    d2h = sum(v.nbytes for v in out.values())

    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_s * 1e3, e2e_single_s * 1e3, ms_x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, e2e_single_ms, ms_x = t.tolist()
        e2e_s, e2e_single_s = e2e_ms / 1e3, e2e_single_ms / 1e3
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
                 "stereo_match": F * (2 * 2000 * 40 + 2000 * 8),
                 "track": F * 2000 * (2 * 60 + 28 + 4 + 8) / 3.0}
    launches_per_step = {"pyramid": 7, "fast_cells": 1, "quadtree": 1, "blur": 1, "orient_describe": 1, "stereo_match": 1, "track": 3}
    top = max(stage_ms, key=lambda k: stage_ms[k])
    top_ms = stage_ms[top] / max(calls, 1) / launches_per_step[top]
    achieved = alg_bytes[top] / (top_ms / 1e3) / 1e9 if top_ms > 0 else 0.0
    # DRAM traffic of that kernel per launch from the committed ncu --set full capture (dram__bytes_read + write, per image
    # there, scaled to this run's images per launch); well above the algorithmic bytes would mean wasted re-reads
    traffic, traffic_src, issue = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        units = F if top in ("stereo_match", "track") else 2 * F
        traffic = tj["kernels"][top]["dram_bytes_per_image"] * units / launches_per_step[top]
        traffic_src = tj["source"]
        # the pipe that actually binds extraction: warp instructions of the kernel (ncu smsp__inst_executed.sum of the same
        # committed capture, per image) over its live CUDA-event time, against the issue peak 4 per clk per SM at the
        # sampled SM clock
        winst = tj["kernels"][top].get("warp_instructions_per_image", 0) * units / launches_per_step[top]
        clk = (sampler.summary() or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        sms = 148
        if winst > 0 and top_ms > 0:
            issue = {"warp_instructions_per_launch": winst, "achieved_gwarp_inst_per_s": winst / (top_ms / 1e3) / 1e9,
                     "peak_gwarp_inst_per_s": 4 * sms * clk * 1e6 / 1e9, "sm_clock_mhz": clk}
            issue["frac"] = issue["achieved_gwarp_inst_per_s"] / issue["peak_gwarp_inst_per_s"]
    except Exception:
        pass
    line = {"metric": "orb_extract_match_stereo_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "l2": f"inputs rotate over {NB} resident batches ({NB * per_batch / 1e6:.0f} MB > 126 MB L2)",
                       "resident_layout": f"row pitch {PITCH} B (sfe_image_pitch)",
                       "sharding": "frames partitioned across GPUs, no collective"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_threads": T, "single_thread_value": world * F * K / e2e_single_s},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg_bytes[top], "peak_source": peak_src,
                         "whole_step_achieved": value / world * B_FRAME / 1e9, "whole_step_frac": value / world * B_FRAME / 1e9 / hbm_peak,
                         "issue": issue,
                         "note": "extraction is integer-issue bound, not HBM bound (SURVEY.md §8d): `issue` = the kernel's warp "
                                 "instructions over its live time against 4 issues/clk/SM; profiles/ holds the pipe utilisation"},
            "stage_ms_per_step": {k: v / max(calls, 1) for k, v in stage_ms.items()},
            "keypoints_per_frame": n_kps / F, "matches_per_s": value * n_match / F,
            "stereo_matches_per_frame": n_stereo / F, "tracked_matches_per_frame": n_track / F,
            "extract_only": {"images_per_s": world * 2 * F / (ms_x / 1e3), "ms_per_call_of_F_images": ms_x / 2,
                             "algorithmic_gbs": world * 2 * F / (ms_x / 1e3) * B_IMG / 1e9,
                             "note": "BASELINE config 3: sfe_extract_batch_dev on resident images, no matching"}}
    if hamming is not None:
        hamming["q1_stream_frac_of_hbm_peak"] = hamming["q1_stream_gbs"] / (world * hbm_peak)
        line["hamming"] = hamming
    if projection is not None:
        projection["frac_of_hbm_peak"] = projection["algorithmic_gbs"] / hbm_peak
        line["projection_match"] = projection
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        frames = args.cpu_frames or 32 * cores   # ~2 s per core: 10-30 s of CPU work in all
        cpu_sample(cores, cores)                 # warm-up
        fps_all, dt_all, _, _ = cpu_sample(frames, cores)
        fps_1, dt_1, _, _ = cpu_sample(8, 1)
        line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"{frames} synthetic stereo frames, one frame per thread ({dt_all:.1f} s)",
                                "single_thread_value": fps_1}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
