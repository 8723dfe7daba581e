"""Multi-GPU plumbing of the front end: one process per GPU.

SURVEY.md §8(e):
  * extraction + per-frame matching shard by FRAMES -- contiguous blocks, no data-path collective;
  * brute-force kNN over a large map shards the database ROWS, replicates the queries, and merges the
    per-rank lexicographic ``(dist, global index)`` top-2 after one all-gather of ``q x 2`` 64-bit keys;
  * ProjectionMatch over a large local map shards the MAP POINTS, replicates the frame, and merges the
    per-keypoint ``(dist, -global query)`` keys after one all-gather of ``m`` keys
    (minimum = "smaller distance, later query wins ties", reference src/matcher.cpp:197-204).
The keys are exact integers and ties break on global indices, so the result does not depend on where the
shard boundaries fall.

Two exchanges, same results:
  * ``exchange="peer"`` (default on GPUs): the library's own step (include/sfe.h ``sfe_comm_*``, ``sfe_knn2_sharded``,
    ``sfe_projection_match_sharded``) -- each rank's kernel stores its keys straight into every peer's inbox over NVLink
    peer memory and the merge kernel waits on flags; ``torch.distributed`` only carries the 64-byte IPC handles once, at
    set-up time;
  * ``exchange="collective"``: ``all_gather_into_tensor`` over NCCL (or gloo on CPU, where tests/test_sharding_gloo.py lets
    the oracle stand in for the kernels) between ``Matcher.knn2_dev`` and ``knn2_merge_dev``.
"""
from __future__ import annotations

import numpy as np

NO_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def block(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of `n` units owned by `rank`: the first n % world ranks hold one more."""
    if world < 1 or not 0 <= rank < world or n < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_keys(local_keys, group=None):
    """all_gather_into_tensor of one rank's key tensor -> (world, *local_keys.shape) on the same device."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    shape = tuple(local_keys.shape)
    flat = local_keys.new_empty((world * shape[0],) + shape[1:])   # ranks concatenated along dim 0
    if world == 1:
        flat.copy_(local_keys)
    else:
        dist.all_gather_into_tensor(flat, local_keys.contiguous(), group=group)
    return flat.view((world,) + shape)


def peer_comm(matcher, group=None):
    """A connected api.Comm for this rank: create, all-gather the 64-byte handles (the only use of the process group), connect."""
    import torch.distributed as dist
    from .api import Comm
    on = dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    comm = Comm(matcher.device, rank, world)
    if world == 1:
        return comm  # nothing to map
    handles = [None] * world
    dist.all_gather_object(handles, comm.export(), group=group)
    return comm.connect(handles)


def as_se3(Tcw):
    """The sharded library call takes predicted_Tcw as the reference holds it (g2o::SE3Quat: qx, qy, qz, qw, tx, ty, tz).
    A matrix is accepted only when its rotation is the identity (no quaternion conversion is guessed here)."""
    a = np.asarray(Tcw, np.float64)
    if a.ndim == 1 and a.size == 7:
        return a
    if a.ndim == 2 and np.array_equal(a[:3, :3], np.eye(3)):
        return np.array([0, 0, 0, 1, a[0, 3], a[1, 3], a[2, 3]], np.float64)
    raise ValueError("pass the pose as (qx, qy, qz, qw, tx, ty, tz)")


def pack_knn_keys(quad: np.ndarray) -> np.ndarray:
    """{idx0, dist0, idx1, dist1} rows (local or global indices) -> uint64 keys dist << 32 | idx, q x 2."""
    quad = np.asarray(quad, np.int64)
    keys = np.empty((len(quad), 2), np.uint64)
    for r in range(2):
        idx, dist = quad[:, 2 * r], quad[:, 2 * r + 1]
        keys[:, r] = np.where(idx < 0, NO_KEY, (dist.astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64))
    return keys


class ShardedDatabase:
    """A descriptor map sharded by rows over the ranks of a process group; queries are replicated.

    `rows_global` counts the whole map; this rank holds rows [start, stop) = block(rows_global, world, rank)
    and was given exactly those rows in `local_rows`."""

    def __init__(self, matcher, local_rows, rows_global, group=None, exchange="peer", comm=None):
        import torch
        import torch.distributed as dist
        self.m, self.group, self.exchange = matcher, group, exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.start, self.stop = block(rows_global, self.world, self.rank)
        if len(local_rows) != self.stop - self.start:
            raise ValueError("local_rows does not match this rank's block")
        self.db = matcher.create_db(local_rows, idx_base=self.start)
        self.device = torch.device("cuda", matcher.device)
        self.comm = comm if comm is not None else (peer_comm(matcher, group) if exchange == "peer" else None)

    def knn2_dev(self, queries_ptr, q, out_ptr):
        """resident queries (the same on every rank) -> out_ptr int32 q x 4 on this rank's GPU; returns once enqueued
        (Matcher.wait() completes it).  The library's one-step exchange."""
        self.m.knn2_sharded(self.comm, self.db, queries_ptr, q, out_ptr)

    def knn2(self, queries: np.ndarray) -> np.ndarray:
        """-> int32 q x 4 {idx0, dist0, idx1, dist1} over the WHOLE map, identical on every rank."""
        import torch
        q = len(queries)
        d_q = torch.from_numpy(np.ascontiguousarray(queries, np.uint8)).to(self.device)
        if self.exchange == "peer":
            out = torch.empty((q, 4), dtype=torch.int32, device=self.device)
            self.knn2_dev(d_q.data_ptr(), q, out.data_ptr())
            self.m.wait()
            return out.cpu().numpy()
        keys = torch.empty((q, 2), dtype=torch.int64, device=self.device)  # uint64 bit patterns
        self.m.knn2_dev(self.db, d_q.data_ptr(), q, keys.data_ptr())
        self.m.wait()                                                       # an asynchronous matcher has only enqueued it
        gathered = gather_keys(keys, self.group)                            # NCCL over NVLink: world x q x 2 keys
        torch.cuda.current_stream(self.device).synchronize()
        out = torch.empty((q, 4), dtype=torch.int32, device=self.device)
        self.m.knn2_merge_dev(gathered.data_ptr(), self.world, q, out.data_ptr())
        self.m.wait()
        return out.cpu().numpy()


class ShardedLocalMap:
    """Map points sharded over the ranks for ProjectionMatch; the frame's keypoints are replicated."""

    def __init__(self, matcher, xw_local, desc_local, points_global, group=None, exchange="peer", comm=None):
        import torch
        import torch.distributed as dist
        self.m, self.group, self.exchange = matcher, group, exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.start, self.stop = block(points_global, self.world, self.rank)
        if len(xw_local) != self.stop - self.start or len(desc_local) != len(xw_local):
            raise ValueError("local map points do not match this rank's block")
        self.device = torch.device("cuda", matcher.device)
        self.xw = torch.from_numpy(np.ascontiguousarray(xw_local, np.float64)).to(self.device)
        self.desc = torch.from_numpy(np.ascontiguousarray(desc_local, np.uint8)).to(self.device)
        self.comm = comm if comm is not None else (peer_comm(matcher, group) if exchange == "peer" else None)

    def projection_match_dev(self, frame, Tcw, radius, to_q_ptr, dist_ptr, best12=0.5):
        """resident frame (api.Frame, the same on every rank), Tcw as SE3 (7 numbers) -> per-keypoint global map-point index /
        distance on this rank's GPU; returns once enqueued.  The library's one-step exchange."""
        self.m.projection_match_sharded(self.comm, frame, self.xw.data_ptr(), self.desc.data_ptr(), None, len(self.xw), self.start,
                                        Tcw, radius, to_q_ptr, dist_ptr, best12)

    def projection_match(self, Tcw, camera, kps, kp_desc, radius, best12=0.5):
        """-> (kp_to_query, kp_dist) with GLOBAL map-point indices, identical on every rank."""
        import torch
        from .api import KP_DTYPE, Frame
        m = len(kps)
        if self.exchange == "peer":
            fr = Frame(self.m, kps, kp_desc, camera)
            to_q = torch.empty((m,), dtype=torch.int32, device=self.device)
            dist_out = torch.empty((m,), dtype=torch.int32, device=self.device)
            self.projection_match_dev(fr, as_se3(Tcw), radius, to_q.data_ptr(), dist_out.data_ptr(), best12)
            self.m.wait()
            return to_q.cpu().numpy(), dist_out.cpu().numpy()
        d_kps = torch.from_numpy(np.ascontiguousarray(kps, KP_DTYPE).view(np.uint8)).to(self.device)
        d_kd = torch.from_numpy(np.ascontiguousarray(kp_desc, np.uint8)).to(self.device)
        keys = torch.empty((m,), dtype=torch.int64, device=self.device)
        self.m.projection_match_keys_dev(self.xw.data_ptr(), self.desc.data_ptr(), None, len(self.xw), self.start, Tcw, camera,
                                         d_kps.data_ptr(), d_kd.data_ptr(), m, radius, keys.data_ptr(), best12)
        self.m.wait()
        gathered = gather_keys(keys, self.group)
        torch.cuda.current_stream(self.device).synchronize()
        to_q = torch.empty((m,), dtype=torch.int32, device=self.device)
        dist_out = torch.empty((m,), dtype=torch.int32, device=self.device)
        self.m.projection_merge_dev(gathered.data_ptr(), self.world, m, to_q.data_ptr(), dist_out.data_ptr())
        self.m.wait()
        return to_q.cpu().numpy(), dist_out.cpu().numpy()
