"""Image-batch sharding across the GPUs of one box.

The reference is strictly batch-1 and single-device (demo/demo_match.py:29); every image is
independent, so the batch is partitioned contiguously over ranks, each rank runs the whole
pipeline locally, and ONE collective gathers the fixed-size keypoint records.  There is no
exchange inside the network (the squeeze-excite mean and the grid gate are per image).

``torch.distributed`` is the plumbing: NCCL over NVLink on the GPU box, gloo in the CPU tests.
``KeypointGather`` is the C-ABI form of the same collective (``balf_gather_keypoints`` on an ``ncclComm_t`` owned by the
library's caller, include/balf_b200.h): the unique id travels through ``torch.distributed``, the data path does not.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous [start, end) of rank's share of n_items (the first n_items % world ranks get one more)."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_records(xy, score, count):
    """xy int32 [B,K,2], score fp32 [B,K], count int32 [B] -> one int32 buffer [B, 3K+1]."""
    b, k, _ = xy.shape
    buf = torch.empty(b, 3 * k + 1, dtype=torch.int32, device=xy.device)
    buf[:, :2 * k] = xy.reshape(b, 2 * k)
    buf[:, 2 * k:3 * k] = score.contiguous().view(torch.int32)
    buf[:, 3 * k] = count
    return buf


def unpack_records(buf):
    b, n = buf.shape
    k = (n - 1) // 3
    xy = buf[:, :2 * k].reshape(b, k, 2)
    score = buf[:, 2 * k:3 * k].contiguous().view(torch.float32)
    return xy, score, buf[:, 3 * k].contiguous()


def gather_keypoints(xy, score, count, group=None):
    """All-gather the per-rank keypoint records (equal local batch on every rank) with a single
    collective -> (xy [B_total,K,2], score [B_total,K], count [B_total]) in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return xy, score, count
    world = dist.get_world_size(group)
    local = pack_records(xy, score, count)
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return unpack_records(out)


class KeypointGather:
    """NCCL all-gather of keypoint records through the C-ABI (``balf_gather_keypoints``).  Needs an initialised
    ``torch.distributed`` group only to hand the 128-byte ncclUniqueId from rank 0 to the other ranks."""

    def __init__(self, device, group=None):
        from . import _capi
        self._capi = _capi
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        uid = [_capi.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=group)
        self.comm = _capi.nccl_comm_create(uid[0], self.world, self.rank, device)

    def __call__(self, xy, score, count):
        return self._capi.gather_keypoints(self.comm, self.world, xy, score, count)

    def close(self):
        if self.comm is not None:
            self._capi.nccl_comm_destroy(self.comm)
            self.comm = None


def gather_matches(ids, n_matches, group=None):
    """ids int32 [P,Mmax,2], n_matches int32 [P] (pairs handled by this rank) -> gathered over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ids, n_matches
    world = dist.get_world_size(group)
    p, m, _ = ids.shape
    local = torch.empty(p, 2 * m + 1, dtype=torch.int32, device=ids.device)
    local[:, :2 * m] = ids.reshape(p, 2 * m)
    local[:, 2 * m] = n_matches
    out = torch.empty((world * p, 2 * m + 1), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:, :2 * m].reshape(world * p, m, 2), out[:, 2 * m].contiguous()
