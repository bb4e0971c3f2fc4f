"""Multi-GPU plumbing of the anchoring path: one process per GPU, torch.distributed.

The only data-path collective is the gather of match rows to rank 0: an all_gather of the
per-rank row counts (one int64) followed by a variable-length gather of 24-byte rows
(NCCL over NVLink on the GPU box; gloo on CPU in the unit tests).  Shards are seed-key prefix
ranges, so no other exchange exists.
"""
import torch
import torch.distributed as dist


def shard_of(rank, world):
    """(shard_index, shard_count) handed to mcu_session_run: rank r owns canonical-key slice r of `world` equal slices"""
    return rank, world


def gather_rows(rows: torch.Tensor, dst: int = 0, group=None):
    """rows: [n, 3] int64 tensor (cuda for nccl, cpu for gloo).  Returns the concatenation of all ranks' rows
    (in rank order) on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    rows = rows.contiguous()
    if rank == dst:
        parts = [torch.empty((c, 3), dtype=torch.int64, device=rows.device) for c in counts]
        parts[dst] = rows
        reqs = [dist.irecv(parts[r], src=r, group=group) for r in range(world) if r != dst and counts[r] > 0]
        for q in reqs:
            q.wait()
        return torch.cat(parts, dim=0) if parts else rows
    if rows.shape[0] > 0:
        dist.send(rows, dst=dst, group=group)
    return None
