"""Multi-GPU plumbing of the anchoring path: one process per GPU, torch.distributed.

Every seed belongs to one rank (a hash of forward ^ reverse-complement mer, csrc/common.cuh seed_owned).  Two exchanges exist
on the data path, both tiny:
(1) a SUM all-reduce of the unique-seed bitmaps (1 bit per genome-0 position; the ranks' bits are
disjoint) between enumeration and extension, because a match is emitted by its leftmost unique
seed whichever rank owns that seed; (2) the gather of match rows to rank 0: an all_gather of the
per-rank row counts followed by a variable-length gather of 24-byte rows.  NCCL over NVLink on the
GPU box; gloo on CPU in the unit tests.
"""
import torch
import torch.distributed as dist


def shard_of(rank, world):
    """(shard_index, shard_count) handed to mcu_session_run: rank r owns canonical-key slice r of `world` equal slices"""
    return rank, world


class _DeviceWords:
    """zero-copy view of a device buffer of int32 words for torch (CUDA array interface)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def allreduce_uniq_bitmap(session, group=None):
    """Combine the ranks' unique-seed bitmaps in place.  Every genome-0 position belongs to exactly one rank's key slice,
    so the bit sets are disjoint and an integer SUM of the words is their OR (NCCL has no bitwise reduction)."""
    ptr, n = session.uniq_bitmap()
    words = torch.as_tensor(_DeviceWords(ptr, n), device="cuda")
    or_disjoint_words(words, group)
    torch.cuda.current_stream().synchronize()


def or_disjoint_words(words: torch.Tensor, group=None):
    """in-place bitwise OR across ranks of int32 words whose set bits are disjoint between ranks (SUM == OR, carry-free)"""
    dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
    return words


def run_sharded(session, seed, rank, world, group=None):
    """One sharded pass: enumerate this rank's slice, combine bitmaps, extend, gather rows to rank 0 and merge there.
    Returns the number of matches on rank 0 (the session then holds the merged list), else this rank's own count."""
    session.enumerate(seed, rank, world)
    allreduce_uniq_bitmap(session, group)
    n = session.finish(uniq_is_global=True)
    rows = torch.empty((n, 3), dtype=torch.int64, device="cuda")
    if n:
        session.download_ptr(rows.data_ptr())
    allrows = gather_rows(rows, 0, group)
    if rank == 0:
        total, _ = session.merge(allrows.data_ptr(), in_device=True, n=allrows.shape[0])
        return total
    return n


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
