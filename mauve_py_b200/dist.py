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
    counts = torch.cat(counts).tolist()   # one device->host read for all ranks' counts
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


# ---- gapped DP and HMM: independent regions / strings, no collective on the data path --------------------------------
def lpt_partition(costs, world):
    """Longest-processing-time-first assignment of independent work items to `world` ranks: items in descending cost order, each to
    the least loaded rank (ties: lowest rank).  Deterministic, so every rank derives the same plan from the same costs.
    Returns `world` ascending index lists."""
    import heapq
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    heap = [(0, r) for r in range(world)]
    parts = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(i)
        heapq.heappush(heap, (load + int(costs[i]), r))
    return [sorted(p) for p in parts]


def gather_bytes(payload: bytes, dst: int = 0, group=None, device=None):
    """variable-length gather of one byte string per rank to `dst` (list in rank order there, None elsewhere); the same counts +
    point-to-point pattern as gather_rows, on `device` (cuda for nccl, cpu for gloo)"""
    import numpy as np
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda") if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.from_numpy(np.frombuffer(payload, dtype=np.uint8).copy()).to(device)
    n = torch.tensor([buf.numel()], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = torch.cat(counts).tolist()
    if rank == dst:
        parts = [torch.empty(c, dtype=torch.uint8, device=device) for c in counts]
        parts[dst] = buf
        reqs = [dist.irecv(parts[r], src=r, group=group) for r in range(world) if r != dst and counts[r] > 0]
        for q in reqs:
            q.wait()
        return [p.cpu().numpy().tobytes() for p in parts]
    if buf.numel() > 0:
        dist.send(buf, dst=dst, group=group)
    return None


def align_sharded(pairs, rank, world, group=None, align=None):
    """muscle::GlobalAlign for a batch of region pairs divided among the ranks by LPT on lenA * lenB.  Every rank holds the
    whole `pairs` list (like the genomes), aligns its share on its GPU and ships (lengths, scores, edges) to rank 0, which returns
    the PWPaths in input order; other ranks return None.  `align` defaults to libmems.GlobalAlignBatch."""
    import numpy as np
    from . import libmems
    align = align or libmems.GlobalAlignBatch
    plan = lpt_partition([len(a) * len(b) for a, b in pairs], world)
    mine = plan[rank]
    paths = align([pairs[i] for i in mine]) if mine else []
    lens = np.array([len(p.edges) for p in paths], dtype=np.int64)
    scores = np.array([p.score for p in paths], dtype=np.int64)
    payload = lens.tobytes() + scores.tobytes() + b"".join(p.edges for p in paths)
    if world == 1:
        return paths
    got = gather_bytes(payload, 0, group)
    if rank != 0:
        return None
    out = [None] * len(pairs)
    for r, blob in enumerate(got):
        k = len(plan[r])
        ln = np.frombuffer(blob[:8 * k], dtype=np.int64)
        sc = np.frombuffer(blob[8 * k:16 * k], dtype=np.int64)
        pos = 16 * k
        for j, i in enumerate(plan[r]):
            out[i] = libmems.PWPath(blob[pos:pos + int(ln[j])], int(sc[j]))
            pos += int(ln[j])
    return out


def hmm_sharded(sequences, params, rank, world, group=None, run_batch=None):
    """run() of the HomologyHMM for a batch of column strings divided among the ranks by LPT on their lengths; rank 0 returns the
    H/N predictions in input order, other ranks None.  `run_batch` defaults to libmems.run_batch."""
    from . import libmems
    run_batch = run_batch or libmems.run_batch
    plan = lpt_partition([len(s) for s in sequences], world)
    mine = plan[rank]
    preds = run_batch([sequences[i] for i in mine], params) if mine else []
    if world == 1:
        return preds
    got = gather_bytes(b"".join(preds), 0, group)
    if rank != 0:
        return None
    out = [None] * len(sequences)
    for r, blob in enumerate(got):
        pos = 0
        for i in plan[r]:
            out[i] = blob[pos:pos + len(sequences[i])]
            pos += len(sequences[i])
    return out
