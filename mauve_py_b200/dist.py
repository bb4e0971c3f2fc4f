"""Multi-GPU plumbing of the anchoring path: one process per GPU.

The data path lives in the library (csrc/comm.cu): NCCL is called from C++ on the session's stream -- the sharded
seed + match + extend step is ONE call (`AnchorSession.run_sharded` -> mcu_session_run_sharded) with no Python between its
phases.  What is left here is start-up and the division of independent work:

* `init_from_env()`     rank / world / device from the torchrun environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_PORT), the
                        128-byte NCCL id travels from rank 0 to the others through a file (one node), then mcu_comm_init.
                        No torch anywhere on this path.
* `NcclComm`            thin face of mcu_comm_* (barrier, all-reduce of host doubles, gather of host bytes to rank 0)
* `TorchComm`           the same three operations over torch.distributed (gloo): lets the CPU test-suite run the host logic below
                        with two and three processes; torch is imported only there
* `lpt_partition`, `align_sharded`, `hmm_sharded`   gapped DP regions and HMM strings are independent: longest-processing-time-first
                        division among the ranks, no collective on the data path, results gathered on rank 0
"""
import os
import time

import numpy as np

from . import _capi
from ._capi import check, lib

SUM, MAX, MIN = 0, 1, 2


# ---- start-up ---------------------------------------------------------------------------------------------------------------------
def _id_path(port):
    # every rank of one launch is a child of the same launcher process (torchrun's agent), so its pid names the launch
    tag = os.environ.get("MCU_RENDEZVOUS_TAG") or "%d" % os.getppid()
    return os.path.join(os.environ.get("MCU_RENDEZVOUS_DIR", "/tmp"), "mcu_nccl_id_%s_%s" % (port, tag))


def exchange_id(rank, world, make_id, timeout=300.0):
    """rank 0 makes the id and publishes it (write + atomic rename); the others poll for the file"""
    path = _id_path(os.environ.get("MASTER_PORT", "0"))
    if rank == 0:
        blob = make_id()
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(blob)
        os.replace(tmp, path)
        return blob, path
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                blob = f.read()
            if len(blob) == _capi.COMM_ID_BYTES:
                return blob, path
        except OSError:
            pass
        if time.time() - t0 > timeout:
            raise TimeoutError("rank %d: no NCCL id at %s after %.0f s" % (rank, path, timeout))
        time.sleep(0.01)


class NcclComm:
    """the library's own NCCL communicator (csrc/comm.cu); one per process"""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def barrier(self):
        check(lib().mcu_comm_barrier())

    def allreduce(self, values, op=SUM):
        """element-wise reduction of a list of floats over the ranks -> list of floats"""
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        check(lib().mcu_comm_allreduce_f64(v.ctypes.data, int(v.size), int(op)))
        return v.tolist()

    def gather_bytes(self, payload: bytes):
        """one byte string per rank -> list in rank order on rank 0, None elsewhere"""
        import ctypes as C
        out = C.c_void_p()
        counts = np.zeros(self.world, dtype=np.uint64)
        buf = np.frombuffer(payload, dtype=np.uint8)
        check(lib().mcu_comm_gather_bytes(buf.ctypes.data if buf.size else None, int(buf.size), C.byref(out), counts.ctypes.data))
        if self.rank != 0:
            return None
        total = int(counts.sum())
        blob = C.string_at(out, total) if total else b""
        lib().mcu_free(out)
        parts, pos = [], 0
        for c in counts.tolist():
            parts.append(blob[pos:pos + int(c)])
            pos += int(c)
        return parts

    def close(self):
        lib().mcu_comm_destroy()


def init_from_env():
    """mcu_init(LOCAL_RANK) + mcu_comm_init for the process layout torchrun (or any launcher setting the same variables) gives.
    Returns an NcclComm; with WORLD_SIZE absent or 1 it is a one-rank communicator and no NCCL call is made."""
    import ctypes as C
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    L = lib()
    check(L.mcu_init(local))
    if world == 1:
        check(L.mcu_comm_init(0, 1, None))
        return NcclComm(0, 1)

    def make_id():
        buf = C.create_string_buffer(_capi.COMM_ID_BYTES)
        check(L.mcu_comm_unique_id(buf))
        return buf.raw

    blob, path = exchange_id(rank, world, make_id)
    check(L.mcu_comm_init(rank, world, blob))
    comm = NcclComm(rank, world)
    comm.barrier()
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return comm


class TorchComm:
    """barrier / allreduce / gather_bytes over an initialised torch.distributed group (gloo on CPU): test plumbing for the host
    logic of this module; the product path uses NcclComm"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self._group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def barrier(self):
        self._dist.barrier(group=self._group)

    def allreduce(self, values, op=SUM):
        import torch
        t = torch.tensor(list(values), dtype=torch.float64)
        ops = {SUM: self._dist.ReduceOp.SUM, MAX: self._dist.ReduceOp.MAX, MIN: self._dist.ReduceOp.MIN}
        self._dist.all_reduce(t, op=ops[op], group=self._group)
        return t.tolist()

    def gather_bytes(self, payload: bytes):
        import torch
        dist = self._dist
        buf = torch.from_numpy(np.frombuffer(payload, dtype=np.uint8).copy())
        n = torch.tensor([buf.numel()], dtype=torch.int64)
        counts = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(counts, n, group=self._group)
        counts = torch.cat(counts).tolist()
        if self.rank == 0:
            parts = [torch.empty(c, dtype=torch.uint8) for c in counts]
            parts[0] = buf
            reqs = [dist.irecv(parts[r], src=r, group=self._group) for r in range(self.world) if r != 0 and counts[r] > 0]
            for q in reqs:
                q.wait()
            return [p.numpy().tobytes() for p in parts]
        if buf.numel() > 0:
            dist.send(buf, dst=0, group=self._group)
        return None


def gather_rows(rows, comm):
    """rows: int64 [n, 3] array per rank -> concatenation in rank order on rank 0 (None elsewhere).  Host-side helper for callers
    that finished shards independently (mcu_merge_matches); the sharded step itself gathers on the device (comm.cu)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    parts = comm.gather_bytes(rows.tobytes())
    if parts is None:
        return None
    return np.frombuffer(b"".join(parts), dtype=np.int64).reshape(-1, 3).copy()


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


def align_sharded(pairs, comm, align=None):
    """muscle::GlobalAlign for a batch of region pairs divided among the ranks by LPT on lenA * lenB.  Every rank holds the
    whole `pairs` list (like the genomes), aligns its share on its GPU and ships (lengths, scores, edges) to rank 0, which returns
    the PWPaths in input order; other ranks return None.  `align` defaults to libmems.GlobalAlignBatch."""
    from . import libmems
    align = align or libmems.GlobalAlignBatch
    rank, world = comm.rank, comm.world
    plan = lpt_partition([len(a) * len(b) for a, b in pairs], world)
    mine = plan[rank]
    paths = align([pairs[i] for i in mine]) if mine else []
    if world == 1:
        return paths
    lens = np.array([len(p.edges) for p in paths], dtype=np.int64)
    scores = np.array([p.score for p in paths], dtype=np.int64)
    got = comm.gather_bytes(lens.tobytes() + scores.tobytes() + b"".join(p.edges for p in paths))
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


def hmm_sharded(sequences, params, comm, run_batch=None):
    """run() of the HomologyHMM for a batch of column strings divided among the ranks by LPT on their lengths; rank 0 returns the
    H/N predictions in input order, other ranks None.  `run_batch` defaults to libmems.run_batch."""
    from . import libmems
    run_batch = run_batch or libmems.run_batch
    rank, world = comm.rank, comm.world
    plan = lpt_partition([len(s) for s in sequences], world)
    mine = plan[rank]
    preds = run_batch([sequences[i] for i in mine], params) if mine else []
    if world == 1:
        return preds
    got = comm.gather_bytes(b"".join(preds))
    if rank != 0:
        return None
    out = [None] * len(sequences)
    for r, blob in enumerate(got):
        pos = 0
        for i in plan[r]:
            out[i] = blob[pos:pos + len(sequences[i])]
            pos += len(sequences[i])
    return out
