"""CPU: host-side containers, synthetic generators, multi-rank gather (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest


def test_synth_deterministic():
    from mauve_py_b200 import synth
    a1, b1 = synth.small_pair(20000, seed=5)
    a2, b2 = synth.small_pair(20000, seed=5)
    assert a1 == a2 and b1 == b2 and a1 != b1
    assert set(a1) <= set(b"ACGT") and set(b1) <= set(b"ACGT")
    x, y = synth.config2_pair(n=200000)
    assert x.size == 200000 and abs(int(y.size) - 200000) < 2000
    x3, y3 = synth.config3_pair(n=300000)
    assert x3.size == 300000 and set(np.unique(y3)) <= set(b"ACGT")
    p = synth.dp_pairs(10, 100, 1000, seed=3)
    assert all(100 <= len(a) <= 1000 for a, _ in p)
    s = synth.hmm_string(1000, seed=2)
    assert len(s) == 1000 and set(s) <= set(b"12345678")


def test_containers():
    import mauve_py_b200 as mp
    m = mp.Match(30, [5, -90])
    assert m.Length() == 30 and m.Start(1) == -90 and m.Orientation(1) == -1 and m.Orientation(0) == 1
    ml = mp.MatchList(seq_table=[b"ACGT", b"ACGT"])
    ml.matches.append(m)
    assert ml.as_array().tolist() == [[30, 5, -90]] and len(ml) == 1
    p = mp.Params.from_array(np.arange(21) / 100.0)
    assert np.array_equal(p.as_array(), np.arange(21) / 100.0)
    with pytest.raises(mp.McuError):
        mp.MemHash().FindMatches(mp.MatchList(seq_table=[b"A", b"C", b"G"]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from mauve_py_b200 import dist as mdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = mdist.TorchComm()
        assert (comm.rank, comm.world) == (rank, world)
        # rank r holds r*3+1 rows (rank 1 of 3 holds none when world == 3 -> exercised by n=0 below)
        n = 0 if (world == 3 and rank == 1) else rank * 3 + 1
        rows = np.arange(n * 3, dtype=np.int64).reshape(n, 3) + 1000 * rank
        out = mdist.gather_rows(rows, comm)
        if rank == 0:
            q.put(out.tolist())
        else:
            assert out is None
        # the three reductions bench.py uses (sum of cells, max of times over ranks)
        assert comm.allreduce([1.0, float(rank)], mdist.SUM) == [float(world), float(sum(range(world)))]
        assert comm.allreduce([float(rank)], mdist.MAX) == [float(world - 1)] and comm.allreduce([float(rank) + 5], mdist.MIN) == [5.0]
        comm.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_rows_gloo(world):
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    expect = []
    for r in range(world):
        n = 0 if (world == 3 and r == 1) else r * 3 + 1
        expect += (np.arange(n * 3).reshape(n, 3) + 1000 * r).tolist()
    assert got == expect


def _shard_worker(rank, world, port, q):
    """DP regions and HMM strings divided among ranks (LPT), results gathered on rank 0: the plumbing of dist.align_sharded /
    dist.hmm_sharded with the CPU restatement standing in for the device calls (tests only)"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import _oracle
    from mauve_py_b200 import dist as mdist, libmems, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.oracle_checker()
        pairs = synth.dp_pairs(23, 30, 400, seed=4)

        def align(ps):
            return [libmems.PWPath(*orc.nw_align(a, b)) for a, b in ps]

        params = orc.hmm_params(0.5, 1e-5, 1e-9, 0.7)
        strings = [synth.hmm_string(50 + 37 * i, seed=i) for i in range(11)] + [b""]

        def run_batch(ss, p):
            return [orc.hmm_run(s, p)[0] if len(s) else b"" for s in ss]

        comm = mdist.TorchComm()
        paths = mdist.align_sharded(pairs, comm, align=align)
        preds = mdist.hmm_sharded(strings, params, comm, run_batch=run_batch)
        if rank == 0:
            q.put(([(p.edges, p.score) for p in paths], preds))
        else:
            assert paths is None and preds is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dp_and_hmm_sharding_gloo(world, orc):
    import torch.multiprocessing as tmp
    from mauve_py_b200 import dist as mdist, synth
    # the plan: every item exactly once, heaviest items spread first, identical on every rank
    costs = [5, 100, 7, 100, 3, 50, 1]
    plan = mdist.lpt_partition(costs, world)
    assert sorted(i for p in plan for i in p) == list(range(len(costs))) and plan == mdist.lpt_partition(costs, world)
    loads = [sum(costs[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(costs)
    assert mdist.lpt_partition([], world) == [[] for _ in range(world)]
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    paths, preds = q.get(timeout=180)
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    pairs = synth.dp_pairs(23, 30, 400, seed=4)
    assert paths == [orc.nw_align(a, b) for a, b in pairs]
    params = orc.hmm_params(0.5, 1e-5, 1e-9, 0.7)
    strings = [synth.hmm_string(50 + 37 * i, seed=i) for i in range(11)] + [b""]
    assert preds == [orc.hmm_run(s, params)[0] if len(s) else b"" for s in strings]


def test_match_list_file_format_round_trips_through_the_reference(refc):
    """--mums / --match-input text format (LM/MatchList.h:526-662): what we write, the reference's ReadList accepts with the
    same rows; what the reference's WriteList prints, we read; both writers agree byte for byte except the match-id column."""
    import ctypes as C
    import io
    import _oracle
    import mauve_py_b200 as mp
    lib = _oracle.ref()
    lib.ref_write_list.restype = C.c_void_p
    lib.ref_write_list.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64]
    lib.ref_read_list.restype = C.c_longlong
    lib.ref_read_list.argtypes = [C.c_char_p, C.POINTER(C.POINTER(_oracle.Match3))]
    rng = np.random.default_rng(3)
    rows = np.stack([rng.integers(21, 5000, 500), rng.integers(1, 4_000_000, 500), rng.integers(1, 4_000_000, 500)], axis=1).astype(np.int64)
    rows[::7, 2] *= -1  # reverse-strand matches carry a negative start
    # ours -> reference
    buf = io.StringIO()
    mp.libmems.WriteList(rows, buf, ("a.fa", "dir with space/b.fa"), (4000000, 4100000))
    out = C.POINTER(_oracle.Match3)()
    n = lib.ref_read_list(buf.getvalue().encode(), C.byref(out))
    assert n == rows.shape[0]
    got = np.array([[out[i].len, out[i].s0, out[i].s1] for i in range(n)], dtype=np.int64)
    lib.ref_free(out)
    assert np.array_equal(got, rows)
    # reference -> ours
    arr = np.ascontiguousarray(rows)
    txt_p = lib.ref_write_list(arr.ctypes.data, rows.shape[0], b"a.fa", b"dir with space/b.fa", 4000000, 4100000)
    txt = C.string_at(txt_p).decode()
    lib.ref_free(txt_p)
    back, names, lens = mp.libmems.ReadList(io.StringIO(txt))
    assert np.array_equal(back, rows) and names == ["a.fa", "dir with space/b.fa"] and lens == [4000000, 4100000]
    strip = lambda t: ["\t".join(l.split("\t")[:3] + l.split("\t")[4:]) if l[:1].isdigit() or l[:1] == "-" else l for l in t.strip().split("\n")]
    assert strip(txt) == strip(buf.getvalue())
    # empty list: the reference writes nothing at all, and rejects an empty file
    e = io.StringIO()
    mp.libmems.WriteList(np.zeros((0, 3), dtype=np.int64), e)
    assert e.getvalue() == "" and lib.ref_read_list(b"", C.byref(out)) == -1
    with pytest.raises(ValueError):
        mp.libmems.ReadList(io.StringIO("FormatVersion\t2\n"))


def test_sml_accessors_against_the_reference_classes(orc):
    """The host-side accessors of the sorted mer list that the reference's callers use besides Read: GetMer, GetSeedMer,
    GetDnaSeedMer (LM/SortedMerList.cpp:321-342, :726-769, RevCompMer :597-614) and FindMer / bsearch (:170-179, :380-394) of the Python
    mirror equal the reference's DNAMemorySML on the same list (oracle/_ref, ref_sml_probe): every probe position, present and absent
    query mers, the rank bsearch stops at included.  The mirror is filled from the oracle's list here (no device on this machine)."""
    import ctypes as C
    import _oracle
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    if not _oracle.have_ref():
        pytest.skip("oracle/_ref/libmauve_ref.so not built (needs /root/reference)")
    ref, lib = _oracle.ref(), _oracle.oracle()
    ref.ref_sml_probe.restype = C.c_longlong
    ref.ref_sml_probe.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64] + [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p] + \
        [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    a, _ = synth.small_pair(30000, seed=12)
    rep = (b"ACGTTGCA" * 400) + a[:3000] + (b"A" * 500) + a[:3000]   # repeats: long equal-mer runs for bsearch to land in
    for seq, (w, r) in ((a, (15, 3)), (a, (11, 0)), (rep, (9, 0)), (a[:40], (5, 0)), (a, (21, 0)), (a, (19, 3))):
        seed = mp.getSeed(w, r)
        pos, mer = orc.sml_build(seq, seed)
        sml = mp.DNAMemorySML()
        sml._pos, sml._mer, sml._seed, sml._length = pos, mer, seed, len(seq)
        sml._packed = np.zeros(int(lib.orc_packed_words(len(seq))), dtype=np.uint32)
        lib.orc_pack(seq, len(seq), sml._packed.ctypes.data)
        probes = np.unique(np.concatenate([rng.integers(0, pos.size, 300), [0, pos.size - 1]])).astype(np.uint64)
        present = mer[rng.integers(0, mer.size, 200)]
        queries = np.concatenate([present, present ^ np.uint64(1), present + np.uint64(1 << 20), rng.integers(0, 2**63, 100).astype(np.uint64) * np.uint64(2),
                                  np.array([0, 2**64 - 1, int(mer[0]), int(mer[-1])], dtype=np.uint64)]).astype(np.uint64)
        found, rank = np.zeros(queries.size, dtype=np.uint8), np.zeros(queries.size, dtype=np.uint64)
        gm, sm, dm = (np.zeros(probes.size, dtype=np.uint64) for _ in range(3))
        n = ref.ref_sml_probe(seq, len(seq), seed, queries.ctypes.data, queries.size, found.ctypes.data, rank.ctypes.data,
                              probes.ctypes.data, probes.size, gm.ctypes.data, sm.ctypes.data, dm.ctypes.data)
        assert n == pos.size
        for k, p in enumerate(probes):
            assert (sml.GetMer(int(p)), sml.GetSeedMer(int(p)), sml.GetDnaSeedMer(int(p))) == (int(gm[k]), int(sm[k]), int(dm[k])), (w, r, int(p))
        # the mer the list holds at a rank is GetDnaSeedMer of its position (MemorySML::operator[], LM/MemorySML.cpp:88-94)
        for i in rng.integers(0, pos.size, 100):
            assert sml.GetDnaSeedMer(int(pos[i])) == int(mer[i])
        for k, q in enumerate(queries):
            ok, at = sml.FindMer(int(q))
            assert (ok, at) == (bool(found[k]), int(rank[k])), (w, r, hex(int(q)))
        c = sml.Clone()
        c._pos[0] ^= 1
        assert sml._pos[0] != c._pos[0] and sml.GetHeader()["seed"] == seed and c.GetHeader()["length"] == len(seq)


def _rendezvous_worker(rank, world, port, tag, stub, q):
    """init_from_env of the product's dist module in a launcher-like environment, with the stand-in library (its communicator exchanges
    files): the NCCL id reaches every rank through the rendezvous file, collectives of the C ABI line up, rank 0 removes the file"""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_PORT=str(port), MCU_RENDEZVOUS_TAG=tag)
    import mauve_py_b200._capi as capi
    capi.LIB_PATH = stub
    from mauve_py_b200 import dist as mdist
    comm = mdist.init_from_env()
    assert (comm.rank, comm.world) == (rank, world)
    assert comm.allreduce([float(rank + 1)], mdist.SUM) == [float(world * (world + 1) // 2)]
    parts = comm.gather_bytes(bytes([65 + rank]) * (rank * 2))   # rank 0 contributes nothing
    if rank == 0:
        q.put([p.decode() for p in parts])
    else:
        assert parts is None
    comm.barrier()
    comm.close()


def test_rendezvous_file_and_library_communicator_two_ranks():
    import multiprocessing as mproc
    import _emu
    from mauve_py_b200 import dist as mdist
    stub = _emu.bench_stub_library()
    ctx = mproc.get_context("spawn")
    q = ctx.Queue()
    port, tag = _free_port(), "pytest%d" % os.getpid()
    procs = [ctx.Process(target=_rendezvous_worker, args=(r, 3, port, tag, stub, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got == ["", "BB", "CCCC"]
    os.environ["MCU_RENDEZVOUS_TAG"] = tag
    try:
        assert not os.path.exists(mdist._id_path(port))     # removed by rank 0 after the first barrier
    finally:
        del os.environ["MCU_RENDEZVOUS_TAG"]
