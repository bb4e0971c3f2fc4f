// Multi-GPU layer of libmauve_cuda.so: one process per GPU, NCCL over NVLink / NVSwitch, called directly from
// this library -- no Python, no torch between the phases of a step.  Precedent for the partition in the reference:
// ParallelMemHash chunks the sorted mer list by mer range (LM/ParallelMemHash.cpp:63-101); here every seed is owned by the
// rank its mer hashes to (seed_owned, common.cuh).
//
// One sharded step (mcu_session_run_sharded), everything enqueued on the session's stream:
//   pack            each rank packs 1/world of both genomes (2-bit), ncclAllGather of the packed words (25 MB per 100 Mbp)
//   enumerate       bucketed seed-match enumeration of the seeds this rank owns -> unique-seed bitmap + pair list
//   ncclAllReduce   SUM of the bitmaps (the ranks' bits are disjoint, so SUM == OR): a match is emitted by its leftmost unique
//                   seed whichever rank owns it, and extension stops at unique seeds of ANY rank
//   extend          candidates + extension of this rank's pairs -> this rank's rows (every match has exactly one emitter)
//   ncclAllGather   of 8 x u64 per rank: row count, pair / candidate / record counts, '-' flag
//   ncclSend/Recv   one group: the rows go to rank 0 (the "NCCL gather" of the north star)
//   merge           rank 0: reference list order + exact replay of order-dependent hash buckets (replay.cu)
// Host synchronisations are the ones a single-GPU step has (counter read-backs that size the next launch) plus one for the counts.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is opened at run time (see load_nccl)

#include <mutex>
#include <new>
#include <vector>

#include "anchor.cuh"

namespace mcu {

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool ok = false;
    cudaStream_t stream = nullptr;  // utility collectives (barrier, host all-reduce / gather)
    DevBuf scratch;
};
static Comm g_comm;

// NCCL is bound with dlopen when the first communicator call is made, not at link time: libmauve_cuda.so must stay loadable in a
// process that brings its own libnccl.so.2 (PyTorch bundles a newer one than the system's; two different libnccl.so.2 cannot share
// a process, and whichever is mapped first wins the soname).  Order: $MAUVE_CUDA_NCCL_LIB, a copy that is already mapped, the
// system's libnccl.so.2.
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.handle) return MCU_OK;
    void* h = nullptr;
    if (const char* e = getenv("MAUVE_CUDA_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2 (%s); set MAUVE_CUDA_NCCL_LIB", dlerror()); return MCU_ENODEV; }
    NcclApi a;
    a.handle = h;
#define MCU_SYM(field, name)                                                           \
    a.field = (decltype(a.field))dlsym(h, name);                                      \
    if (!a.field) { set_error("libnccl.so.2 has no %s", name); return MCU_ENODEV; }
    MCU_SYM(GetUniqueId, "ncclGetUniqueId")
    MCU_SYM(CommInitRank, "ncclCommInitRank")
    MCU_SYM(CommDestroy, "ncclCommDestroy")
    MCU_SYM(GetErrorString, "ncclGetErrorString")
    MCU_SYM(AllReduce, "ncclAllReduce")
    MCU_SYM(AllGather, "ncclAllGather")
    MCU_SYM(Send, "ncclSend")
    MCU_SYM(Recv, "ncclRecv")
    MCU_SYM(GroupStart, "ncclGroupStart")
    MCU_SYM(GroupEnd, "ncclGroupEnd")
    MCU_SYM(GetVersion, "ncclGetVersion")
#undef MCU_SYM
    g_nccl = a;
    return MCU_OK;
}

#define MCU_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            mcu::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return MCU_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

static int comm_ready()
{
    if (!g_comm.ok) { set_error("mcu_comm_init has not been called"); return MCU_EINVAL; }
    return ensure_device();
}

// the other ranks' chunks of the packed genomes (in place: this rank's chunk already lies at its offset)
static int allgather_packed(Session& s)
{
    for (int g = 0; g < 2; ++g) {
        const u64 chunk = pack_chunk_words(s.n[g], s.pack_world);
        u32* base = s.packed[g].as<u32>();
        MCU_NCCL(g_nccl.AllGather(base + chunk * (u64)s.pack_rank, base, chunk, ncclUint32, g_comm.comm, s.stream));
    }
    return MCU_OK;
}

static int run_sharded(Session& s, u64 seed, float* stage_ms, u64* stats)
{
    Comm& c = g_comm;
    if (c.world == 1) return session_run(s, seed, 0, 1, stage_ms, stats);
    const int W = c.world;
    if (!s.h_comm) MCU_CUDA(cudaHostAlloc((void**)&s.h_comm, (size_t)(W + 1) * 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    MCU_TRY(s.comm_small.reserve((size_t)(W + 1) * 8 * sizeof(unsigned long long)));

    // ---- pack (1 / world per rank) + all-gather, enumeration of this rank's seeds ----
    s.pack_rank = c.rank;
    s.pack_world = W;
    s.after_pack = allgather_packed;
    s.defer_gap_error = true;
    int r = session_enumerate(s, seed, c.rank, W);
    s.pack_world = 1;
    s.pack_rank = 0;
    s.after_pack = nullptr;
    s.defer_gap_error = false;
    MCU_TRY(r);

    // ---- unique-seed bitmaps of all ranks ----
    MCU_NCCL(g_nccl.AllReduce(s.uniq.p, s.uniq.p, s.run.uniq_words, ncclUint32, ncclSum, c.comm, s.stream));

    // ---- candidates + extension of this rank's pairs ----
    float st[16];
    u64 mine[8];
    MCU_TRY(session_finish(s, true, st, mine));
    const u64 nmine = s.match_count;

    // ---- counts ----
    unsigned long long* h_send = s.h_comm;
    unsigned long long* h_all = s.h_comm + 8;
    unsigned long long* d_send = s.comm_small.as<unsigned long long>();
    unsigned long long* d_all = d_send + 8;
    h_send[0] = nmine; h_send[1] = mine[0]; h_send[2] = mine[4]; h_send[3] = mine[5]; h_send[4] = mine[3];
    h_send[5] = s.gap_seen ? 1 : 0; h_send[6] = 0; h_send[7] = 0;
    MCU_CUDA(cudaMemcpyAsync(d_send, h_send, 64, cudaMemcpyHostToDevice, s.stream));
    MCU_NCCL(g_nccl.AllGather(d_send, d_all, 8, ncclUint64, c.comm, s.stream));
    MCU_CUDA(cudaMemcpyAsync(h_all, d_all, (size_t)W * 64, cudaMemcpyDeviceToHost, s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    u64 total = 0, pairs = 0, cands = 0, recs = 0, repeat = 0, gap = 0;
    for (int k = 0; k < W; ++k) {
        total += h_all[8 * k]; pairs += h_all[8 * k + 1]; cands += h_all[8 * k + 2]; recs += h_all[8 * k + 3];
        repeat |= h_all[8 * k + 4]; gap |= h_all[8 * k + 5];
    }
    if (gap) { set_error("gap character '-' in a genome sequence (input must be unaligned)"); return MCU_EGAP; }

    // ---- rows to rank 0 ----
    if (c.rank == 0) MCU_TRY(s.gathered.reserve((total + 1) * sizeof(mcu_match)));
    MCU_NCCL(g_nccl.GroupStart());
    if (c.rank == 0) {
        u64 off = h_all[0];
        for (int k = 1; k < W; ++k) {
            const u64 cnt = h_all[8 * k];
            if (cnt) MCU_NCCL(g_nccl.Recv(s.gathered.as<mcu_match>() + off, cnt * 3, ncclInt64, k, c.comm, s.stream));
            off += cnt;
        }
    } else if (nmine)
        MCU_NCCL(g_nccl.Send(s.matches.p, nmine * 3, ncclInt64, 0, c.comm, s.stream));
    MCU_NCCL(g_nccl.GroupEnd());

    // ---- merge on rank 0 ----
    u64 unclean = 0, dups = 0;
    if (c.rank == 0) {
        if (nmine) MCU_CUDA(cudaMemcpyAsync(s.gathered.p, s.matches.p, nmine * sizeof(mcu_match), cudaMemcpyDeviceToDevice, s.stream));
        MCU_TRY(session_merge(s, s.gathered.as<mcu_match>(), total, &unclean, &dups));
    }
    MCU_CUDA(cudaEventRecord(s.ev[7], s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    MCU_CUDA(cudaGetLastError());
    if (stage_ms) {
        for (int i = 0; i < 16; ++i) stage_ms[i] = st[i];
        cudaEventElapsedTime(&stage_ms[6], s.ev[0], s.ev[7]);  // the whole step on this rank, collectives and merge included
    }
    if (stats) {
        const u64 nfinal = c.rank == 0 ? s.match_count : total;
        stats[0] = pairs; stats[1] = nfinal; stats[2] = pairs - nfinal; stats[3] = repeat; stats[4] = cands; stats[5] = recs;
        stats[6] = unclean; stats[7] = dups;
    }
    return MCU_OK;
}

}  // namespace mcu

using namespace mcu;

extern "C" {

int mcu_comm_unique_id(void* id_out)
{
    if (!id_out) return MCU_EINVAL;
    MCU_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == MCU_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    MCU_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return MCU_OK;
}

int mcu_comm_init(int rank, int world, const void* id)
{
    MCU_TRY(ensure_device());
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) { set_error("mcu_comm_init: bad rank / world / id"); return MCU_EINVAL; }
    if (g_comm.ok) { set_error("mcu_comm_init: already initialised"); return MCU_EINVAL; }
    g_comm.rank = rank;
    g_comm.world = world;
    if (world > 1) {
        MCU_TRY(load_nccl());
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        MCU_NCCL(g_nccl.CommInitRank(&g_comm.comm, world, uid, rank));
        MCU_CUDA(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
    }
    g_comm.ok = true;
    return MCU_OK;
}

void mcu_comm_destroy(void)
{
    if (!g_comm.ok) return;
    if (g_comm.comm) {
        if (g_comm.stream) cudaStreamSynchronize(g_comm.stream);
        g_nccl.CommDestroy(g_comm.comm);
        g_comm.comm = nullptr;
    }
    if (g_comm.stream) { cudaStreamDestroy(g_comm.stream); g_comm.stream = nullptr; }
    g_comm.scratch.release();
    g_comm.ok = false;
    g_comm.rank = 0;
    g_comm.world = 1;
}

int mcu_comm_rank(void) { return g_comm.ok ? g_comm.rank : 0; }
int mcu_comm_world(void) { return g_comm.ok ? g_comm.world : 1; }

int mcu_device_synchronize(void)
{
    MCU_TRY(ensure_device());
    MCU_CUDA(cudaDeviceSynchronize());
    return MCU_OK;
}

int mcu_comm_allreduce_f64(double* v, int n, int op)
{
    MCU_TRY(comm_ready());
    if (n < 0 || (n && !v) || op < 0 || op > 2) return MCU_EINVAL;
    if (g_comm.world == 1 || n == 0) return MCU_OK;
    const ncclRedOp_t ops[3] = {ncclSum, ncclMax, ncclMin};
    MCU_TRY(g_comm.scratch.reserve((size_t)n * sizeof(double)));
    MCU_CUDA(cudaMemcpyAsync(g_comm.scratch.p, v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g_comm.stream));
    MCU_NCCL(g_nccl.AllReduce(g_comm.scratch.p, g_comm.scratch.p, (size_t)n, ncclDouble, ops[op], g_comm.comm, g_comm.stream));
    MCU_CUDA(cudaMemcpyAsync(v, g_comm.scratch.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g_comm.stream));
    MCU_CUDA(cudaStreamSynchronize(g_comm.stream));
    return MCU_OK;
}

int mcu_comm_barrier(void)
{
    MCU_TRY(comm_ready());
    MCU_CUDA(cudaDeviceSynchronize());
    double one = 1.0;
    MCU_TRY(mcu_comm_allreduce_f64(&one, 1, 0));
    return MCU_OK;
}

int mcu_comm_gather_bytes(const void* send, uint64_t n, void** out, uint64_t* counts_out)
{
    MCU_TRY(comm_ready());
    if ((n && !send) || !out) return MCU_EINVAL;
    const int W = g_comm.world, rank = g_comm.rank;
    *out = nullptr;
    if (W == 1) {
        void* r = malloc(n ? n : 1);
        if (!r) return MCU_ENOMEM;
        if (n) memcpy(r, send, n);
        if (counts_out) counts_out[0] = n;
        *out = r;
        return MCU_OK;
    }
    cudaStream_t st = g_comm.stream;
    // sizes
    std::vector<unsigned long long> cnt((size_t)W);
    MCU_TRY(g_comm.scratch.reserve((size_t)(W + 1) * 8 + 16));
    unsigned long long* d = g_comm.scratch.as<unsigned long long>();
    unsigned long long mine = n;
    MCU_CUDA(cudaMemcpyAsync(d, &mine, 8, cudaMemcpyHostToDevice, st));
    MCU_NCCL(g_nccl.AllGather(d, d + 1, 1, ncclUint64, g_comm.comm, st));
    MCU_CUDA(cudaMemcpyAsync(cnt.data(), d + 1, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    u64 total = 0;
    for (int k = 0; k < W; ++k) total += cnt[k];
    if (counts_out)
        for (int k = 0; k < W; ++k) counts_out[k] = cnt[k];
    // payloads through device staging
    DevBuf stage;
    MCU_TRY(stage.reserve((rank == 0 ? total : n) + 16));
    u64 off0 = 0;
    if (n) MCU_CUDA(cudaMemcpyAsync(stage.p, send, n, cudaMemcpyHostToDevice, st));
    MCU_NCCL(g_nccl.GroupStart());
    if (rank == 0) {
        u64 off = cnt[0];
        for (int k = 1; k < W; ++k) {
            if (cnt[k]) MCU_NCCL(g_nccl.Recv(stage.as<char>() + off, cnt[k], ncclUint8, k, g_comm.comm, st));
            off += cnt[k];
        }
    } else if (n)
        MCU_NCCL(g_nccl.Send(stage.as<char>() + off0, n, ncclUint8, 0, g_comm.comm, st));
    MCU_NCCL(g_nccl.GroupEnd());
    int rc = MCU_OK;
    if (rank == 0) {
        void* r = malloc(total ? total : 1);
        if (!r) rc = MCU_ENOMEM;
        else {
            if (total && cudaMemcpyAsync(r, stage.p, total, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = MCU_ECUDA;
            *out = r;
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = MCU_ECUDA;
    stage.release();
    if (rc != MCU_OK && *out) { free(*out); *out = nullptr; }
    if (rc == MCU_ECUDA) set_error("mcu_comm_gather_bytes: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int mcu_session_upload_sharded(mcu_session* h, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(comm_ready());
    if (g_comm.world == 1) return session_upload(h->s, seq0, n0, seq1, n1);
    return session_upload_slice(h->s, seq0, n0, seq1, n1, g_comm.rank, g_comm.world);
}

int mcu_session_run_sharded(mcu_session* h, uint64_t seed, float* stage_ms, uint64_t* stats)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(comm_ready());
    return run_sharded(h->s, seed, stage_ms, stats);
}

int mcu_find_mums_sharded(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match* rows_out,
                          uint64_t cap, uint64_t* n_out, uint64_t* stats)
{
    std::lock_guard<std::mutex> lk(api_mutex());
    MCU_TRY(comm_ready());
    if (!n_out || (cap && !rows_out)) { set_error("mcu_find_mums_sharded: NULL output pointer"); return MCU_EINVAL; }
    if (rule != MCU_RULE_PAIRWISE && rule != MCU_RULE_MEMHASH) { set_error("mcu_find_mums_sharded: unknown rule %d", rule); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    if (g_comm.world == 1) MCU_TRY(session_upload(*s, seq0, n0, seq1, n1));
    else MCU_TRY(session_upload_slice(*s, seq0, n0, seq1, n1, g_comm.rank, g_comm.world));
    u64 st[8];
    MCU_TRY(run_sharded(*s, seed, nullptr, st));
    if (stats) memcpy(stats, st, sizeof st);
    *n_out = st[1];
    if (g_comm.rank != 0) return MCU_OK;
    const u64 m = s->match_count;
    if (m > cap) { set_error("mcu_find_mums_sharded: %llu rows do not fit the caller's %llu", (unsigned long long)m, (unsigned long long)cap); return MCU_ESMALL; }
    if (m) {
        MCU_CUDA(cudaMemcpyAsync(rows_out, s->matches.p, m * sizeof(mcu_match), cudaMemcpyDeviceToHost, s->stream));
        MCU_CUDA(cudaStreamSynchronize(s->stream));
    }
    return MCU_OK;
}

// Sorted-mer-list build sharded by key range (the materialised position array): every rank scans the genome, keeps the
// seeds of its range (anchor.cu SmlSplit), sorts them; the ranks' lists, one after the other, are the sorted list.  One
// ncclAllGather of the lengths, one group of ncclSend / ncclRecv of the 4-byte positions to rank 0.  pos_out (rank 0): n - L + 1
// positions; equal mers come in unspecified order inside their run (the reference's own order there is std::sort's).
int mcu_sml_build_sharded(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* sml_len_out, float* ms_out)
{
    std::lock_guard<std::mutex> lk(api_mutex());
    MCU_TRY(comm_ready());
    if (n && !seq) { set_error("mcu_sml_build_sharded: NULL sequence"); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    Comm& c = g_comm;
    const int W = c.world;
    u64 mine = 0;
    cudaEvent_t e0 = s->ev[0], e1 = s->kev[11];   // ev[0]: recorded by the build once the genome is in HBM; (kev 9..11 are not used by the enumeration)
    if (W == 1) {
        MCU_TRY(sml_build_device(*s, seq, n, seed, pos_out, nullptr, nullptr, &mine));
        if (sml_len_out) *sml_len_out = mine;
        if (ms_out) cudaEventElapsedTime(ms_out, e0, s->ev[3]);   // ev[3]: the sort is done (the copy of the positions to the host follows it)
        return MCU_OK;
    }
    MCU_TRY(sml_build_device(*s, seq, n, seed, nullptr, nullptr, nullptr, &mine, c.rank, W));
    if (!s->h_comm) MCU_CUDA(cudaHostAlloc((void**)&s->h_comm, (size_t)(W + 1) * 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    MCU_TRY(s->comm_small.reserve((size_t)(W + 1) * 8 * sizeof(unsigned long long)));
    unsigned long long* h_send = s->h_comm;
    unsigned long long* h_all = s->h_comm + 8;
    unsigned long long* d_send = s->comm_small.as<unsigned long long>();
    unsigned long long* d_all = d_send + 8;
    h_send[0] = mine;
    MCU_CUDA(cudaMemcpyAsync(d_send, h_send, 8, cudaMemcpyHostToDevice, s->stream));
    MCU_NCCL(g_nccl.AllGather(d_send, d_all, 1, ncclUint64, c.comm, s->stream));
    MCU_CUDA(cudaMemcpyAsync(h_all, d_all, (size_t)W * 8, cudaMemcpyDeviceToHost, s->stream));
    MCU_CUDA(cudaStreamSynchronize(s->stream));
    u64 total = 0;
    for (int k = 0; k < W; ++k) total += h_all[k];
    if (c.rank == 0) MCU_TRY(s->gathered.reserve((total + 1) * sizeof(u32)));
    MCU_NCCL(g_nccl.GroupStart());
    if (c.rank == 0) {
        u64 off = h_all[0];
        for (int k = 1; k < W; ++k) {
            if (h_all[k]) MCU_NCCL(g_nccl.Recv(s->gathered.as<u32>() + off, h_all[k], ncclUint32, k, c.comm, s->stream));
            off += h_all[k];
        }
    } else if (mine)
        MCU_NCCL(g_nccl.Send(s->sml_vals, mine, ncclUint32, 0, c.comm, s->stream));
    MCU_NCCL(g_nccl.GroupEnd());
    if (c.rank == 0) {
        if (mine) MCU_CUDA(cudaMemcpyAsync(s->gathered.p, s->sml_vals, mine * sizeof(u32), cudaMemcpyDeviceToDevice, s->stream));
        MCU_CUDA(cudaEventRecord(e1, s->stream));
        if (pos_out && total) MCU_CUDA(cudaMemcpyAsync(pos_out, s->gathered.p, total * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
    } else
        MCU_CUDA(cudaEventRecord(e1, s->stream));
    MCU_CUDA(cudaStreamSynchronize(s->stream));
    MCU_CUDA(cudaGetLastError());
    if (sml_len_out) *sml_len_out = total;
    if (ms_out) cudaEventElapsedTime(ms_out, e0, e1);
    return MCU_OK;
}

}  // extern "C"
