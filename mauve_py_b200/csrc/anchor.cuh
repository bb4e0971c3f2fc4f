// Device pipeline of the anchoring path: pack -> seedgen -> sort -> join -> extend -> order.
#pragma once
#include "common.cuh"
#include "radix.cuh"

#include <mutex>
#include <vector>

namespace mcu {

struct Session {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_aux = nullptr;  // cross-stream ordering (chunked upload)
    cudaEvent_t kev[12] = {};  // per-kernel marks: 0 start, 1 hist1, 2 scatter1, 3 hist2, 4 scatter2, 5 group, 6/7 candidate, 8 extend
    DevBuf ascii[2], packed[2];
    DevBuf keys_a, keys_b, vals_a, vals_b;
    DevBuf uniq, pairs, cand, raw_matches, ord_keys_a, ord_keys_b, ord_vals_a, ord_vals_b, ord_primary, matches, counters;
    // bucketed enumeration (bucket.cu)
    DevBuf bk_a, bk_b, bk_tab1, bk_tab2, bk_spill, bk_tileseg;
    DevBuf bt_tab, bt_seg_a, bt_seg_b;  // batched gap search (batch.cu)
    u64 bk_spilled = 0, bk_direct = 0, bk_fallbacks = 0;  // fallbacks: runs that overflowed the fixed-capacity layout and were redone exactly
    u64 bk_group_fwd = 0, bk_group_rev = 0;  // leading forward / reverse pairs of the pair list that bk_group emitted
    bool bk_exact = false;  // the last bucketed run used exact (counted) bucket sizes: forced, or after an overflow of the fixed layout
    bool bk_aux = false;  // bucket records carried neighbour bases: pending pairs only need the unique-seed bitmap test
    // bucket replay (replay.cu)
    DevBuf rp_ctr, rp_bitmap, rp_list, rp_canon, rp_keys_b, rp_idx_a, rp_idx_b, rp_p0, rp_row, rp_bkeys, rp_pool, rp_extra, rp_prefix, rp_vinfo, rp_out;
    unsigned long long* h_replay = nullptr;  // pinned, 8 entries
    RadixScratch radix;
    // where the last single-genome sorted list lies (set by sml_build_device) + seed occurrence list / anchor scores (sol.cu)
    const void* sml_keys = nullptr;
    const u32* sml_vals = nullptr;
    u64 sml_npos = 0;
    int sml_key_bytes = 0;
    float sml_ms[3] = {0, 0, 0};  // last sml_build_device: pack, seed generation, radix sort (CUDA events)
    int sml_passes = 0, sml_key_bytes_last = 0;
    DevBuf sol_raw, sol_freq[2], as_rows, as_match, as_off, as_lcb;
    u64 n[2] = {0, 0};
    u64 match_count = 0;
    u64 launches = 0;
    unsigned long long* h_counters = nullptr;  // pinned, 8 entries
    struct RunState {  // carried from mcu_session_enumerate to mcu_session_finish / mcu_session_merge
        SeedParams sp;
        bool enumerated = false, sharded = false, bucketed = false, uniq_global = false;
        int passes = 0;
        u64 nsort = 0, pair_cap = 0, uniq_words = 0;
    } run;
    bool ok = false;
    bool use_buckets = true;  // MAUVE_CUDA_SORT_PATH=1 forces the radix-sort + join enumeration
    // multi-GPU runs (comm.cu): this rank packs words [pack_rank * chunk, (pack_rank + 1) * chunk) of every genome and
    // `after_pack` all-gathers the chunks; a '-' in a rank's slice is reported after the ranks exchanged their flags
    int pack_rank = 0, pack_world = 1;
    int (*after_pack)(Session&) = nullptr;
    bool defer_gap_error = false, gap_seen = false;
    DevBuf gathered, comm_small;            // rank 0: rows of all ranks; per-rank stat vectors
    unsigned long long* h_comm = nullptr;   // pinned, (world + 1) * 8 entries
    // chunked upload (session_upload_begin): the genomes arrive on `copy_stream` in `up_chunks` pieces per genome, one event each;
    // pack and the level-1 scatter of a piece start as soon as it (and the piece after it) is on the device
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> up_events;
    int up_chunks = 0;                      // 0: the whole genomes are resident (session_upload)
    u64 up_chunk_bases[2] = {0, 0};         // bases per piece (a multiple of 4096)
};

// packed words rank r of `world` produces for a genome of n bases: [r * chunk, (r + 1) * chunk), chunk a multiple of 4 words
static inline u64 pack_chunk_words(u64 n, int world)
{
    const u64 words = div_up(n, 16) + 2;
    return (div_up(words, (u64)world) + 3) & ~3ull;
}

int session_init(Session& s);
int default_session(Session** out);  // the process-wide session behind the one-call entry points (api.cu); hold api_mutex()
std::mutex& api_mutex();
void session_destroy(Session& s);
int session_upload(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1);
// H2D of the bases this rank packs in a run sharded over `world` ranks (asynchronous on the session stream)
int session_upload_slice(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1, int rank, int world);
// asynchronous H2D in pieces on the copy stream; the next session_run overlaps pack + level-1 scatter with the copies
int session_upload_begin(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1, int chunks);
int session_run(Session& s, u64 seed, int shard_index, int shard_count, float* stage_ms, u64* stats);
int session_enumerate(Session& s, u64 seed, int shard_index, int shard_count);
int session_finish(Session& s, bool uniq_is_global, float* stage_ms, u64* stats);
int session_merge(Session& s, const mcu_match* rows_dev, u64 n, u64* unclean, u64* dup_rows);
// sorts rows (device, n of them) into reference list order; result in s.matches (device)
int order_matches(Session& s, const mcu_match* rows_dev, u64 n, u64 max_start0);
// join of an already sorted (key, position) array of 64-bit keys: appends to s.uniq / s.pairs / counters (anchor.cu)
int join_sorted_u64(Session& s, const u64* keys, const u32* vals, u64 n, u64 pair_cap);
// bucketed seed-match enumeration (bucket.cu); *used == false when the plan does not apply (caller sorts instead)
int bucket_group(Session& s, const SeedParams& sp, int shard_index, int shard_count, u64 pair_cap, cudaEvent_t ev_scatter1, cudaEvent_t ev_scatter2,
                 bool* used, u64* nrecords);
// Looks for hash buckets whose content depends on the reference's insertion order (replay.cu) and, when `can_replay`
// (genomes + unique-seed bitmap of the whole key space are in the session), replays those buckets exactly.
// s.matches / s.match_count are updated in place.
int replay_unclean(Session& s, const SeedParams* sp, bool can_replay, u64* unclean_buckets, u64* duplicate_rows);

// batched gap search (batch.cu): n_seg sequence pairs given as two concatenations + offsets, one seed for all of them.
// Host outputs: rows (segment-local coordinates, reference list order inside every segment, segments ascending), the
// segment of every row, and the segments whose hash buckets are order dependent (to be redone one by one).
int batch_find_mums(Session& s, const char* cat0, const u64* off0, const char* cat1, const u64* off1, u32 n_seg, u64 seed,
                    std::vector<mcu_match>* rows, std::vector<u32>* row_seg, std::vector<u32>* unclean_segs, u64* seed_pairs);
int run_pack_genome(Session& s, int g, u32* err_flag);
int run_pack_piece(Session& s, int g, int c, u32* err_flag, u64* bases_ready);
// true when bucket_group will take the fixed-capacity bucketed path for these sizes (decided from sizes alone)
bool bucket_plan_applies(const SeedParams& sp, u64 npos0, u64 npos1, int shard_count);

// single-genome SML (stable): outputs on device in s.keys_*/vals_* ; returns which buffer
int sml_build_device(Session& s, const char* seq, u64 n, u64 seed, u32* pos_out, u64* mer_out, u32* packed_out, u64* len_out, int shard = 0, int nshard = 1);

#ifdef __CUDACC__
// ---- probing one diagonal of the two packed genomes (shared by join/extend and the bucket replay) ----
struct ExtendArgs {
    const u32* g0;
    const u32* g1;
    u64 npos0, npos1;
    const u32* uniq;
    u64* cand;          // candidate list p0 | p1 << 32: forward-strand from the front, reverse-strand from the back
    u64 nfwd, nrev, cap;
    mcu_match* out;
    unsigned long long* counters;
};

// valid seed start positions of the two sequences being compared: [lo0, hi0) in genome 0, [lo1, hi1) in genome 1
// (the whole genomes, or one segment pair of a batched gap search)
struct DiagBounds {
    i64 lo0, hi0, lo1, hi1;
};

__device__ __forceinline__ bool probe_hit_in(const ExtendArgs& a, const SeedParams& sp, bool rev, i64 d, i64 t, i64& other, const DiagBounds& b)
{
    if (t < b.lo0 || t >= b.hi0) return false;
    other = rev ? d - t : t + d;
    if (other < b.lo1 || other >= b.hi1) return false;
    u64 f0 = extract_seed(load_mer32(a.g0, (u64)t), sp);
    u64 x1 = extract_seed(load_mer32(a.g1, (u64)other), sp);
    if (!rev) return f0 == x1;
    if (f0 != revcomp_seed(x1, sp.w)) return false;
    return f0 != revcomp_seed(f0, sp.w);
}

__device__ __forceinline__ bool probe_hit(const ExtendArgs& a, const SeedParams& sp, bool rev, i64 d, i64 t, i64& other)
{
    if (t < 0 || t >= (i64)a.npos0) return false;
    other = rev ? d - t : t + d;
    if (other < 0 || other >= (i64)a.npos1) return false;
    u64 f0 = extract_seed(load_mer32(a.g0, (u64)t), sp);
    u64 x1 = extract_seed(load_mer32(a.g1, (u64)other), sp);
    if (!rev) return f0 == x1;
    if (f0 != revcomp_seed(x1, sp.w)) return false;
    return f0 != revcomp_seed(f0, sp.w);
}

__device__ __forceinline__ bool uniq_bit(const u32* __restrict__ uniq, i64 t) { return (__ldg(uniq + (t >> 5)) >> (t & 31)) & 1u; }
#endif

}  // namespace mcu

struct mcu_session { mcu::Session s; };
