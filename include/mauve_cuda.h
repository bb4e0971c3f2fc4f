/*
 * mauve_cuda.h -- C ABI of libmauve_cuda.so: the B200 (sm_100a) implementation of
 * progressiveMauve's anchoring hot path.
 *
 * The reference (Wyss/mauve-py, mauve/src/libMems) has no FFI today: everything on this
 * path is a C++ virtual call inside one process.  Each entry point below states the
 * reference interface it replaces (paths relative to /root/reference/mauve/src/;
 * LM = libMems/libMems, MU = muscle/libMUSCLE).  INTEGRATION.md shows the C++ adapter
 * classes a maintainer adds on the reference side to bind them.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0
 * on success and a negative MCU_E* code on failure; mcu_last_error() gives the text for
 * the calling thread.  Host buffers are caller-owned; buffers returned through `**out`
 * are library-owned until mcu_free().  There is NO CPU fallback: if no CUDA device is
 * usable every compute entry point fails with MCU_ENODEV.
 */
#ifndef MAUVE_CUDA_H_
#define MAUVE_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MCU_OK 0
#define MCU_ENODEV (-1)   /* no CUDA device / driver */
#define MCU_ECUDA (-2)    /* CUDA runtime error (see mcu_last_error) */
#define MCU_EINVAL (-3)   /* bad argument (seed pattern, sizes, NULL pointer) */
#define MCU_EGAP (-4)     /* '-' in a genome sequence: the reference throws (LM/SortedMerList.cpp:433-437) */
#define MCU_ENOMEM (-5)
#define MCU_EALPHA (-6)   /* DP input outside ACGT (integer-exact kernel only) */
#define MCU_ESMALL (-7)   /* caller-provided output buffer too small; the required size is reported */

/* One ungapped match, exactly the fields of a reference match-list row
 * (LM/MatchList.h:617-662 WriteList: length, start0, start1; 1-based, start1 < 0 = reverse strand). */
typedef struct mcu_match {
    int64_t len;
    int64_t start0;
    int64_t start1;
} mcu_match;

/* ---- lifecycle ------------------------------------------------------------------------ */
/* Selects the CUDA device for the calling process (one process per GPU). */
int mcu_init(int device);
void mcu_shutdown(void);
const char* mcu_last_error(void);
void mcu_free(void* p);
/* Pinned host buffers for callers that want overlap-capable transfers. */
int mcu_host_alloc(void** out, uint64_t bytes);
void mcu_host_free(void* p);

/* ---- seed patterns: LM/SeedMasks.h:298-321 getSeed, :389-401 getDefaultSeedWeight,
 *      :335-373 getSeedLength/getSeedWeight (host-side table logic, no device work) -------- */
#define MCU_SOLID_SEED 0x7fffffff
#define MCU_CODING_SEED 3
uint64_t mcu_get_seed(int weight, int seed_rank);
unsigned mcu_default_seed_weight(uint64_t avg_sequence_length);
int mcu_seed_length(uint64_t seed);
int mcu_seed_weight(uint64_t seed);

/* ---- sorted mer list: replaces mems::DNAMemorySML::Create (LM/MemorySML.cpp:45-60:
 *      SortedMerList::Create + FillDnaSeedSML/FillDnaSML + std::sort(bmer_lessthan)) and the
 *      data MemorySML::Read / operator[] recompute (LM/MemorySML.cpp:62-94).
 *      seq: n ASCII bases.  Outputs (each optional, may be NULL):
 *        pos_out[n-L+1]    positions in sorted order (ties position-ascending)
 *        mer_out[n-L+1]    bmer::mer of each rank (canonical seed left-aligned | strand bit)
 *        packed_out[ceil(n/16)+2]  2-bit sequence exactly as SortedMerList::sequence
 *      *sml_len_out = n-L+1 (0 when n < L).                                                  */
int mcu_sml_build(const char* seq, uint64_t n, uint64_t seed,
                  uint32_t* pos_out, uint64_t* mer_out, uint32_t* packed_out, uint64_t* sml_len_out);

/* One shard of the sorted mer list (the materialised position array sharded by mer range): the seeds whose key
 * (canonical mer) lies in range `shard` of `n_shards` (<= 16) ranges that tile the key space in order, so that the shards' lists,
 * one after the other, are the list mcu_sml_build returns -- except that equal mers come in unspecified order inside their run
 * (the reference's own order there is std::sort's).  *shard_len_out = this shard's length; pos_out / mer_out need that much room
 * (n - L + 1 always suffices).  mcu_sml_build_sharded is the collective form over the communicator of mcu_comm_init: every rank
 * passes the same sequence and builds its shard, the positions are gathered on rank 0 (pos_out, n - L + 1 entries, used there only);
 * ms_out (optional): device ms on this rank from the genome being in HBM to the gathered list being in rank 0's HBM (pack, scan,
 * sort, gather; CUDA events). */
int mcu_sml_build_shard(const char* seq, uint64_t n, uint64_t seed, int shard, int n_shards, uint32_t* pos_out, uint64_t* mer_out,
                        uint64_t* shard_len_out);
int mcu_sml_build_sharded(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* sml_len_out, float* ms_out);

/* device times of the last mcu_sml_build call (6 floats): [0] pack ms, [1] seed generation ms, [2] radix sort ms (CUDA events
 * on the launching stream), [3] radix passes, [4] key bytes (4 or 8), [5] list length */
void mcu_sml_last_stats(float* out6);

/* ---- seed-match enumeration + extension: replaces MemHash::FindMatches(MatchList&)
 *      (LM/MemHash.cpp:109-127) for two genomes, i.e. MatchFinder::SearchRange
 *      (LM/MatchFinder.cpp:172-340) -> EnumerateMatches (LM/PairwiseMatchFinder.cpp:37-71 for
 *      rule 0; LM/MemHash.cpp:139-162 with tolerances 0/1 for rule 1) -> HashMatch/SetDirection
 *      (:167-203) -> AddHashEntry (:209-251) -> ExtendMatch (LM/MatchFinder.h:218-374)
 *      -> GetMatchList (LM/MemHash.h:183-203).  Rows come back in the reference's list order.
 *      stats (optional, 8 x uint64): [0] seed pairs (unique in both genomes), [1] matches,
 *      [2] collisions = [0]-[1] (MemHash::MemCollisionCount), [3] 1 if some join run exceeds
 *      MER_REPEAT_LIMIT=1000: the reference's skip-ahead branch (LM/MatchFinder.cpp:253-277) is not
 *      reproduced.  The dropped run itself never produces seed pairs, but the reference resumes the other
 *      genome's list mid-run (the `&&` at :117) and can then report a few spurious matches between repeat
 *      copies whose identity depends on std::sort's tie order; they are left out,
 *      [4] extension candidates,
 *      [5] sorted (key, position) entries, [6] hash buckets replayed in insertion order,
 *      [7] duplicate rows the replay added (the reference stores some matches twice).            */
#define MCU_RULE_PAIRWISE 0
#define MCU_RULE_MEMHASH 1
int mcu_find_mums(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule,
                  mcu_match** out, uint64_t* n_out, uint64_t* stats);

/* The same call with the rows written into CALLER memory (pinned memory from mcu_host_alloc makes the device-to-host copy a
 * plain DMA): rows_out has room for cap rows; *n_out = rows found; MCU_ESMALL (nothing copied) when they do not fit.
 * Inputs of 8 MB and more are uploaded in pieces on a copy stream and the 2-bit pack + first partition pass of a piece run
 * under the copies of the next ones (mcu_find_mums does the same).                                                        */
int mcu_find_mums_into(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule,
                       mcu_match* rows_out, uint64_t cap, uint64_t* n_out, uint64_t* stats);

/* ---- batched gap search: replaces the per-gap calls of recursive anchoring, pairwiseAnchorSearch
 *      (LM/ProgressiveAligner.cpp:590-679, called for every gap by recurseOnPairs :681-924) and the pairwise part of
 *      SearchLCBGaps (LM/Aligner.cpp:784-930): per gap two DNAMemorySML::Create + MemHash::FindMatches with the MUM
 *      settings.  Pair i = seq0[off0[i], off0[i+1]) vs seq1[off1[i], off1[i+1]) searched with pattern seeds[i]
 *      (the caller picks it as the reference does: getSeed(getDefaultSeedWeight(average gap length), 0), :617-626;
 *      seeds[i] == 0 skips the pair, as the reference does for weights < 5, :627).  Pairs sharing a pattern are
 *      processed in one launch sequence (segment-tagged keys, one radix sort).  Output: *out = all rows, pair by
 *      pair, each pair's rows in the reference's list order with coordinates local to the pair's sequences;
 *      out_off (n_pairs + 1, caller-allocated): rows of pair i are [out_off[i], out_off[i+1]).  stats (optional,
 *      4 x uint64): [0] unique seed pairs, [1] matches, [2] pairs redone one by one because a hash bucket was
 *      order dependent (csrc/replay.cu), [3] distinct patterns (= launch sequences).                              */
int mcu_find_mums_batch(uint64_t n_pairs, const char* seq0, const uint64_t* off0, const char* seq1, const uint64_t* off1,
                        const uint64_t* seeds, int rule, mcu_match** out, uint64_t* out_off, uint64_t* stats);

/* ---- device-resident session (measurement + multi-GPU sharding) -------------------------
 * Same computation as mcu_find_mums, split so the timed region can start with both genomes
 * already in HBM.  shard_index/shard_count partition the seeds among ranks by a
 * hash of forward ^ reverse-complement mer, identical for a mer and its reverse complement: every
 * rank keeps the seeds it owns; the per-rank match lists are merged on rank 0 (mcu_session_merge
 * after the two-phase run below, or mcu_merge_matches for independently finished shards).   */
typedef struct mcu_session mcu_session;
int mcu_session_create(mcu_session** out);
void mcu_session_destroy(mcu_session* s);
/* H2D of both genomes (async on the session stream, then synchronised). */
int mcu_session_upload(mcu_session* s, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1);
/* The same copy in `chunks` pieces per genome on a copy stream, returning at once: the next mcu_session_run packs and
 * partitions every piece as it lands.  The host buffers must stay valid until that run has returned.                 */
int mcu_session_upload_begin(mcu_session* s, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, int chunks);
/* pack + enumerate + extend + order on the session stream.  stage_ms (optional, 16 floats, CUDA-event
 * times in ms): [0] pack, [1] seed generation (+ level-1 scatter), [2] sort (or level-2 scatter),
 * [3] join (or in-bucket grouping), [4] candidates + extend, [5] order + replay, [6] total,
 * [7] radix passes run (-1: bucketed enumeration, csrc/bucket.cu), per-kernel times: [8] bk_hist1,
 * [9] bk_scatter1, [10] bk_hist2, [11] bk_scatter2, [12] bk_group, [13] candidate, [14] extend, [15] records spilled to the sort path.
 * Leaves the ordered match list on the device.                                               */
int mcu_session_run(mcu_session* s, uint64_t seed, int shard_index, int shard_count,
                    float* stage_ms, uint64_t* stats);
/* The same run in two phases, for multi-GPU use.  mcu_session_enumerate: pack + enumeration of the unique seed pairs
 * whose (mixed) mer falls in this rank's slice.  Extension needs to know about unique seeds of ALL slices (a match is
 * emitted by its leftmost unique seed, wherever that seed's mer hashes), so between the phases the ranks combine their
 * unique-seed bitmaps: mcu_session_uniq_bitmap exposes the device buffer (n_words x uint32); the slices' bits are
 * disjoint, so a SUM all-reduce of the words is their OR.  mcu_session_finish(uniq_is_global != 0) then runs candidates +
 * extension + ordering; every match is found by exactly one rank.  mcu_session_run == enumerate + finish(0).            */
int mcu_session_enumerate(mcu_session* s, uint64_t seed, int shard_index, int shard_count);
int mcu_session_uniq_bitmap(mcu_session* s, void** device_words_out, uint64_t* n_words_out);
int mcu_session_finish(mcu_session* s, int uniq_is_global, float* stage_ms, uint64_t* stats);
/* Rank-0 merge for runs finished with a global bitmap: `rows` = concatenation of the ranks' lists (device or host
 * pointer).  Orders them into the reference list order and replays order-dependent hash buckets exactly (csrc/replay.cu),
 * using this session's genomes and bitmap.  The result replaces the session's match list (mcu_session_match_count /
 * mcu_session_download).  stats2 (optional, 2 x uint64): [0] buckets replayed, [1] duplicate rows added.          */
int mcu_session_merge(mcu_session* s, const mcu_match* rows, uint64_t n, int in_device, uint64_t* stats2);
/* number of matches produced by the last run */
uint64_t mcu_session_match_count(const mcu_session* s);
/* copy of the match list into caller memory, host or device (n = mcu_session_match_count rows). */
int mcu_session_download(mcu_session* s, mcu_match* out);
/* device pointer to the match rows (3 x int64 each) for NCCL gathers by the host language */
const void* mcu_session_matches_device(const mcu_session* s);
/* kernels launched by this session since creation (bench.py's gpu_launches claim) */
uint64_t mcu_session_launch_count(const mcu_session* s);
/* Rank-0 merge of per-shard lists gathered into one device or host array: sorts into the
 * reference list order and drops the duplicates that shards produce for one maximal match
 * (the distributed form of AddHashEntry's containment check).  in_device != 0 means `rows`
 * is a device pointer.  Result in library-owned host memory (*out, *n_out).
 * *unclean_buckets_out (optional) = number of hash buckets whose content depends on the reference's
 * insertion order (see csrc/replay.cu); when it is non-zero the merged list may differ from the
 * reference's (duplicate rows / order inside those buckets) and the caller re-runs the pair unsharded
 * (mcu_session_run with shard_count 1), which replays such buckets exactly.                          */
int mcu_merge_matches(const mcu_match* rows, uint64_t n, int in_device, mcu_match** out, uint64_t* n_out,
                      uint64_t* unclean_buckets_out);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink inside this library.  The reference has no distributed
 *      match finder; its closest relative is ParallelMemHash (LM/ParallelMemHash.cpp:63-101: OpenMP threads over mer ranges of
 *      the sorted lists, one MemHash per thread, merged afterwards).  Here every seed is owned by the rank its mer hashes to.
 *      Start-up: rank 0 calls mcu_comm_unique_id and hands the 128 bytes to the other ranks by any means (file, socket, MPI,
 *      torch.distributed store); every rank then calls mcu_init(local device) and mcu_comm_init(rank, world, id).
 *      All entry points below are COLLECTIVE: every rank calls them with the same arguments, in the same order.            */
#define MCU_COMM_ID_BYTES 128
int mcu_comm_unique_id(void* id_out);
int mcu_comm_init(int rank, int world, const void* id);
void mcu_comm_destroy(void);
int mcu_comm_rank(void);
int mcu_comm_world(void);
/* device synchronise + a one-word all-reduce: every rank has finished what it enqueued before any rank returns */
int mcu_comm_barrier(void);
/* in-place all-reduce of n host doubles; op: 0 sum, 1 max, 2 min (timing: max over ranks) */
int mcu_comm_allreduce_f64(double* v, int n, int op);
/* variable-length gather of host bytes to rank 0 over NCCL send/recv (DP paths, HMM predictions of a sharded batch).
 * Rank 0: *out = concatenation in rank order (library-owned, mcu_free), counts_out[world] = bytes per rank; others: *out = NULL */
int mcu_comm_gather_bytes(const void* send, uint64_t n, void** out, uint64_t* counts_out);
int mcu_device_synchronize(void);
/* H2D of the slice of both genomes this rank packs (1 / world of each, asynchronous on the session stream); with world == 1
 * the same as mcu_session_upload.  A session filled by mcu_session_upload (whole genomes on every rank) works as well.    */
int mcu_session_upload_sharded(mcu_session* s, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1);
/* One seed + match + extend pass sharded over the ranks, all on the session's stream: pack 1 / world + ncclAllGather of the
 * packed genomes, enumeration of this rank's seeds, ncclAllReduce(SUM == OR) of the unique-seed bitmaps, extension,
 * ncclAllGather of the row counts, ncclSend / ncclRecv of the rows to rank 0, merge there (reference list order + exact
 * bucket replay).  Afterwards rank 0's session holds the complete list (mcu_session_match_count / mcu_session_download);
 * stats as in mcu_find_mums, summed over the ranks; stage_ms[6] = the whole step on this rank.                             */
int mcu_session_run_sharded(mcu_session* s, uint64_t seed, float* stage_ms, uint64_t* stats);
/* mcu_find_mums_into as a collective: host buffers in, rows in rank 0's rows_out (other ranks: *n_out only). */
int mcu_find_mums_sharded(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule,
                          mcu_match* rows_out, uint64_t cap, uint64_t* n_out, uint64_t* stats);

/* ---- gapped DP: replaces muscle::GlobalAlign (MU/glbalign.cpp:69-81 -> NWSmall
 *      MU/nwsmall.cpp:500-670 + BitTraceBack MU/bittraceback.cpp:138-) for batches of
 *      two single-sequence ACGT profiles (the 2-genome case).
 *      a/b: concatenated sequences; a_off/b_off: n+1 offsets.  path_off (n+1, input): where each
 *      problem's path goes in path_out (capacity la+lb per problem).  Outputs: path_out edge
 *      types 'M','D','I' first-to-last edge (PWPath order), path_len[n], score[n] (max of
 *      MAB/DAB/IAB -- NWSmall itself returns 0).  gcups_ms (optional): device time in ms.     */
int mcu_nw_batch(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off,
                 const uint64_t* path_off, char* path_out, uint32_t* path_len, int64_t* score, float* device_ms);
/* counters of the last mcu_nw_batch call (5 x uint64): [0] DP cells (sum la*lb), [1] forward kernel
 * launches, [2] traceback-side launches, [3] sub-batches, [4] traceback bytes of the largest one */
void mcu_nw_last_stats(uint64_t* out5);

/* The same call for regions whose sequences contain DNA wildcards (N X M R W S Y K V H D B, either case): there
 * NWSmall's scores are not integers, and the path is reproduced by executing the reference's float operations in the reference's
 * order (MSA::GetFractionalWeightedCounts MU/msa2.cpp:20-90 incl. its treatment of 'X', ProfileFromMSA MU/profilefrommsa.cpp:246-322,
 * ScoreProfPos2SPN MU/scorepp.cpp:80-92, NWSmall, BitTraceBack) as a float32 wavefront: a warp per region, the cell update of
 * NWSmall in round-to-nearest operations without contraction, 4 traceback bits per cell.  Meant for the ranges mcu_nw_batch refuses
 * with MCU_EALPHA; pure ACGT regions give the same paths here, only slower.  score: max(MAB, DAB, IAB) as the reference's float.
 * Sequences of up to 2^26 - 1 letters each (MCU_EINVAL beyond; MCU_ENOMEM when one region's traceback does not fit the device);
 * a byte that is no DNA letter or wildcard gives MCU_EALPHA.   */
int mcu_nw_batch_wild(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off,
                      const uint64_t* path_off, char* path_out, uint32_t* path_len, float* score, float* device_ms);

/* ---- HomologyHMM: replaces run() (LM/HomologyHMM/homologymain.cc:24-62 = Forward
 *      homology.cc:307-394 + Backward :400-547 + posterior threshold 0.9) for a batch of
 *      symbol strings over '1'..'8' (encoder: LM/Islands.h:90-155).
 *      params: 21 doubles in struct Params order (homology.h:169-177): iStartHomologous,
 *      iGoHomologous, iGoUnrelated, iGoStopFromUnrelated, iGoStopFromHomologous,
 *      aEmitHomologous[8], aEmitUnrelated[8].  sym/off: concatenated strings + n+1 offsets.
 *      pred_out: 'H'/'N' per column; post_out (optional): posterior of "homologous".        */
int mcu_hmm_params(double gc_content, double go_homologous, double go_unrelated, double pct_identity, double* params_out);
int mcu_hmm_batch(uint64_t n, const char* sym, const uint64_t* off, const double* params,
                  char* pred_out, double* post_out, float* device_ms);

/* ---- seed occurrence list: replaces mems::SeedOccurrenceList::construct
 *      (LM/SeedOccurrenceList.h:22-78, smoothFrequencies :103-119), called once per genome on the match finder's sorted mer
 *      list (LM/ProgressiveAligner.cpp:3908-3912).  freq_out[n]: what getFrequency(position) returns (:81-84): the
 *      multiplicity of the seed starting at each position (1 for the last L-1 positions), averaged over the L seeds that
 *      start at position-L+1 .. position; the last position keeps its raw count.  Bit-identical floats.
 *      n < L (no seed at all): the reference reads an uninitialised count; all ones here.                          */
int mcu_sol_build(const char* seq, uint64_t n, uint64_t seed, float* freq_out);

/* ---- anchor scores: replaces mems::GetPairwiseAnchorScore (LM/GreedyBreakpointElimination.h:403-476, penalize_gaps = false:
 *      the form LM/ProgressiveAligner.cpp:1825 and :3422 call) for the LCBs of one genome pair made of ungapped matches.
 *      rows[n_rows]: matches (1-based starts, negative = reverse strand); LCB l owns rows [lcb_off[l], lcb_off[l+1]).
 *      freq0/freq1: the genomes' seed occurrence lists from mcu_sol_build (n0 / n1 floats); NULL = built here from `seed`.
 *      matrix: 16 x int32 substitution scores [A,C,G,T][A,C,G,T] (NULL = hoxd_matrix, LM/SubstitutionMatrix.h:23-33).
 *      penalize_repeats: the reference's global of that name (LM/GreedyBreakpointElimination.cpp:37, default false).
 *      Outputs: lcb_score_out[n_lcb] (the function's return value per LCB), match_score_out[n_rows] (optional: m_score of
 *      every match, an integer).  Sequence bytes outside the IUPAC DNA alphabet in reverse-strand rows are not supported
 *      (gnFilter::ReverseFilter deletes them, which shifts the reference's columns).                                  */
int mcu_anchor_scores(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, const float* freq0, const float* freq1,
                      const mcu_match* rows, uint64_t n_rows, const uint64_t* lcb_off, uint64_t n_lcb, const int32_t* matrix,
                      int penalize_repeats, double* lcb_score_out, int64_t* match_score_out);

/* ---- the step after the match list, two genomes.  Rows are (len > 0, start0 > 0, start1 != 0 signed).
 *      mcu_eliminate_overlaps replaces mems::EliminateOverlaps_v2(ml, eliminate_both) (LM/ProgressiveAligner.h:300-406) followed, when
 *      min_length > 0, by ml.LengthFilter(min_length) (LM/MatchList.h:680-692): the sequence pairwiseAnchorSearch runs after every gap
 *      search (LM/ProgressiveAligner.cpp:656-660) and, with eliminate_both, the pairwise LCB set-up after the initial anchoring
 *      (:3408-3410).  rows_out (n rows of room): what the reference's list holds afterwards, in its order.
 *      mcu_lcbs replaces mems::IdentifyBreakpoints + ComputeLCBs_v2 (LM/GreedyBreakpointElimination.h:161-250): sorted_out = the list
 *      ordered on genome 0, breakpoints_out (n entries of room) = index of the last match of every LCB, ascending: LCB l is
 *      sorted_out[breakpoints[l-1] + 1 .. breakpoints[l]].
 *      The reference orders its lists with std::sort on one start coordinate; where starts tie, the outcome depends on where
 *      libstdc++'s introsort leaves the tied rows, and that algorithm is what runs here then (csrc/lcb.cu).  ties_out (optional): how
 *      many adjacent ties the orderings met (0: any correct sort gives this result). */
int mcu_eliminate_overlaps(const mcu_match* rows, uint64_t n, int eliminate_both, uint64_t min_length, mcu_match* rows_out, uint64_t* n_out,
                           uint64_t* ties_out);
int mcu_lcbs(const mcu_match* rows, uint64_t n, mcu_match* sorted_out, uint64_t* breakpoints_out, uint64_t* n_breakpoints_out, uint64_t* ties_out);

/* ---- anchor columns of alignment windows: replaces muscle::FindAnchorColsPP (MU/anchoredpp.cpp:354-409), the step
 *      of AnchoredProfileProfile (:443-552) that decides which ranges of a window the gapped DP re-aligns:
 *      LetterObjScoreXP (:256-329: ScoreSeqPairLetters :19-93 and the per-site ScoreSeqPairGaps :96-250 for every pair of rows, summed
 *      with the rows' weights), WindowSmooth (MU/anchors.cpp:9-47), FindBestColsComboPP (:335-351), MergeBestCols (MU/anchors.cpp:137-186).
 *      All scores are the reference's floats, every sum in the reference's order: columns, scores and smoothed scores are identical.
 *      The settings are MUSCLE's globals at the time of the call (mcu_anchor_default_params: what MuscleInterface::ProfileAlignFast
 *      leaves in force for DNA, LM/MuscleInterface.cpp:1086-1106).                                                              */
#define MCU_AC_GAP 255u
typedef struct {
    float subst[4][4];        /* (*g_ptrScoreMatrix)[a][b] on the four residues (g_AlphaSize = 4)               */
    float gap_open;           /* g_scoreGapOpen                                                                 */
    float gap_extend;         /* g_scoreGapExtend                                                               */
    float term_gap;           /* TermGapScore(true), MU/objscore2.cpp:22-40                                     */
    float smooth_ceil;        /* g_dSmoothScoreCeil                                                             */
    float min_best_col;       /* g_dMinBestColScore                                                             */
    float min_smooth;         /* g_dMinSmoothScore                                                              */
    uint32_t smooth_window;   /* g_uSmoothWindowLength as FindAnchorColsPP sets it: 21 (odd, MCU_EINVAL if not) */
    uint32_t anchor_spacing;  /* g_uAnchorSpacing as FindAnchorColsPP sets it: 96                               */
    uint8_t letter_of_char[256]; /* CharToLetterEx per character: 0..3 residues, MCU_AC_GAP where IsGapChar,
                                    any other value = a letter outside the alphabet (wildcards: scored 0)      */
} mcu_anchor_params;
void mcu_anchor_default_params(mcu_anchor_params* p);
/* n windows.  Window i is two alignments of ncol[i] columns with n1[i] and n2[i] rows: its (n1 + n2) rows of ncol characters lie
 * one after the other from rows + row_off[i], the first alignment's rows first (characters as the MSA holds them after FixAlpha).
 * weights: MSA::GetSeqWeight of every row, window after window in row order (NULL: all 1, the two-genome case).
 * params NULL = mcu_anchor_default_params.  col_off (n + 1 offsets, col_off[i+1] - col_off[i] >= ncol[i]): where window i's outputs
 * go in cols_out (its anchor columns, ascending; n_cols_out[i] of them), score_out and smooth_out (optional: MatchScore[] and
 * SmoothScore[] of FindAnchorColsPP, ncol[i] floats).  device_ms (optional): the kernel's time.
 * A window whose alignments differ in length has no anchor columns in the reference (:358-362): the caller answers that itself. */
int mcu_anchor_cols_batch(uint64_t n, const char* rows, const uint64_t* row_off, const uint32_t* ncol, const uint32_t* n1, const uint32_t* n2,
                          const float* weights, const mcu_anchor_params* params, const uint64_t* col_off, uint32_t* cols_out,
                          uint32_t* n_cols_out, float* score_out, float* smooth_out, float* device_ms);

/* ---- test hooks (exercise single kernels through the ABI) -------------------------------- */
/* stable LSD radix sort of (key,val) pairs on the low `bits` bits; key_bytes is 4 or 8 */
int mcu_test_sort_pairs(void* keys, uint32_t* vals, uint64_t n, int key_bytes, int bits);
/* INT32 issue-rate microbenchmark (the measured denominator of the gapped-DP roofline): 8 independent add/max chains
 * per thread, register resident, one full-device launch timed with CUDA events.  *gops_out = thread-level integer instructions / s / 1e9
 * (ptxas fuses every add+max pair of the loop into one VIADDMNMX: one instruction, two operations). */
int mcu_test_int32_peak(double* gops_out, float* ms_out);
/* HomologyHMM, few long strings (one warp per chain, csrc/hmm.cu): of the last mcu_hmm_batch call, out3[0] = columns the chains
 * stepped through (both directions), out3[1] = chain rounds (a round ends at a column whose exponents move or whose FP32 products are
 * hazardous), out3[2] = columns that fell back from the FP32 form of the recurrence to the operation-by-operation FP64 form.
 * MAUVE_CUDA_HMM_FP64=1 in the environment sends every column down the FP64 form (A/B in the tests); MAUVE_CUDA_HMM_TEST_FAULT=N flips
 * the last mantissa bit of the chain's state at the end of every N-th group of eight columns, so that the tests can watch the kernel
 * repair its chain (the result stays bit-identical). */
int mcu_test_hmm_counters(uint64_t* out3);
/* counters of the last mcu_anchor_cols_batch call (8 x uint64): [0] 32-column segments of the smoothing chain that were finished in
 * exact arithmetic by 32 lanes at once, [1] segments that ran as the serial float chain (csrc/anchorcols.cu), [2..7] SM cycles the
 * first window spent scoring the columns, smoothing, selecting the best columns, searching the groups' ends, walking the groups,
 * picking the anchor of every group */
int mcu_test_anchor_counters(uint64_t* out8);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MAUVE_CUDA_H_ */
