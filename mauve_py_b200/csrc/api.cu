// C ABI of libmauve_cuda.so (include/mauve_cuda.h): argument checking, device selection,
// the process-wide default session, seed-pattern tables.  No compute happens on the host.
#include <stdarg.h>
#include <math.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <new>

#include "anchor.cuh"
#include "dp.cuh"
#include "hmm.cuh"
#include "lcb.cuh"
#include "anchorcols.cuh"
#include "sol.cuh"

namespace mcu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static int g_device = -1;
static int g_sms = 0;
static std::mutex g_mu;
static std::recursive_mutex g_dev_mu;  // device selection (mcu_init / ensure_device); g_mu serialises the default session

int ensure_device()
{
    std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
    if (g_device >= 0) {
        cudaError_t e = cudaSetDevice(g_device);
        if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", g_device, cudaGetErrorString(e)); return MCU_ENODEV; }
        return MCU_OK;
    }
    return mcu_init(0);
}
int sm_count() { return g_sms > 0 ? g_sms : 148; }

// ---- seed patterns (LM/SeedMasks.h:44-260; low 32 bits, the high words are all zero) ----------
// Quirks kept: weight-11 rank-0 entry equals the weight-12 one; weight-21 rank-1 is 0xaeb3f.
static const u32 SEED_TABLE[32][6] = {
    {0}, {0}, {0},
    {0xb}, {0x3b},
    {0x6b, 0x139, 0x193, 0x6b},
    {0x58D, 0x653, 0x1AB, 0xdb},
    {0x1953, 0x588d, 0x688b, 0x17d, 0x164d},
    {0x3927, 0x1CA7, 0x6553, 0xb6d},
    {0x7497, 0x1c927, 0x72a7, 0x6fb, 0x16ed},
    {0x1d297, 0x3A497, 0xE997, 0x6D5B},
    {0x7954f, 0x75257, 0x1c9527, 0x5bed, 0x5b26d},
    {0x7954f, 0x3D32F, 0x768B7, 0x5B56D},
    {0x792a4f, 0x1d64d7, 0x1d3597, 0x1b7db, 0x75ad7},
    {0x1e6acf, 0xF59AF, 0x3D4CAF, 0x35AD6B},
    {0x7ac9af, 0x7b2a6f, 0x79aacf, 0x16df6d, 0x6b5d6b},
    {0xf599af, 0xEE5A77, 0x7CD59F, 0xEB5AD7},
    {0x6dbedb},
    {0x3E6B59F, 0x3EB335F, 0x7B3566F},
    {0x7b974ef, 0x7d6735f, 0x1edd74f},
    {0x1F59B35F, 0x3EDCEDF, 0xFAE675F},
    {0x7ddaddf, 0xaeb3f, 0x7eb76bf},
    {0x003fffff}, {0x007fffff}, {0x00ffffff}, {0x01ffffff}, {0x03ffffff},
    {0x07ffffff}, {0x0fffffff}, {0x1fffffff}, {0x3fffffff}, {0x7fffffff}};

int make_seed_params(u64 seed, SeedParams* out)
{
    SeedParams sp;
    memset(&sp, 0, sizeof sp);
    sp.seed = seed;
    sp.L = mcu_seed_length(seed);
    sp.w = mcu_seed_weight(seed);
    // GetSeedMer scans the pattern from bit L-1 down to bit 0 (LM/SortedMerList.cpp:737-757) and
    // GetMer holds at most 32 bases; a pattern with trailing zeros or L > 31 is not a valid DNA seed.
    if (seed == 0 || !(seed & 1) || sp.L > 31 || sp.w < 1) {
        set_error("unusable seed pattern 0x%llx (length %d, weight %d)", (unsigned long long)seed, sp.L, sp.w);
        return MCU_EINVAL;
    }
    int j = 0, cum = 0;
    while (j < sp.L) {
        if (!((seed >> (sp.L - 1 - j)) & 1)) { ++j; continue; }
        int len = 0;
        while (j + len < sp.L && ((seed >> (sp.L - 1 - (j + len))) & 1)) ++len;
        if (sp.nruns >= MCU_MAX_RUNS) { set_error("seed pattern has too many runs"); return MCU_EINVAL; }
        sp.shift_in[sp.nruns] = (u8)(2 * (sp.L - j - len));
        sp.shift_out[sp.nruns] = (u8)(2 * (sp.w - cum - len));
        sp.mask[sp.nruns] = len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1);
        ++sp.nruns;
        cum += len;
        j += len;
    }
    u64 rev = 0;
    for (int b = 0; b < sp.L; ++b) rev |= ((seed >> b) & 1) << (sp.L - 1 - b);
    sp.palindromic = rev == seed;
    *out = sp;
    return MCU_OK;
}

static Session* g_default_session = nullptr;

std::mutex& api_mutex() { return g_mu; }

int default_session(Session** out)
{
    if (!g_default_session) {
        Session* s = new (std::nothrow) Session();
        if (!s) return MCU_ENOMEM;
        int r = session_init(*s);
        if (r != MCU_OK) { delete s; return r; }
        g_default_session = s;
    }
    *out = g_default_session;
    return MCU_OK;
}

}  // namespace mcu

using namespace mcu;

extern "C" {

int mcu_init(int device)
{
    std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no usable CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return MCU_ENODEV;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (0..%d)", device, count - 1); return MCU_EINVAL; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); return MCU_ENODEV; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return MCU_ENODEV; }
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return MCU_ENODEV;
    }
    g_device = device;
    g_sms = prop.multiProcessorCount;
    return MCU_OK;
}

void mcu_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_default_session) {
        session_destroy(*g_default_session);
        delete g_default_session;
        g_default_session = nullptr;
    }
    // the gapped-DP, HMM and LCB entries keep their device buffers between calls: give them back too
    nw_release();
    nwf_release();
    hmm_release();
    lcb_release();
    ac_release();
}

const char* mcu_last_error(void) { return get_error(); }
void mcu_free(void* p) { free(p); }

int mcu_host_alloc(void** out, uint64_t bytes)
{
    MCU_TRY(ensure_device());
    MCU_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return MCU_OK;
}
void mcu_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- seeds ------------------------------------------------------------------------------
uint64_t mcu_get_seed(int weight, int rank)
{
    auto solid = [](int w) -> u64 { return w >= 64 ? ~0ull : ((1ull << w) - 1); };
    if (weight < 0 || rank < 0) return 0;
    if (rank == MCU_SOLID_SEED) return solid(weight);
    if (weight > 31) return solid(32);
    if (rank > 5) return solid(weight);
    if (SEED_TABLE[weight][rank] == 0) return solid(weight);
    return SEED_TABLE[weight][rank];
}

unsigned mcu_default_seed_weight(uint64_t avg_len)
{
    unsigned w = (unsigned)ceil((log((double)avg_len) / log(2.0)) / 1.5);
    if (!(w & 1)) ++w;
    if (w < 5) w = 0;
    if (avg_len == 0) w = 0;
    if (w > 31) w = 31;
    return w;
}

int mcu_seed_length(uint64_t seed)
{
    if (!seed) return 0;
    int hi = 63 - __builtin_clzll(seed), lo = __builtin_ctzll(seed);
    return hi - lo + 1;
}

int mcu_seed_weight(uint64_t seed) { return __builtin_popcountll(seed); }

// ---- SML --------------------------------------------------------------------------------
int mcu_sml_build(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* mer_out, uint32_t* packed_out,
                  uint64_t* sml_len_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (n && !seq) { set_error("mcu_sml_build: NULL sequence"); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    return sml_build_device(*s, seq, n, seed, pos_out, mer_out, packed_out, sml_len_out);
}

int mcu_sml_build_shard(const char* seq, uint64_t n, uint64_t seed, int shard, int n_shards, uint32_t* pos_out, uint64_t* mer_out, uint64_t* shard_len_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (n && !seq) { set_error("mcu_sml_build_shard: NULL sequence"); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    return sml_build_device(*s, seq, n, seed, pos_out, mer_out, nullptr, shard_len_out, shard, n_shards);
}

void mcu_sml_last_stats(float* out6)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!out6 || !g_default_session) return;
    const Session& s = *g_default_session;
    out6[0] = s.sml_ms[0]; out6[1] = s.sml_ms[1]; out6[2] = s.sml_ms[2];
    out6[3] = (float)s.sml_passes; out6[4] = (float)s.sml_key_bytes_last; out6[5] = (float)s.sml_npos;
}

// ---- MUMs -------------------------------------------------------------------------------
int mcu_find_mums(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match** out,
                  uint64_t* n_out, uint64_t* stats)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (!out || !n_out) { set_error("mcu_find_mums: NULL output pointer"); return MCU_EINVAL; }
    if (rule != MCU_RULE_PAIRWISE && rule != MCU_RULE_MEMHASH) { set_error("mcu_find_mums: unknown rule %d", rule); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    if (n0 + n1 >= (8ull << 20) && getenv("MAUVE_CUDA_NO_OVERLAP") == nullptr) MCU_TRY(session_upload_begin(*s, seq0, n0, seq1, n1, 8));
    else MCU_TRY(session_upload(*s, seq0, n0, seq1, n1));
    MCU_TRY(session_run(*s, seed, 0, 1, nullptr, stats));
    u64 m = s->match_count;
    mcu_match* r = (mcu_match*)malloc((m ? m : 1) * sizeof(mcu_match));
    if (!r) { set_error("out of host memory"); return MCU_ENOMEM; }
    if (m) {
        cudaError_t e = cudaMemcpyAsync(r, s->matches.p, m * sizeof(mcu_match), cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) { free(r); set_error("mcu_find_mums: copy of the rows failed: %s", cudaGetErrorString(e)); return MCU_ECUDA; }
    }
    *out = r;
    *n_out = m;
    return MCU_OK;
}

int mcu_find_mums_into(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match* rows_out,
                       uint64_t cap, uint64_t* n_out, uint64_t* stats)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (!n_out || (cap && !rows_out)) { set_error("mcu_find_mums_into: NULL output pointer"); return MCU_EINVAL; }
    if (rule != MCU_RULE_PAIRWISE && rule != MCU_RULE_MEMHASH) { set_error("mcu_find_mums_into: unknown rule %d", rule); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    // large inputs arrive in pieces on a copy stream; pack and the level-1 partition of a piece run under the next copies
    if (n0 + n1 >= (8ull << 20) && getenv("MAUVE_CUDA_NO_OVERLAP") == nullptr) MCU_TRY(session_upload_begin(*s, seq0, n0, seq1, n1, 8));
    else MCU_TRY(session_upload(*s, seq0, n0, seq1, n1));
    MCU_TRY(session_run(*s, seed, 0, 1, nullptr, stats));
    const u64 m = s->match_count;
    *n_out = m;
    if (m > cap) { set_error("mcu_find_mums_into: %llu rows do not fit the caller's %llu", (unsigned long long)m, (unsigned long long)cap); return MCU_ESMALL; }
    if (m) {
        MCU_CUDA(cudaMemcpyAsync(rows_out, s->matches.p, m * sizeof(mcu_match), cudaMemcpyDeviceToHost, s->stream));
        MCU_CUDA(cudaStreamSynchronize(s->stream));
    }
    return MCU_OK;
}

// ---- batched gap search -------------------------------------------------------------------
int mcu_find_mums_batch(uint64_t n_pairs, const char* seq0, const uint64_t* off0, const char* seq1, const uint64_t* off1, const uint64_t* seeds,
                        int rule, mcu_match** out, uint64_t* out_off, uint64_t* stats)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (!out || !out_off) { set_error("mcu_find_mums_batch: NULL output pointer"); return MCU_EINVAL; }
    if (n_pairs && (!seq0 || !seq1 || !off0 || !off1 || !seeds)) { set_error("mcu_find_mums_batch: NULL input pointer"); return MCU_EINVAL; }
    if (rule != MCU_RULE_PAIRWISE && rule != MCU_RULE_MEMHASH) { set_error("mcu_find_mums_batch: unknown rule %d", rule); return MCU_EINVAL; }
    if (n_pairs >= 0xFFFFFFFFull) { set_error("mcu_find_mums_batch: too many pairs"); return MCU_EINVAL; }
    for (uint64_t i = 0; i < n_pairs; ++i)
        if (off0[i + 1] < off0[i] || off1[i + 1] < off1[i]) { set_error("mcu_find_mums_batch: offsets not monotone at pair %llu", (unsigned long long)i); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    // pairs that share a seed pattern go through the device together (LM/ProgressiveAligner.cpp:617-626: the weight depends on
    // the gap length only, so a round of recursive anchoring uses a handful of patterns); seed 0 = "no search" (:627)
    std::map<uint64_t, std::vector<uint32_t> > groups;
    for (uint64_t i = 0; i < n_pairs; ++i)
        if (seeds[i]) groups[seeds[i]].push_back((uint32_t)i);
    std::vector<mcu_match> all_rows;
    std::vector<uint32_t> all_pair;
    std::vector<uint32_t> redo;
    uint64_t seed_pairs_total = 0;
    for (auto& g : groups) {
        const std::vector<uint32_t>& idx = g.second;
        // chunks keep the concatenations below the 32-bit position limit of the reference's match coordinates
        size_t first = 0;
        while (first < idx.size()) {
            std::string c0, c1;
            std::vector<u64> o0(1, 0), o1(1, 0);
            size_t last = first;
            while (last < idx.size()) {
                const uint64_t i = idx[last];
                const u64 l0 = off0[i + 1] - off0[i], l1 = off1[i + 1] - off1[i];
                if (last > first && (c0.size() + l0 >= 0xF0000000ull || c1.size() + l1 >= 0xF0000000ull)) break;
                c0.append(seq0 + off0[i], l0);
                c1.append(seq1 + off1[i], l1);
                o0.push_back(c0.size());
                o1.push_back(c1.size());
                ++last;
            }
            std::vector<mcu_match> rows;
            std::vector<u32> seg, unclean;
            u64 sp = 0;
            MCU_TRY(batch_find_mums(*s, c0.data(), o0.data(), c1.data(), o1.data(), (u32)(last - first), g.first, &rows, &seg, &unclean, &sp));
            seed_pairs_total += sp;
            std::vector<char> is_unclean(last - first, 0);
            for (u32 u : unclean) { is_unclean[u] = 1; redo.push_back(idx[first + u]); }
            for (size_t r = 0; r < rows.size(); ++r) {
                if (is_unclean[seg[r]]) continue;
                all_rows.push_back(rows[r]);
                all_pair.push_back(idx[first + seg[r]]);
            }
            first = last;
        }
    }
    // order-dependent hash buckets (csrc/replay.cu): those pairs take the single-pair path, which replays them exactly
    for (uint32_t i : redo) {
        MCU_TRY(session_upload(*s, seq0 + off0[i], off0[i + 1] - off0[i], seq1 + off1[i], off1[i + 1] - off1[i]));
        MCU_TRY(session_run(*s, seeds[i], 0, 1, nullptr, nullptr));
        const u64 m = s->match_count;
        std::vector<mcu_match> rows(m);
        if (m) {
            MCU_CUDA(cudaMemcpyAsync(rows.data(), s->matches.p, m * sizeof(mcu_match), cudaMemcpyDeviceToHost, s->stream));
            MCU_CUDA(cudaStreamSynchronize(s->stream));
        }
        for (u64 r = 0; r < m; ++r) { all_rows.push_back(rows[r]); all_pair.push_back(i); }
    }
    // rows of one pair stay in the order they were produced; pairs ascending
    for (uint64_t i = 0; i <= n_pairs; ++i) out_off[i] = 0;
    for (uint32_t p : all_pair) out_off[p + 1]++;
    for (uint64_t i = 0; i < n_pairs; ++i) out_off[i + 1] += out_off[i];
    const u64 total = all_rows.size();
    mcu_match* r = (mcu_match*)malloc((total ? total : 1) * sizeof(mcu_match));
    if (!r) { set_error("out of host memory"); return MCU_ENOMEM; }
    std::vector<u64> cursor(out_off, out_off + n_pairs);
    for (u64 k = 0; k < total; ++k) r[cursor[all_pair[k]]++] = all_rows[k];
    *out = r;
    if (stats) { stats[0] = seed_pairs_total; stats[1] = total; stats[2] = redo.size(); stats[3] = groups.size(); }
    return MCU_OK;
}

// ---- sessions ---------------------------------------------------------------------------

int mcu_session_create(mcu_session** out)
{
    MCU_TRY(ensure_device());
    if (!out) return MCU_EINVAL;
    mcu_session* h = new (std::nothrow) mcu_session();
    if (!h) return MCU_ENOMEM;
    int r = session_init(h->s);
    if (r != MCU_OK) { delete h; return r; }
    *out = h;
    return MCU_OK;
}

void mcu_session_destroy(mcu_session* h)
{
    if (!h) return;
    session_destroy(h->s);
    delete h;
}

int mcu_session_upload(mcu_session* h, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(ensure_device());
    return session_upload(h->s, seq0, n0, seq1, n1);
}

int mcu_session_upload_begin(mcu_session* h, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, int chunks)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(ensure_device());
    return session_upload_begin(h->s, seq0, n0, seq1, n1, chunks);
}

int mcu_session_run(mcu_session* h, uint64_t seed, int shard_index, int shard_count, float* stage_ms, uint64_t* stats)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(ensure_device());
    return session_run(h->s, seed, shard_index, shard_count, stage_ms, stats);
}

int mcu_session_enumerate(mcu_session* h, uint64_t seed, int shard_index, int shard_count)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(ensure_device());
    return session_enumerate(h->s, seed, shard_index, shard_count);
}

int mcu_session_uniq_bitmap(mcu_session* h, void** device_words_out, uint64_t* n_words_out)
{
    if (!h || !device_words_out || !n_words_out) return MCU_EINVAL;
    *device_words_out = h->s.uniq.p;
    *n_words_out = h->s.run.uniq_words;
    return MCU_OK;
}

int mcu_session_finish(mcu_session* h, int uniq_is_global, float* stage_ms, uint64_t* stats)
{
    if (!h) return MCU_EINVAL;
    MCU_TRY(ensure_device());
    return session_finish(h->s, uniq_is_global != 0, stage_ms, stats);
}

int mcu_session_merge(mcu_session* h, const mcu_match* rows, uint64_t n, int in_device, uint64_t* stats2)
{
    if (!h || (n && !rows)) return MCU_EINVAL;
    MCU_TRY(ensure_device());
    Session& s = h->s;
    const mcu_match* dev_rows = rows;
    if (!in_device && n) {
        MCU_TRY(s.raw_matches.reserve(n * sizeof(mcu_match)));
        MCU_CUDA(cudaMemcpyAsync(s.raw_matches.p, rows, n * sizeof(mcu_match), cudaMemcpyHostToDevice, s.stream));
        dev_rows = s.raw_matches.as<mcu_match>();
    }
    u64 unclean = 0, dups = 0;
    MCU_TRY(session_merge(s, dev_rows, n, &unclean, &dups));
    if (stats2) { stats2[0] = unclean; stats2[1] = dups; }
    return MCU_OK;
}

uint64_t mcu_session_match_count(const mcu_session* h) { return h ? h->s.match_count : 0; }

int mcu_session_download(mcu_session* h, mcu_match* out)
{
    if (!h) return MCU_EINVAL;
    if (h->s.match_count == 0) return MCU_OK;
    if (!out) return MCU_EINVAL;
    MCU_CUDA(cudaMemcpyAsync(out, h->s.matches.p, h->s.match_count * sizeof(mcu_match), cudaMemcpyDefault, h->s.stream));
    MCU_CUDA(cudaStreamSynchronize(h->s.stream));
    return MCU_OK;
}

const void* mcu_session_matches_device(const mcu_session* h) { return h ? h->s.matches.p : nullptr; }
uint64_t mcu_session_launch_count(const mcu_session* h) { return h ? h->s.launches : 0; }

int mcu_merge_matches(const mcu_match* rows, uint64_t n, int in_device, mcu_match** out, uint64_t* n_out, uint64_t* unclean_buckets_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (!out || !n_out || (n && !rows)) { set_error("mcu_merge_matches: NULL pointer"); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    const mcu_match* dev_rows = rows;
    if (!in_device && n) {
        MCU_TRY(s->raw_matches.reserve(n * sizeof(mcu_match)));
        MCU_CUDA(cudaMemcpyAsync(s->raw_matches.p, rows, n * sizeof(mcu_match), cudaMemcpyHostToDevice, s->stream));
        dev_rows = s->raw_matches.as<mcu_match>();
    }
    MCU_TRY(order_matches(*s, dev_rows, n, 0));
    mcu_match* r = (mcu_match*)malloc((n ? n : 1) * sizeof(mcu_match));
    if (!r) { set_error("out of host memory"); return MCU_ENOMEM; }
    if (n) {
        cudaError_t e = cudaMemcpyAsync(r, s->matches.p, n * sizeof(mcu_match), cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) { free(r); set_error("mcu_merge_matches: copy of the rows failed: %s", cudaGetErrorString(e)); return MCU_ECUDA; }
    }
    // shards rediscover the same maximal match from their own seeds: equal rows are adjacent now
    u64 k = 0;
    for (u64 i = 0; i < n; ++i)
        if (k == 0 || r[i].len != r[k - 1].len || r[i].start0 != r[k - 1].start0 || r[i].start1 != r[k - 1].start1) r[k++] = r[i];
    // buckets the reference would fill order-dependently cannot be replayed from rows alone: report them
    u64 unclean = 0, dups = 0;
    if (k) {
        cudaError_t e = cudaMemcpyAsync(s->matches.p, r, k * sizeof(mcu_match), cudaMemcpyHostToDevice, s->stream);
        if (e != cudaSuccess) { free(r); set_error("mcu_merge_matches: %s", cudaGetErrorString(e)); return MCU_ECUDA; }
        s->match_count = k;
        const int rc = replay_unclean(*s, nullptr, false, &unclean, &dups);
        if (rc != MCU_OK) { free(r); return rc; }
    }
    if (unclean_buckets_out) *unclean_buckets_out = unclean;
    *out = r;
    *n_out = k;
    return MCU_OK;
}

// ---- DP / HMM ---------------------------------------------------------------------------
int mcu_nw_batch(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off, const uint64_t* path_off,
                 char* path_out, uint32_t* path_len, int64_t* score, float* device_ms)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    return nw_batch(n, a, a_off, b, b_off, path_off, path_out, path_len, score, device_ms);
}

void mcu_nw_last_stats(uint64_t* out5)
{
    if (out5) nw_last_stats(out5);
}

int mcu_hmm_params(double gc, double go_homologous, double go_unrelated, double pct_identity, double* out)
{
    if (!out) return MCU_EINVAL;
    return hmm_params(gc, go_homologous, go_unrelated, pct_identity, out);
}

int mcu_hmm_batch(uint64_t n, const char* sym, const uint64_t* off, const double* params, char* pred_out, double* post_out,
                  float* device_ms)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    return hmm_batch(n, sym, off, params, pred_out, post_out, device_ms);
}

int mcu_eliminate_overlaps(const mcu_match* rows, uint64_t n, int eliminate_both, uint64_t min_length, mcu_match* rows_out, uint64_t* n_out,
                           uint64_t* ties_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (!n_out || (n && (!rows || !rows_out))) { set_error("mcu_eliminate_overlaps: NULL pointer"); return MCU_EINVAL; }
    u64 cnt = 0, ties = 0;
    MCU_TRY(lcb_eliminate_overlaps(rows, n, eliminate_both, min_length, rows_out, &cnt, &ties));
    *n_out = cnt;
    if (ties_out) *ties_out = ties;
    return MCU_OK;
}

int mcu_lcbs(const mcu_match* rows, uint64_t n, mcu_match* sorted_out, uint64_t* breakpoints_out, uint64_t* n_breakpoints_out, uint64_t* ties_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (!n_breakpoints_out || (n && (!rows || !sorted_out || !breakpoints_out))) { set_error("mcu_lcbs: NULL pointer"); return MCU_EINVAL; }
    u64 nb = 0, ties = 0;
    MCU_TRY(lcb_breakpoints(rows, n, sorted_out, (u64*)breakpoints_out, &nb, &ties));
    *n_breakpoints_out = nb;
    if (ties_out) *ties_out = ties;
    return MCU_OK;
}

void mcu_anchor_default_params(mcu_anchor_params* p)
{
    if (p) ac_default_params(p);
}

int mcu_anchor_cols_batch(uint64_t n, const char* rows, const uint64_t* row_off, const uint32_t* ncol, const uint32_t* n1, const uint32_t* n2,
                          const float* weights, const mcu_anchor_params* params, const uint64_t* col_off, uint32_t* cols_out,
                          uint32_t* n_cols_out, float* score_out, float* smooth_out, float* device_ms)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (n && (!rows || !row_off || !ncol || !n1 || !n2 || !col_off || !cols_out || !n_cols_out)) {
        set_error("mcu_anchor_cols_batch: NULL pointer");
        return MCU_EINVAL;
    }
    return ac_batch(n, rows, (const u64*)row_off, ncol, n1, n2, weights, params, (const u64*)col_off, cols_out, n_cols_out, score_out, smooth_out, device_ms);
}

int mcu_test_anchor_counters(uint64_t* out8)
{
    if (!out8) return MCU_EINVAL;
    ac_last_counters((u64*)out8);
    return MCU_OK;
}

int mcu_test_hmm_counters(uint64_t* out3)
{
    if (!out3) return MCU_EINVAL;
    u64 c[3];
    hmm_last_counters(c);
    out3[0] = c[0];
    out3[1] = c[1];
    out3[2] = c[2];
    return MCU_OK;
}

int mcu_nw_batch_wild(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off, const uint64_t* path_off,
                      char* path_out, uint32_t* path_len, float* score, float* device_ms)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    return nw_batch_wild(n, a, a_off, b, b_off, path_off, path_out, path_len, score, device_ms);
}

// ---- seed occurrence list + anchor scores (sol.cu) ------------------------------------------------
int mcu_sol_build(const char* seq, uint64_t n, uint64_t seed, float* freq_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (n && (!seq || !freq_out)) { set_error("mcu_sol_build: NULL pointer"); return MCU_EINVAL; }
    Session* s;
    MCU_TRY(default_session(&s));
    return sol_build(*s, seq, n, seed, freq_out);
}

int mcu_anchor_scores(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, const float* freq0, const float* freq1,
                      const mcu_match* rows, uint64_t n_rows, const uint64_t* lcb_off, uint64_t n_lcb, const int32_t* matrix,
                      int penalize_repeats, double* lcb_score_out, int64_t* match_score_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if ((n0 && !seq0) || (n1 && !seq1) || (n_rows && !rows) || (n_lcb && (!lcb_off || !lcb_score_out))) {
        set_error("mcu_anchor_scores: NULL pointer");
        return MCU_EINVAL;
    }
    Session* s;
    MCU_TRY(default_session(&s));
    return anchor_scores(*s, seq0, n0, seq1, n1, seed, freq0, freq1, rows, n_rows, lcb_off, n_lcb, matrix, penalize_repeats, lcb_score_out,
                         (i64*)match_score_out);
}

// ---- test hooks -------------------------------------------------------------------------
int mcu_test_int32_peak(double* gops_out, float* ms_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    return int32_peak(gops_out, ms_out);
}

int mcu_test_sort_pairs(void* keys, uint32_t* vals, uint64_t n, int key_bytes, int bits)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MCU_TRY(ensure_device());
    if (key_bytes != 4 && key_bytes != 8) return MCU_EINVAL;
    Session* s;
    MCU_TRY(default_session(&s));
    size_t kb = (size_t)key_bytes;
    MCU_TRY(s->keys_a.reserve((n + 1) * kb));
    MCU_TRY(s->keys_b.reserve((n + 1) * kb));
    MCU_TRY(s->vals_a.reserve((n + 1) * 4));
    MCU_TRY(s->vals_b.reserve((n + 1) * 4));
    if (n) {
        MCU_CUDA(cudaMemcpyAsync(s->keys_a.p, keys, n * kb, cudaMemcpyHostToDevice, s->stream));
        MCU_CUDA(cudaMemcpyAsync(s->vals_a.p, vals, n * 4, cudaMemcpyHostToDevice, s->stream));
    }
    bool in_a = true;
    if (key_bytes == 4)
        MCU_TRY(radix_sort_pairs<u32>(s->radix, s->keys_a.as<u32>(), s->vals_a.as<u32>(), s->keys_b.as<u32>(), s->vals_b.as<u32>(), n, bits,
                                      false, s->stream, &in_a, nullptr));
    else
        MCU_TRY(radix_sort_pairs<u64>(s->radix, s->keys_a.as<u64>(), s->vals_a.as<u32>(), s->keys_b.as<u64>(), s->vals_b.as<u32>(), n, bits,
                                      false, s->stream, &in_a, nullptr));
    if (n) {
        MCU_CUDA(cudaMemcpyAsync(keys, in_a ? s->keys_a.p : s->keys_b.p, n * kb, cudaMemcpyDeviceToHost, s->stream));
        MCU_CUDA(cudaMemcpyAsync(vals, in_a ? s->vals_a.p : s->vals_b.p, n * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    MCU_CUDA(cudaStreamSynchronize(s->stream));
    MCU_CUDA(cudaGetLastError());
    return MCU_OK;
}

}  // extern "C"
