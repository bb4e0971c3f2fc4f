/* TEST INFRASTRUCTURE ONLY -- never shipped, never loaded by the product.
 *
 * Stand-in for ALL of libmauve_cuda.so's symbols (include/mauve_cuda.h), answering from the CPU restatement
 * (oracle/libmauve_oracle.so).  tests/test_bench_dryrun.py loads it in place of the product library so that bench.py's and
 * libmems.py's own host code (argument marshalling, roofline arithmetic, JSON line, secondary objects) executes end to end in a
 * container without a GPU.  Nothing measured through it is a benchmark number.
 * It extends mcu_stub.c (the LD_PRELOAD stand-in of the C++ adapters) with the session, HMM, seed-table and test entry points. */
#define mcu_nw_batch mcu_nw_batch_base
#include "mcu_stub.c"
#undef mcu_nw_batch

#include <string.h>

uint64_t orc_get_seed(int weight, int rank);
int orc_seed_length(uint64_t seed);
int orc_seed_weight(uint64_t seed);
unsigned orc_default_seed_weight(uint64_t avg_len);
void orc_hmm_params(double gc, double go_homologous, double go_unrelated, double pct_id, double* out);
int orc_hmm_run(const char* sym, uint64_t len, const double* p, char* pred_out, double* post_out);

void mcu_shutdown(void) {}
int mcu_host_alloc(void** out, uint64_t bytes) { *out = malloc(bytes ? bytes : 1); return *out ? 0 : -5; }
void mcu_host_free(void* p) { free(p); }
uint64_t mcu_get_seed(int weight, int rank) { return orc_get_seed(weight, rank); }
unsigned mcu_default_seed_weight(uint64_t n) { return orc_default_seed_weight(n); }
int mcu_seed_length(uint64_t seed) { return orc_seed_length(seed); }
int mcu_seed_weight(uint64_t seed) { return orc_seed_weight(seed); }

typedef struct {
    const char* seq[2];
    uint64_t n[2];
    mcu_match* rows;
    uint64_t n_rows;
    uint64_t launches;
    uint64_t seed;
    int shard;
} stub_session;

int mcu_session_create(void** out) { *out = calloc(1, sizeof(stub_session)); return *out ? 0 : -5; }
void mcu_session_destroy(void* h) { stub_session* s = (stub_session*)h; if (s) { free(s->rows); free(s); } }
int mcu_session_upload(void* h, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1)
{
    stub_session* s = (stub_session*)h;
    s->seq[0] = seq0; s->n[0] = n0; s->seq[1] = seq1; s->n[1] = n1;
    return 0;
}
int mcu_session_run(void* h, uint64_t seed, int shard_index, int shard_count, float* stage_ms, uint64_t* stats)
{
    stub_session* s = (stub_session*)h;
    uint64_t st[4] = {0, 0, 0, 0};
    long long n;
    int i;
    if (shard_count != 1 || shard_index != 0) return -3;
    free(s->rows);
    s->rows = NULL;
    n = orc_find_mums(s->seq[0], s->n[0], s->seq[1], s->n[1], seed, 0, &s->rows, st);
    if (n < 0) return -3;
    s->n_rows = (uint64_t)n;
    s->launches += 17;
    if (stage_ms) {
        for (i = 0; i < 16; ++i) stage_ms[i] = 0.f;
        for (i = 0; i < 6; ++i) stage_ms[i] = 0.5f;
        stage_ms[6] = 3.0f;
        stage_ms[7] = -1.0f;  /* bucketed enumeration */
        for (i = 8; i < 15; ++i) stage_ms[i] = 0.25f;
    }
    if (stats) {
        const int L = orc_seed_length(seed);
        for (i = 0; i < 8; ++i) stats[i] = 0;
        stats[0] = st[3]; stats[1] = (uint64_t)n; stats[2] = st[0]; stats[4] = (uint64_t)n;
        stats[5] = (s->n[0] >= (uint64_t)L ? s->n[0] - L + 1 : 0) + (s->n[1] >= (uint64_t)L ? s->n[1] - L + 1 : 0);
    }
    return 0;
}
/* the two-phase form (multi-GPU runs): rank 0's stand-in produces the whole list, the other ranks none; merge keeps what it is given */
static void stub_fill(stub_session* s, uint64_t seed, long long n, const uint64_t* st, float* stage_ms, uint64_t* stats)
{
    int i;
    s->launches += 17;
    if (stage_ms) {
        for (i = 0; i < 16; ++i) stage_ms[i] = 0.f;
        for (i = 0; i < 6; ++i) stage_ms[i] = 0.5f;
        stage_ms[6] = 3.0f;
        stage_ms[7] = -1.0f;  /* bucketed enumeration */
        for (i = 8; i < 15; ++i) stage_ms[i] = 0.25f;
    }
    if (stats) {
        const int L = orc_seed_length(seed);
        for (i = 0; i < 8; ++i) stats[i] = 0;
        stats[0] = st[3]; stats[1] = (uint64_t)n; stats[2] = st[0]; stats[4] = (uint64_t)n;
        stats[5] = (s->n[0] >= (uint64_t)L ? s->n[0] - L + 1 : 0) + (s->n[1] >= (uint64_t)L ? s->n[1] - L + 1 : 0);
    }
}
int mcu_session_enumerate(void* h, uint64_t seed, int shard_index, int shard_count)
{
    stub_session* s = (stub_session*)h;
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return -3;
    s->seed = seed; s->shard = shard_index;
    return 0;
}
int mcu_session_uniq_bitmap(void* h, void** p, uint64_t* n) { static uint32_t words[4]; (void)h; *p = words; *n = 4; return 0; }
int mcu_session_finish(void* h, int uniq_is_global, float* stage_ms, uint64_t* stats)
{
    stub_session* s = (stub_session*)h;
    uint64_t st[4] = {0, 0, 0, 0};
    long long n = 0;
    (void)uniq_is_global;
    free(s->rows);
    s->rows = NULL;
    if (s->shard == 0) n = orc_find_mums(s->seq[0], s->n[0], s->seq[1], s->n[1], s->seed, 0, &s->rows, st);
    if (n < 0) return -3;
    s->n_rows = (uint64_t)n;
    stub_fill(s, s->seed, n, st, stage_ms, stats);
    return 0;
}
int mcu_session_merge(void* h, const mcu_match* r, uint64_t n, int in_device, uint64_t* st)
{
    stub_session* s = (stub_session*)h;
    mcu_match* copy = (mcu_match*)malloc((n + 1) * sizeof(mcu_match));
    (void)in_device;
    if (n) memcpy(copy, r, n * sizeof(mcu_match));
    free(s->rows);
    s->rows = copy;
    s->n_rows = n;
    if (st) st[0] = st[1] = 0;
    return 0;
}
uint64_t mcu_session_match_count(const void* h) { return ((const stub_session*)h)->n_rows; }
int mcu_session_download(void* h, mcu_match* out)
{
    stub_session* s = (stub_session*)h;
    if (s->n_rows) memcpy(out, s->rows, s->n_rows * sizeof(mcu_match));
    return 0;
}
const void* mcu_session_matches_device(const void* h) { return ((const stub_session*)h)->rows; }
uint64_t mcu_session_launch_count(const void* h) { return ((const stub_session*)h)->launches; }
int mcu_merge_matches(const mcu_match* r, uint64_t n, int d, mcu_match** out, uint64_t* n_out, uint64_t* u)
{
    (void)r; (void)n; (void)d; (void)out; (void)n_out; (void)u;
    return -1;
}
int mcu_find_mums_batch(uint64_t n, const void* a, const void* b, const void* c, const void* d, const void* e, int rule, mcu_match** out, void* f, void* g)
{
    (void)n; (void)a; (void)b; (void)c; (void)d; (void)e; (void)rule; (void)out; (void)f; (void)g;
    return -1;
}

static uint64_t g_last_cells = 0, g_last_n = 0;
int mcu_nw_batch(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off, const uint64_t* path_off, char* path_out,
                 uint32_t* path_len, int64_t* score, float* device_ms)
{
    uint64_t i;
    const int rc = mcu_nw_batch_base(n, a, a_off, b, b_off, path_off, path_out, path_len, score, device_ms);
    g_last_cells = 0;
    for (i = 0; i < n; ++i) g_last_cells += (a_off[i + 1] - a_off[i]) * (b_off[i + 1] - b_off[i]);
    g_last_n = n;
    if (device_ms) *device_ms = 1.0f;  /* bench.py divides by it */
    return rc;
}
void mcu_nw_last_stats(uint64_t* out5)
{
    out5[0] = g_last_cells; out5[1] = 1; out5[2] = 1; out5[3] = 1; out5[4] = g_last_n;
}

int mcu_hmm_params(double gc, double go_h, double go_u, double pct, double* out) { orc_hmm_params(gc, go_h, go_u, pct, out); return 0; }
int mcu_test_sort_pairs(void* k, void* v, uint64_t n, int bits, int kb) { (void)k; (void)v; (void)n; (void)bits; (void)kb; return -1; }
int mcu_test_int32_peak(double* gops, float* ms) { if (gops) *gops = 1000.0; if (ms) *ms = 1.0f; return 0; }
int mcu_sml_build_shard(const char* seq, uint64_t n, uint64_t seed, int shard, int n_shards, uint32_t* pos_out, uint64_t* mer_out, uint64_t* len_out)
{
    (void)shard; (void)n_shards;   /* the stand-in has one shard */
    return mcu_sml_build(seq, n, seed, pos_out, mer_out, NULL, len_out);
}
int mcu_sml_build_sharded(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* len_out, float* ms_out)
{
    if (ms_out) *ms_out = 1.0f;
    return mcu_sml_build(seq, n, seed, pos_out, NULL, NULL, len_out);
}
int mcu_test_hmm_counters(uint64_t* out3) { if (out3) out3[0] = out3[1] = out3[2] = 0; return 0; }
int mcu_test_anchor_counters(uint64_t* out8) { int i; if (out8) for (i = 0; i < 8; ++i) out8[i] = 0; return 0; }

/* ---- round 2 entry points: caller-buffer form, chunked upload, the library's own communicator ------------------------------------ */
#include <stdio.h>
#include <unistd.h>
#include <sys/stat.h>

int mcu_find_mums_into(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match* rows_out, uint64_t cap,
                       uint64_t* n_out, uint64_t* stats)
{
    mcu_match* r = NULL;
    uint64_t n = 0;
    const int rc = mcu_find_mums(seq0, n0, seq1, n1, seed, rule, &r, &n, stats);
    if (rc) return rc;
    *n_out = n;
    if (n > cap) { free(r); return -7; }
    if (n) memcpy(rows_out, r, n * sizeof(mcu_match));
    free(r);
    return 0;
}
void mcu_sml_last_stats(float* out6) { out6[0] = 0.1f; out6[1] = 1.0f; out6[2] = 2.0f; out6[3] = 5.0f; out6[4] = 8.0f; out6[5] = 0.0f; }
int mcu_session_upload_begin(void* h, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, int chunks)
{
    (void)chunks;
    return mcu_session_upload(h, seq0, n0, seq1, n1);
}
int mcu_device_synchronize(void) { return 0; }

/* communicator over files: every collective writes one file per rank and reads the others' (same directory, same sequence number) */
static struct { int rank, world, ok; unsigned long long seq; char tag[80]; } g_sc;
int mcu_comm_unique_id(void* id_out)
{
    memset(id_out, 0, 128);
    snprintf((char*)id_out, 64, "%ld_%ld", (long)getpid(), (long)time(NULL));
    return 0;
}
int mcu_comm_init(int rank, int world, const void* id)
{
    g_sc.rank = rank; g_sc.world = world; g_sc.ok = 1; g_sc.seq = 0;
    snprintf(g_sc.tag, sizeof g_sc.tag, "/tmp/mcu_stub_comm_%s", world > 1 ? (const char*)id : "solo");
    return 0;
}
void mcu_comm_destroy(void) { g_sc.ok = 0; }
int mcu_comm_rank(void) { return g_sc.ok ? g_sc.rank : 0; }
int mcu_comm_world(void) { return g_sc.ok ? g_sc.world : 1; }
static void sc_path(char* out, size_t cap, unsigned long long seq, int rank) { snprintf(out, cap, "%s_%llu_%d", g_sc.tag, seq, rank); }
/* all-gather of one blob per rank: out[r] = malloc'ed copy, len[r] */
static int sc_exchange(const void* mine, uint64_t n, void** out, uint64_t* len)
{
    char path[200], tmp[220];
    int r;
    const unsigned long long seq = g_sc.seq++;
    sc_path(path, sizeof path, seq, g_sc.rank);
    snprintf(tmp, sizeof tmp, "%s.tmp", path);
    FILE* f = fopen(tmp, "wb");
    if (!f) return -2;
    if (n) fwrite(mine, 1, n, f);
    fclose(f);
    rename(tmp, path);
    for (r = 0; r < g_sc.world; ++r) {
        struct stat st;
        int tries = 0;
        sc_path(path, sizeof path, seq, r);
        while (stat(path, &st) != 0) { usleep(2000); if (++tries > 150000) return -2; }
        len[r] = (uint64_t)st.st_size;
        out[r] = malloc(len[r] ? len[r] : 1);
        f = fopen(path, "rb");
        if (len[r] && fread(out[r], 1, len[r], f) != len[r]) { fclose(f); return -2; }
        fclose(f);
    }
    /* second phase: nobody removes a file before every rank has read all of them */
    {
        const unsigned long long seq2 = g_sc.seq++;
        sc_path(path, sizeof path, seq2, g_sc.rank);
        f = fopen(path, "wb"); fclose(f);
        for (r = 0; r < g_sc.world; ++r) {
            struct stat st;
            int tries = 0;
            sc_path(path, sizeof path, seq2, r);
            while (stat(path, &st) != 0) { usleep(2000); if (++tries > 150000) return -2; }
        }
        sc_path(path, sizeof path, seq, g_sc.rank);
        remove(path);
    }
    return 0;
}
int mcu_comm_allreduce_f64(double* v, int n, int op)
{
    void* out[64];
    uint64_t len[64];
    int r, i;
    if (!g_sc.ok) return -3;
    if (g_sc.world == 1 || n == 0) return 0;
    if (sc_exchange(v, (uint64_t)n * sizeof(double), out, len)) return -2;
    for (i = 0; i < n; ++i) {
        double acc = ((double*)out[0])[i];
        for (r = 1; r < g_sc.world; ++r) {
            const double x = ((double*)out[r])[i];
            acc = op == 0 ? acc + x : op == 1 ? (x > acc ? x : acc) : (x < acc ? x : acc);
        }
        v[i] = acc;
    }
    for (r = 0; r < g_sc.world; ++r) free(out[r]);
    return 0;
}
int mcu_comm_barrier(void) { double one = 1.0; return g_sc.ok ? mcu_comm_allreduce_f64(&one, 1, 0) : -3; }
int mcu_comm_gather_bytes(const void* send, uint64_t n, void** out, uint64_t* counts_out)
{
    void* parts[64];
    uint64_t len[64], total = 0, pos = 0;
    int r;
    if (!g_sc.ok) return -3;
    *out = NULL;
    if (g_sc.world == 1) { parts[0] = malloc(n ? n : 1); if (n) memcpy(parts[0], send, n); len[0] = n; }
    else if (sc_exchange(send, n, parts, len)) return -2;
    for (r = 0; r < g_sc.world; ++r) { total += len[r]; if (counts_out) counts_out[r] = len[r]; }
    if (g_sc.rank == 0) {
        char* cat = (char*)malloc(total ? total : 1);
        for (r = 0; r < g_sc.world; ++r) { if (len[r]) memcpy(cat + pos, parts[r], len[r]); pos += len[r]; }
        *out = cat;
    }
    for (r = 0; r < g_sc.world; ++r) free(parts[r]);
    return 0;
}
int mcu_session_upload_sharded(void* h, const char* seq0, uint64_t n0, const char* seq1, uint64_t n1) { return mcu_session_upload(h, seq0, n0, seq1, n1); }
/* every rank's stand-in computes the whole list (they hold the whole genomes); rank 0's is "the merged one" */
int mcu_session_run_sharded(void* h, uint64_t seed, float* stage_ms, uint64_t* stats) { return mcu_session_run(h, seed, 0, 1, stage_ms, stats); }
int mcu_find_mums_sharded(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match* rows_out, uint64_t cap,
                          uint64_t* n_out, uint64_t* stats)
{
    mcu_match* r = NULL;
    uint64_t n = 0;
    const int rc = mcu_find_mums(seq0, n0, seq1, n1, seed, rule, &r, &n, stats);
    if (rc) return rc;
    *n_out = n;
    if (g_sc.rank == 0) {
        if (n > cap) { free(r); return -7; }
        if (n) memcpy(rows_out, r, n * sizeof(mcu_match));
    }
    free(r);
    return 0;
}
