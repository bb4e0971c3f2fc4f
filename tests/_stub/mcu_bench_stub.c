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
int mcu_hmm_batch(uint64_t n, const char* sym, const uint64_t* off, const double* params, char* pred_out, double* post_out, float* device_ms)
{
    uint64_t i;
    for (i = 0; i < n; ++i)
        if (orc_hmm_run(sym + off[i], off[i + 1] - off[i], params, pred_out + off[i], post_out ? post_out + off[i] : NULL) != 0) return -3;
    if (device_ms) *device_ms = 1.0f;
    return 0;
}
int mcu_test_sort_pairs(void* k, void* v, uint64_t n, int bits, int kb) { (void)k; (void)v; (void)n; (void)bits; (void)kb; return -1; }
int mcu_test_int32_peak(double* gops, float* ms) { if (gops) *gops = 1000.0; if (ms) *ms = 1.0f; return 0; }
