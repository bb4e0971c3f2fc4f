/* TEST INFRASTRUCTURE ONLY -- never shipped, never loaded by the product.
 *
 * LD_PRELOAD stand-in for the few libmauve_cuda.so entry points the C++ adapters of SURVEY.md 8f-2 and the DP seam
 * (adapters/seams/anchoredpp_batch.cpp) call, answering from the CPU restatement (oracle/libmauve_oracle.so).  It lets the CPU
 * suite run oracle/_ref/dropin_check_next and oracle/_ref/progressiveMauve_cuda -- i.e. the adapters' own host code (2-bit
 * unpacking, temp-file mapping, LCB marshalling, range batching) next to / inside the reference -- in a container without a
 * GPU.  The GPU tests run the same binaries against the real library.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <time.h>
static double g_stub_seconds = 0;
static double stub_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct { int64_t len, start0, start1; } mcu_match;
long long orc_sol_build(const char* seq, uint64_t n, uint64_t seed, float* freq_out);
int orc_anchor_scores(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, const float* freq0, const float* freq1, const mcu_match* m,
                      uint64_t n_matches, const uint64_t* lcb_off, uint64_t n_lcb, const int* matrix, int penalize_repeats, double* lcb_score_out,
                      int64_t* match_score_out);

int mcu_init(int device) { (void)device; return 0; }
const char* mcu_last_error(void) { return "stub"; }
int mcu_sol_build(const char* seq, uint64_t n, uint64_t seed, float* freq_out) { return orc_sol_build(seq, n, seed, freq_out) == (long long)n ? 0 : -3; }
int mcu_anchor_scores(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, const float* freq0, const float* freq1,
                      const mcu_match* rows, uint64_t n_rows, const uint64_t* lcb_off, uint64_t n_lcb, const int32_t* matrix, int penalize_repeats,
                      double* lcb_score_out, int64_t* match_score_out)
{
    static const int32_t hoxd[16] = {91, -114, -31, -123, -114, 100, -125, -31, -31, -125, 100, -114, -123, -31, -114, 91};  /* LM/SubstitutionMatrix.h:23-33 */
    float *f0 = NULL, *f1 = NULL;
    int rc;
    /* frequencies not given: built from the seed, as the library does (csrc/sol.cu anchor_scores) */
    if (!freq0) { f0 = (float*)malloc((n0 + 1) * sizeof(float)); if (orc_sol_build(seq0, n0, seed, f0) != (long long)n0) { free(f0); return -3; } freq0 = f0; }
    if (!freq1) { f1 = (float*)malloc((n1 + 1) * sizeof(float)); if (orc_sol_build(seq1, n1, seed, f1) != (long long)n1) { free(f0); free(f1); return -3; } freq1 = f1; }
    rc = orc_anchor_scores(seq0, n0, seq1, n1, freq0, freq1, rows, n_rows, lcb_off, n_lcb, matrix ? matrix : hoxd, penalize_repeats, lcb_score_out, match_score_out) == 0 ? 0 : -3;
    free(f0);
    free(f1);
    return rc;
}

long long orc_nw_align(const char* a, unsigned la, const char* b, unsigned lb, char* path_out, int64_t* score_out);
static unsigned long long g_nw_calls = 0, g_nw_problems = 0;
int mcu_nw_batch(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off, const uint64_t* path_off, char* path_out,
                 uint32_t* path_len, int64_t* score, float* device_ms)
{
    uint64_t i;
    const double t0 = stub_now();
    ++g_nw_calls;
    g_nw_problems += n;
    for (i = 0; i < n; ++i) {
        long long r = orc_nw_align(a + a_off[i], (unsigned)(a_off[i + 1] - a_off[i]), b + b_off[i], (unsigned)(b_off[i + 1] - b_off[i]),
                                   path_out + path_off[i], &score[i]);
        if (r < 0) return -6;
        path_len[i] = (uint32_t)r;
    }
    if (device_ms) *device_ms = 0;
    g_stub_seconds += stub_now() - t0;
    return 0;
}
#include <stdio.h>
#include <stdlib.h>
__attribute__((destructor)) static void stub_report(void)
{
    if (getenv("MAUVE_CUDA_SEAM_REPORT")) fprintf(stderr, "stub: %llu mcu_nw_batch calls, %llu problems; %.2f s inside the stand-in entry points\n", g_nw_calls, g_nw_problems, g_stub_seconds);
}

/* match finder (adapters/seams/memhash_seam.cpp, CudaMatchFinder.h) */
long long orc_find_mums(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match** out, uint64_t* stats);
int mcu_find_mums(const char* seq0, uint64_t n0, const char* seq1, uint64_t n1, uint64_t seed, int rule, mcu_match** out, uint64_t* n_out, uint64_t* stats)
{
    uint64_t st[4] = {0, 0, 0, 0};
    const double t0 = stub_now();
    long long n = orc_find_mums(seq0, n0, seq1, n1, seed, rule, out, st);
    g_stub_seconds += stub_now() - t0;
    if (n < 0) return -3;
    *n_out = (uint64_t)n;
    if (stats) { int i; for (i = 0; i < 8; ++i) stats[i] = 0; stats[0] = st[3]; stats[1] = st[1]; stats[2] = st[0]; }
    return 0;
}
void mcu_free(void* p) { free(p); }

/* sorted mer list (adapters/seams/filesml_seam.cpp, CudaDNAMemorySML.h) */
long long orc_sml_build(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* mer_out);
int orc_pack(const char* seq, uint64_t n, uint32_t* out);
int mcu_sml_build(const char* seq, uint64_t n, uint64_t seed, uint32_t* pos_out, uint64_t* mer_out, uint32_t* packed_out, uint64_t* sml_len_out)
{
    const double t0 = stub_now();
    long long r = orc_sml_build(seq, n, seed, pos_out, mer_out);
    g_stub_seconds += stub_now() - t0;
    if (r < 0) return -3;
    if (packed_out && orc_pack(seq, n, packed_out) != 0) return -4;
    if (sml_len_out) *sml_len_out = (uint64_t)r;
    return 0;
}

/* regions with DNA wildcard columns (CudaGlobalAlign.h, seams with MAUVE_CUDA_WILD=1) */
long long orc_eliminate_overlaps(const mcu_match* rows, uint64_t n, int eliminate_both, uint64_t min_length, mcu_match* out, uint64_t* ties);
long long orc_lcbs(const mcu_match* rows, uint64_t n, mcu_match* sorted_out, uint64_t* bp_out, uint64_t* ties);
int mcu_eliminate_overlaps(const mcu_match* rows, uint64_t n, int eliminate_both, uint64_t min_length, mcu_match* rows_out, uint64_t* n_out, uint64_t* ties_out)
{
    long long k = orc_eliminate_overlaps(rows, n, eliminate_both, min_length, rows_out, ties_out);
    if (k < 0) return -3;
    *n_out = (uint64_t)k;
    return 0;
}
int mcu_lcbs(const mcu_match* rows, uint64_t n, mcu_match* sorted_out, uint64_t* bp_out, uint64_t* n_bp_out, uint64_t* ties_out)
{
    long long k = orc_lcbs(rows, n, sorted_out, bp_out, ties_out);
    if (k < 0) return -3;
    *n_bp_out = (uint64_t)k;
    return 0;
}

/* anchor columns of windows (CudaAnchorCols.h, the AnchoredProfileProfile seam): mcu_anchor_params == orc_anchor_params */
typedef struct { float subst[4][4]; float gap_open, gap_extend, term_gap, smooth_ceil, min_best_col, min_smooth; unsigned smooth_window, anchor_spacing; unsigned char letter_of_char[256]; } mcu_anchor_params;
void orc_anchor_default_params(mcu_anchor_params* p);
long long orc_anchor_cols(const unsigned char* rows, unsigned n1, unsigned n2, unsigned ncol, const float* w, const mcu_anchor_params* p,
                          unsigned* cols_out, float* score_out, float* smooth_out);
void mcu_anchor_default_params(mcu_anchor_params* p) { orc_anchor_default_params(p); }
int mcu_anchor_cols_batch(uint64_t n, const char* rows, const uint64_t* row_off, const uint32_t* ncol, const uint32_t* n1, const uint32_t* n2,
                          const float* weights, const mcu_anchor_params* params, const uint64_t* col_off, uint32_t* cols_out,
                          uint32_t* n_cols_out, float* score_out, float* smooth_out, float* device_ms)
{
    mcu_anchor_params dflt;
    uint64_t i, woff = 0;
    if (!params) { orc_anchor_default_params(&dflt); params = &dflt; }
    for (i = 0; i < n; ++i) {
        long long k = orc_anchor_cols((const unsigned char*)rows + row_off[i], n1[i], n2[i], ncol[i], weights ? weights + woff : NULL, params,
                                      cols_out + col_off[i], score_out ? score_out + col_off[i] : NULL, smooth_out ? smooth_out + col_off[i] : NULL);
        if (k < 0) return -3;
        n_cols_out[i] = (uint32_t)k;
        woff += (uint64_t)n1[i] + n2[i];
    }
    if (device_ms) *device_ms = 1.0f;
    return 0;
}

long long orc_nw_align_f(const char* a, unsigned la, const char* b, unsigned lb, char* path_out, float* score_out);
static unsigned long long g_nwf_problems = 0;
int orc_hmm_run(const char* sym, uint64_t len, const double* p, char* pred_out, double* post_out);
int mcu_hmm_batch(uint64_t n, const char* sym, const uint64_t* off, const double* params, char* pred_out, double* post_out, float* device_ms)
{
    uint64_t i;
    for (i = 0; i < n; ++i)
        if (orc_hmm_run(sym + off[i], off[i + 1] - off[i], params, pred_out + off[i], post_out ? post_out + off[i] : NULL) != 0) return -3;
    if (device_ms) *device_ms = 1.0f;
    return 0;
}

int mcu_nw_batch_wild(uint64_t n, const char* a, const uint64_t* a_off, const char* b, const uint64_t* b_off, const uint64_t* path_off, char* path_out,
                      uint32_t* path_len, float* score, float* device_ms)
{
    uint64_t i;
    const double t0 = stub_now();
    g_nwf_problems += n;
    for (i = 0; i < n; ++i) {   /* the argument checks of nw_batch_wild (csrc/dpwild.cu): empty region, sequence too long */
        const uint64_t la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        if (a_off[i + 1] <= a_off[i] || b_off[i + 1] <= b_off[i] || la > 0x3FFFFFFull || lb > 0x3FFFFFFull) return -3;   /* MCU_EINVAL */
    }
    for (i = 0; i < n; ++i) {
        long long r = orc_nw_align_f(a + a_off[i], (unsigned)(a_off[i + 1] - a_off[i]), b + b_off[i], (unsigned)(b_off[i + 1] - b_off[i]),
                                     path_out + path_off[i], &score[i]);
        if (r < 0) return -6;
        path_len[i] = (uint32_t)r;
    }
    if (device_ms) *device_ms = 0;
    g_stub_seconds += stub_now() - t0;
    return 0;
}
__attribute__((destructor)) static void stub_report_wild(void)
{
    if (getenv("MAUVE_CUDA_SEAM_REPORT") && g_nwf_problems) fprintf(stderr, "stub: %llu mcu_nw_batch_wild problems\n", g_nwf_problems);
}
