/* TEST INFRASTRUCTURE ONLY -- never shipped, never loaded by the product.
 *
 * LD_PRELOAD stand-in for the few libmauve_cuda.so entry points the C++ adapters of SURVEY.md 8f-2 call, answering from the
 * CPU restatement (oracle/libmauve_oracle.so).  It lets the CPU suite run oracle/_ref/dropin_check_next -- i.e. the adapters'
 * own host code (2-bit unpacking, temp-file mapping, LCB marshalling) next to the reference classes -- in a container
 * without a GPU.  The GPU tests run the same binary against the real library.
 */
#include <stdint.h>
#include <stddef.h>

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
    (void)seed;
    if (!freq0 || !freq1) return -3;
    return orc_anchor_scores(seq0, n0, seq1, n1, freq0, freq1, rows, n_rows, lcb_off, n_lcb, matrix, penalize_repeats, lcb_score_out, match_score_out) == 0 ? 0 : -3;
}
