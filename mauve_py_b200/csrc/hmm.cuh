// Pairwise HomologyHMM: Forward + Backward posteriors of the 2-state model.
#pragma once
#include "common.cuh"

namespace mcu {

int hmm_params(double gc, double go_homologous, double go_unrelated, double pct_identity, double* out21);
int hmm_batch(u64 n, const char* sym, const u64* off, const double* params, char* pred_out, double* post_out, float* device_ms);
void hmm_release();
void hmm_last_counters(u64* out3);  // warp chains of the last call: columns, chain rounds, columns evaluated by hmm_exact_step

}  // namespace mcu
