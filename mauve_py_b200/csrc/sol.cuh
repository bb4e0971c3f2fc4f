// Seed occurrence list + anchor scores (sol.cu): the consumers of the sorted mer list after matching.
#pragma once
#include "anchor.cuh"

namespace mcu {

// SeedOccurrenceList::construct (LM/SeedOccurrenceList.h:22-78) of one sequence: n floats to the host
int sol_build(Session& s, const char* seq, u64 n, u64 seed, float* freq_out);
// the same, left on the device in s.sol_freq[slot]
int sol_build_device(Session& s, const char* seq, u64 n, u64 seed, int slot);
// GetPairwiseAnchorScore (LM/GreedyBreakpointElimination.h:403-476) for the LCBs of one genome pair
int anchor_scores(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1, u64 seed, const float* freq0, const float* freq1,
                  const mcu_match* rows, u64 n_rows, const u64* lcb_off, u64 n_lcb, const int* matrix, int penalize_repeats,
                  double* lcb_score_out, i64* match_score_out);

}  // namespace mcu
