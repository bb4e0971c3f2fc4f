// The step after the match list for two genomes: EliminateOverlaps_v2 + LengthFilter, IdentifyBreakpoints + ComputeLCBs_v2.
#pragma once
#include "common.cuh"

namespace mcu {

int lcb_eliminate_overlaps(const mcu_match* rows, u64 n, int eliminate_both, u64 min_length, mcu_match* rows_out, u64* n_out, u64* ties_out);
int lcb_breakpoints(const mcu_match* rows, u64 n, mcu_match* sorted_out, u64* bp_out, u64* n_bp_out, u64* ties_out);

void lcb_release();

}  // namespace mcu
