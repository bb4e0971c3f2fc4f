// Anchor columns of alignment windows: muscle::FindAnchorColsPP on the device.
#pragma once
#include <vector>

#include "common.cuh"

namespace mcu {

void ac_default_params(mcu_anchor_params* p);
int ac_batch(u64 n, const char* rows, const u64* row_off, const u32* ncol, const u32* n1, const u32* n2, const float* weights,
             const mcu_anchor_params* params, const u64* col_off, u32* cols_out, u32* n_cols_out, float* score_out, float* smooth_out, float* device_ms);
void ac_release();
void ac_last_counters(u64* out8);   // of the last call: smoothing segments finished without / with the serial chain, then 6 phase times (SM cycles, first window)

}  // namespace mcu
