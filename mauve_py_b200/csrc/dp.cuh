// Gapped DP (NWSmall + BitTraceBack restated for two single-sequence ACGT profiles).
#pragma once
#include "common.cuh"

namespace mcu {

int nw_batch(u64 n, const char* a, const u64* a_off, const char* b, const u64* b_off, const u64* path_off, char* path_out,
             u32* path_len, i64* score, float* device_ms);

// counters bench.py reads through mcu_dp_last_stats: [0] cells, [1] forward launches, [2] traceback launches,
// [3] sub-batches, [4] traceback bytes
void nw_last_stats(u64* out5);

// NWSmall in the reference's float arithmetic for regions with DNA wildcard columns (dpwild.cu): one thread per region
int nw_batch_wild(u64 n, const char* a, const u64* a_off, const char* b, const u64* b_off, const u64* path_off, char* path_out, u32* path_len,
                  float* score, float* device_ms);

// register-resident add/max microbenchmark: Gops/s of thread-level INT32 instructions the device sustains (DP roofline denominator)
int int32_peak(double* gops_out, float* ms_out);

void nw_release();
void nwf_release();

}  // namespace mcu
