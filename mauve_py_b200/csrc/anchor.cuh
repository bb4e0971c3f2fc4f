// Device pipeline of the anchoring path: pack -> seedgen -> sort -> join -> extend -> order.
#pragma once
#include "common.cuh"
#include "radix.cuh"

namespace mcu {

struct Session {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    DevBuf ascii[2], packed[2];
    DevBuf keys_a, keys_b, vals_a, vals_b;
    DevBuf uniq, pairs, cand, raw_matches, ord_keys_a, ord_keys_b, ord_vals_a, ord_vals_b, ord_primary, matches, counters;
    RadixScratch radix;
    u64 n[2] = {0, 0};
    u64 match_count = 0;
    u64 launches = 0;
    unsigned long long* h_counters = nullptr;  // pinned, 8 entries
    bool ok = false;
};

int session_init(Session& s);
void session_destroy(Session& s);
int session_upload(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1);
int session_run(Session& s, u64 seed, int shard_index, int shard_count, float* stage_ms, u64* stats);
// sorts rows (device, n of them) into reference list order; result in s.matches (device)
int order_matches(Session& s, const mcu_match* rows_dev, u64 n);

// single-genome SML (stable): outputs on device in s.keys_*/vals_* ; returns which buffer
int sml_build_device(Session& s, const char* seq, u64 n, u64 seed, u32* pos_out, u64* mer_out, u32* packed_out, u64* len_out);

}  // namespace mcu
