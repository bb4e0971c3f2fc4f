// Stable LSD radix sort of (key, u32 value) pairs, 8 bits per pass, one global read + one
// global write of every pair per pass ("onesweep": chained-scan with decoupled look-back, so
// the digit histogram of a tile is combined with its predecessors' while the tile is still in
// shared memory).  Replaces std::sort(bmer_lessthan) in MemorySML::Create
// (LM/MemorySML.cpp:54) / FileSML::Create (LM/FileSML.cpp:429).  HBM-bound by construction:
// per pass 2*(sizeof(K)+4) bytes per pair.
//
// Layout per tile (TILE = BLOCK*IPT pairs, "warp-striped": warp w owns the contiguous chunk
// [w*32*IPT, (w+1)*32*IPT), lane l item i sits at chunk + i*32 + l, so every load instruction
// of a warp is one contiguous 128/256-byte request):
//   1. rank keys inside the warp with match.any on the digit (stable: item-major, lane-minor)
//   2. per-digit exclusive scan over warps -> tile digit counts
//   3. publish counts, look back over earlier tiles (status word = epoch|state|value), publish
//      inclusive prefix
//   4. place keys/values at their in-tile sorted slot in shared memory, then stream them out:
//      consecutive threads write consecutive addresses inside each digit run (coalesced).
#pragma once
#include "common.cuh"

namespace mcu {

constexpr int RS_BLOCK = 512;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_RADIX = 256;

template <typename K>
struct RsCfg {
    static constexpr int IPT = sizeof(K) == 4 ? 16 : 12;
    static constexpr int TILE = RS_BLOCK * IPT;
    static constexpr size_t SMEM = (size_t)TILE * (sizeof(K) + 4) + RS_WARPS * RS_RADIX * 4 + RS_RADIX * (4 + 8) + 64;
};

// status word: [63:50] epoch (14 bits) | [49:48] state | [47:0] value
constexpr u64 RS_STATE_AGG = 1ull << 48;
constexpr u64 RS_STATE_PREFIX = 2ull << 48;
constexpr u64 RS_VALUE_MASK = (1ull << 48) - 1;
constexpr int RS_EPOCH_SHIFT = 50;
constexpr u32 RS_EPOCH_MAX = (1u << 14) - 1;

// Histogram of every digit position in one read of the keys (used when the producer of the
// keys did not already accumulate it).  hist[pass*256 + digit].
template <typename K>
__global__ void __launch_bounds__(256) rs_histogram_kernel(const K* __restrict__ keys, u64 n, int passes, int first_shift, u64* __restrict__ hist)
{
    extern __shared__ u32 sh[];  // passes*256
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        K k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * RS_RADIX + (u32)((k >> (first_shift + 8 * p)) & 255)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd((unsigned long long*)&hist[i], (unsigned long long)sh[i]);
}

// In-place exclusive scan of each pass's 256-bin histogram; one block of 256 threads per pass.
__global__ void rs_scan_hist_kernel(u64* hist);

template <typename K>
__global__ void __launch_bounds__(RS_BLOCK, 2) rs_onesweep_kernel(const K* __restrict__ kin, const u32* __restrict__ vin,
                                                              K* __restrict__ kout, u32* __restrict__ vout, u64 n, int shift,
                                                              const u64* __restrict__ ghist_excl, volatile u64* status,
                                                              u32* tile_counter, u32 epoch)
{
    constexpr int IPT = RsCfg<K>::IPT;
    constexpr int TILE = RsCfg<K>::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* skeys = (K*)smem_raw;
    u32* svals = (u32*)(smem_raw + (size_t)TILE * sizeof(K));
    u32* whist = svals + TILE;                        // [RS_WARPS][256]
    u32* tile_excl = whist + RS_WARPS * RS_RADIX;     // [256] exclusive scan of tile digit counts
    u64* gbase = (u64*)(tile_excl + RS_RADIX);        // [256] global destination of slot 0 of each digit run
    u32* s_misc = (u32*)(gbase + RS_RADIX);           // [0] tile id, [1..16] warp scan scratch

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_misc[0] = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_BLOCK) whist[i] = 0;
    __syncthreads();
    const u32 tile = s_misc[0];
    const u64 tile_base = (u64)tile * TILE;
    const u64 remaining = n - tile_base;
    const u32 tile_n = remaining < (u64)TILE ? (u32)remaining : (u32)TILE;

    // ---- load (warp-striped) ----
    K key[IPT];
    const u32 chunk = warp * 32 * IPT;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 loc = chunk + i * 32 + lane;
        key[i] = loc < tile_n ? __ldcs(kin + tile_base + loc) : (K)~(K)0;   // read once: streaming, so that the lines below stay in L1
    }
    // the values are not touched before the keys are ranked and the tile's place is known: their lines are asked for now (one 128-byte
    // line per warp and item, no register spent), and the loads of the placement step find them in L1
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (lane == 0 && chunk + i * 32 < tile_n) asm volatile("prefetch.global.L1 [%0];" ::"l"(vin + tile_base + chunk + i * 32));

    // ---- rank inside the warp ----
    u32 rank[IPT];
    u32* myhist = whist + warp * RS_RADIX;
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 loc = chunk + i * 32 + lane;
        bool valid = loc < tile_n;
        u32 d = (u32)((key[i] >> shift) & 255);
        u32 peers = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
        u32 leader = __ffs(peers) - 1;
        u32 old = 0;
        if (valid && lane == leader) {
            old = myhist[d];
            myhist[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive scan over warps, tile totals ----
    u32 my_count = 0;
    if (tid < RS_RADIX) {
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            u32 c = whist[w * RS_RADIX + tid];
            whist[w * RS_RADIX + tid] = run;
            run += c;
        }
        my_count = run;
    }
    // block exclusive scan of my_count over the 256 digit threads (warps 0..7)
    u32 incl = my_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (tid < RS_RADIX && lane == 31) s_misc[1 + warp] = incl;
    __syncthreads();
    if (tid < RS_RADIX) {
        u32 add = 0;
        for (u32 w = 0; w < warp; ++w) add += s_misc[1 + w];
        u32 excl = incl - my_count + add;
        tile_excl[tid] = excl;

        // ---- decoupled look-back on this digit ----
        const u64 tag = (u64)epoch << RS_EPOCH_SHIFT;
        u64 prefix = 0;
        if (tile == 0) {
            status[(u64)tid] = tag | RS_STATE_PREFIX | (u64)my_count;
        } else {
            status[(u64)tile * RS_RADIX + tid] = tag | RS_STATE_AGG | (u64)my_count;
            i64 look = (i64)tile - 1;
            while (true) {
                u64 s = status[(u64)look * RS_RADIX + tid];
                if ((s >> RS_EPOCH_SHIFT) != (u64)epoch || (s & (RS_STATE_AGG | RS_STATE_PREFIX)) == 0) continue;  // not published yet
                prefix += s & RS_VALUE_MASK;
                if (s & RS_STATE_PREFIX) break;
                --look;
            }
            status[(u64)tile * RS_RADIX + tid] = tag | RS_STATE_PREFIX | (prefix + my_count);
        }
        gbase[tid] = ghist_excl[tid] + prefix - excl;  // dst(slot j of digit d) = gbase[d] + j
    }
    __syncthreads();

    // ---- place into shared memory at the in-tile sorted slot ----
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 loc = chunk + i * 32 + lane;
        if (loc < tile_n) {
            u32 d = (u32)((key[i] >> shift) & 255);
            u32 slot = tile_excl[d] + myhist[d] + rank[i];
            skeys[slot] = key[i];
            svals[slot] = vin[tile_base + loc];  // values are only touched here: keeps them out of the ranking registers
        }
    }
    __syncthreads();

    // ---- stream out ----
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 j = i * RS_BLOCK + tid;
        if (j < tile_n) {
            K k = skeys[j];
            u32 d = (u32)((k >> shift) & 255);
            u64 dst = gbase[d] + j;
            kout[dst] = k;
            vout[dst] = svals[j];
        }
    }
}

// Host-side driver state: scratch shared by all sorts of one context.
struct RadixScratch {
    DevBuf hist;     // [8 passes][256] u64
    DevBuf status;   // [tiles][256] u64
    DevBuf counters; // [16] u32
    u32 epoch = 0;
    u64 launches = 0;
};

// Sorts n pairs on key bits [0, bits).  keys_a/vals_a hold the input; *_b are same-size scratch.
// If `hist_ready`, scratch.hist already holds the (un-scanned) per-pass digit counts.
// Returns through *out_in_a whether the sorted data ended up in the a-buffers.
template <typename K>
int radix_sort_pairs(RadixScratch& sc, K* keys_a, u32* vals_a, K* keys_b, u32* vals_b, u64 n, int bits, bool hist_ready,
                     cudaStream_t stream, bool* out_in_a, int* passes_out);

int radix_clear_hist(RadixScratch& sc, cudaStream_t stream);

}  // namespace mcu
