// Shared host/device plumbing for libmauve_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#include "../../include/mauve_cuda.h"

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef uint8_t u8;

namespace mcu {

// ---- error state ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define MCU_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            mcu::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return MCU_ECUDA;                                                                   \
        }                                                                                       \
    } while (0)

#define MCU_TRY(call)                  \
    do {                               \
        int r__ = (call);              \
        if (r__ != MCU_OK) return r__; \
    } while (0)

int ensure_device();  // MCU_OK or MCU_ENODEV; lazily selects device 0 when mcu_init was not called
int sm_count();

// ---- growable device buffer (cached across calls: cudaMalloc is far slower than the kernels) ----
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return MCU_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            return MCU_ENOMEM;
        }
        cap = want;
        return MCU_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return (T*)p; }
};

static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }

// ---- seed pattern decomposition (host) -> constant-style parameter block passed by value ----
#define MCU_MAX_RUNS 16
struct SeedParams {
    u64 seed;
    int L, w;
    int nruns;
    int palindromic;          // pattern equals its own bit reversal
    u8 shift_in[MCU_MAX_RUNS];   // right shift applied to the 2L-bit window
    u8 shift_out[MCU_MAX_RUNS];  // left shift into the 2w-bit seed
    u64 mask[MCU_MAX_RUNS];      // (1 << 2*runlen) - 1
};
int make_seed_params(u64 seed, SeedParams* out);  // MCU_EINVAL for unusable patterns

// ---- device helpers shared by the kernels ------------------------------------------------
#ifdef __CUDACC__
// 32 bases starting at base `pos`, left-aligned (base pos in bits 63..62). packed is MSB-first
// per 32-bit word with >= 2 zero pad words (SortedMerList::SetSequence layout).
__device__ __forceinline__ u64 load_mer32(const u32* __restrict__ packed, u64 pos)
{
    u64 word = pos >> 4;
    u32 bit = (u32)(pos & 15) * 2;
    u32 w0 = __ldg(packed + word), w1 = __ldg(packed + word + 1), w2 = __ldg(packed + word + 2);
    u64 hi = ((u64)w0 << 32) | w1;
    return bit ? ((hi << bit) | ((u64)w2 >> (32 - bit))) : hi;
}

// 2-bit code of base i
__device__ __forceinline__ u32 base_at(const u32* __restrict__ packed, i64 i) { return (__ldg(packed + (i >> 4)) >> (30 - 2 * (int)(i & 15))) & 3u; }

// spaced seed (2w bits, right-aligned) of the L bases held left-aligned in mer32
__device__ __forceinline__ u64 extract_seed(u64 mer32, const SeedParams& sp)
{
    u64 win = mer32 >> (64 - 2 * sp.L);
    u64 f = 0;
#pragma unroll 4
    for (int r = 0; r < sp.nruns; ++r) f |= ((win >> sp.shift_in[r]) & sp.mask[r]) << sp.shift_out[r];
    return f;
}

// reverse complement of a right-aligned 2w-bit seed
__device__ __forceinline__ u64 revcomp_seed(u64 f, int w)
{
    u64 x = __brevll(~f);  // complement, reverse all bits
    x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);  // restore bit order inside each base
    return x >> (64 - 2 * w);
}

// Which rank of a sharded run owns a seed: a hash of the low words of forward ^ reverse-complement, which is the same for a mer and
// its reverse complement, so the decision needs neither the canonical form nor the bucket of the mer.  The low word of the forward
// seed is its last min(16, w) bases, the low word of the reverse complement comes from its first min(16, w) bases: a scan over
// consecutive positions gets both from two sliding windows (bkf_scatter1_kernel's sharded path) -- the non-owners' whole cost.
// Every enumeration path uses it, so ranks that end up on different paths (bucket overflow fallback) still agree on the partition.
__device__ __forceinline__ bool seed_owned_x(u32 x, u32 shard, u32 nshard)   // x = (u32)forward ^ (u32)reverse complement
{
    return __umulhi(x * 0x9E3779B1u, nshard) == shard;
}
__device__ __forceinline__ bool seed_owned(u64 f, u64 rc, u32 shard, u32 nshard)
{
    if (nshard <= 1) return true;
    return seed_owned_x((u32)f ^ (u32)rc, shard, nshard);
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif

}  // namespace mcu
