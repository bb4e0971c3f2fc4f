// Bucketed seed-match enumeration (sm_100a): the B200-first replacement of "sort both mer lists, then merge".
//
// MatchFinder::SearchRange (LM/MatchFinder.cpp:172-340) only needs EQUAL mers of the two genomes to meet; the
// total order std::sort produces (LM/MemorySML.cpp:54) is needed by MemorySML::Read, not by match finding.
// So instead of P = ceil((2w+2)/8) full LSD passes over (key, position) pairs (2(Kb+4) bytes/pair/pass), the
// seeds are PARTITIONED by the top T bits of the canonical mer (two scatter passes of d1 and d2 bits over
// 8-byte records) into ~2^T buckets of ~2k records, and every bucket is finished inside one CTA in shared
// memory (counting split on the next 8 bits, then warp-level match.any on the remaining key bits):
//
//   bk_hist1     count canonical mers per level-1 bucket          reads packed genomes only (25 MB @100 Mbp)
//   bk_scan1     offsets + this rank's bucket range (sharding by exact counts, identical on every rank)
//   bk_scatter1  recompute the mer, build the 8-byte record, scatter by level-1 digit      8 B/seed written
//   bk_hist2     per level-1 bucket: count level-2 digits                                  8 B/seed read
//   bk_scan2     exclusive scan of the (bucket, digit) table
//   bk_scatter2  scatter by level-2 digit                                                  8 + 8 B/seed
//   bk_group     one CTA per final bucket: unique-in-both keys -> uniq bitmap + pair list  8 B/seed + 8 B/pair
//
// 40 B/seed instead of 12 + 5*24 + 12 = 144 B/seed (w = 19): the path is HBM-bound, so this is the lever.
// record = keyrem(2w - d1 bits) | position(pbits) | strand | genome; the top d1 key bits are implied by the
// bucket.  Scatter passes are unordered (atomic cursors, no look-back chain): grouping needs no stability.
// Buckets that do not fit shared memory (low-complexity sequence) are spilled to the radix-sort + join path
// (radix.cuh / join_kernel), which handles any size; plans that do not fit 64-bit records use that path too.
// Output is identical to join_kernel's: uniq bitmap, pairs (forward from the front / reverse from the back),
// counters[0]/[6] pair counts, counters[1] MER_REPEAT_LIMIT flag.
#include "anchor.cuh"

#include <math.h>

namespace mcu {

constexpr int BK_THREADS = 256;
constexpr int BK_IPT = 16;
constexpr int BK_TILE = BK_THREADS * BK_IPT;   // records per scatter tile
constexpr int BK_SUPER = 16 * BK_TILE;         // records per histogram block iteration
constexpr int BK_CAP = 2048;                   // largest final bucket finished in shared memory
constexpr int BK_HALF = BK_CAP / 2;            // fixed-capacity layout: records of genome 0 in the first half of a final bucket, genome 1 in the second
constexpr int BK_MAXB = 2048;                  // max bins per level
constexpr int BK_D3 = 10;

struct BkPlan {
    int kbits, d1, d2, d3, pbits, rem1;
    int aux;     // 4 neighbour-base bits per record (solid seeds only, when they fit), else 0
    int kshift;  // pbits + 2 + aux: where the key remainder starts
    int w;       // seed weight
    u32 B1, B2;
    u64 npos0, npos1, ntot;
    u64 npad0, nidx;  // genome-0 positions padded to a multiple of BK_TILE: a thread's 16 consecutive positions share 3 packed words and a scatter tile holds one genome
    // fixed-capacity layout (histogram-free path): this rank owns the level-1 bins [b_lo, b_hi); level-1 bucket b keeps the records of
    // genome g in its own segment of cap1g[g] records (segment g of bucket b starts at b * (cap1g[0] + cap1g[1]) + g * cap1g[0]);
    // a final bucket has BK_HALF records of room per genome.  cap1 = max(cap1g)
    u32 b_lo, b_hi, cap1, cap1g[2];
    u32 shard, nshard;  // seed ownership in a sharded run (seed_owned, common.cuh)
};

struct BkMeta {  // written by bk_scan1_kernel
    u32 b_lo, b_hi;
    u64 nrec;    // records of this rank
    u64 nfinal;  // (b_hi - b_lo) * B2
};

// Canonical mers are far from uniform (min(f, rc) piles up at small values, base composition skews the
// leading bases), and only EQUALITY of mers matters here, so buckets are cut on a bijective mix of the
// canonical mer (xorshift, odd multiply on 2w bits): bucket sizes become Poisson-tight whatever the
// genome looks like; identical mers still share a bucket.
__device__ __forceinline__ u64 bk_mix(u64 x, int kbits)
{
    const u64 mask = kbits >= 64 ? ~0ull : ((1ull << kbits) - 1);
    x ^= x >> (kbits / 2 + 1);                      // high bits reach the low half
    return (x * 0x9E3779B97F4A7C15ull) & mask;      // every output bit depends on all lower input bits: the top (bucket) bits on all of them
}

// false when another rank owns the seed
__device__ __forceinline__ bool seed_canon32(u64 mer32, const SeedParams& sp, u32 shard, u32 nshard, u64& key, u32& strand)
{
    const u64 f = extract_seed(mer32, sp);
    const u64 rc = revcomp_seed(f, sp.w);
    if (!seed_owned(f, rc, shard, nshard)) return false;
    strand = rc < f;  // GetDnaSeedMer: forward wins ties
    key = bk_mix(strand ? rc : f, 2 * sp.w);
    return true;
}

// 16 consecutive positions starting at a multiple of 16 read the same three packed words
struct BkWindow {
    u64 hi;
    u32 lo;
    __device__ __forceinline__ void load(const u32* __restrict__ packed, u64 pos16)
    {
        const u64 k = pos16 >> 4;
        hi = ((u64)__ldg(packed + k) << 32) | __ldg(packed + k + 1);
        lo = __ldg(packed + k + 2);
    }
    __device__ __forceinline__ u64 mer(int j) const { return j ? ((hi << (2 * j)) | ((u64)lo >> (32 - 2 * j))) : hi; }
    __device__ __forceinline__ u32 base(int j) const { return j < 32 ? (u32)(hi >> (62 - 2 * j)) & 3u : (lo >> (94 - 2 * j)) & 3u; }  // j < 48
};

// ---- level 1 ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BK_THREADS) bk_hist1_kernel(const u32* __restrict__ g0, const u32* __restrict__ g1, BkPlan pl, SeedParams sp,
                                                             unsigned long long* __restrict__ count1)
{
    extern __shared__ u32 sh[];
    for (u32 i = threadIdx.x; i < pl.B1; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 ngroups = pl.nidx >> 4;
    for (u64 grp = (u64)blockIdx.x * blockDim.x + threadIdx.x; grp < ngroups; grp += (u64)gridDim.x * blockDim.x) {
        const u64 idx = grp << 4;
        const bool g = idx >= pl.npad0;
        const u64 pos = g ? idx - pl.npad0 : idx, npos = g ? pl.npos1 : pl.npos0;
        BkWindow w;
        w.load(g ? g1 : g0, pos);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (pos + j < npos) {
                u64 canon;
                u32 strand;
                if (seed_canon32(w.mer(j), sp, pl.shard, pl.nshard, canon, strand)) atomicAdd(&sh[(u32)(canon >> pl.rem1)], 1u);
            }
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < pl.B1; i += blockDim.x)
        if (sh[i]) atomicAdd(&count1[i], (unsigned long long)sh[i]);
}

// offsets of the level-1 buckets (counts of this rank's seeds)
__global__ void bk_scan1_kernel(const unsigned long long* __restrict__ count1, BkPlan pl, int shard, int nshard, u64* __restrict__ off1,
                                unsigned long long* __restrict__ cursor1, BkMeta* __restrict__ meta)
{
    if (threadIdx.x || blockIdx.x) return;
    u64 total = 0;
    for (u32 b = 0; b < pl.B1; ++b) total += count1[b];
    const u32 edge_lo = 0, edge_hi = pl.B1;  // ownership is decided per seed (seed_owned): every bucket holds this rank's share
    u64 run = 0;
    for (u32 b = 0; b < pl.B1; ++b) {
        off1[b] = run;
        cursor1[b] = run;
        if (b >= edge_lo && b < edge_hi) run += count1[b];
    }
    off1[pl.B1] = run;
    meta->b_lo = edge_lo;
    meta->b_hi = edge_hi;
    meta->nrec = run;
    meta->nfinal = (u64)(edge_hi - edge_lo) * pl.B2;
}

// Shared scatter machinery: every thread holds up to IPT (record, bin) items of a tile; ranks come from
// shared-memory atomics, one global reservation per (tile, bin), then the tile is staged bin-major in shared
// memory and streamed out so that consecutive threads write consecutive addresses inside a bin run.
struct BkScatterSmem {
    u64* stage;        // [BK_TILE]
    unsigned short* sbin;  // [BK_TILE]
    u32* cnt;          // [bins]
    u32* sofs;         // [bins]
    u64* gbase;        // [bins]
    u32* warp_tot;     // [8]
};

__device__ __forceinline__ BkScatterSmem bk_carve(unsigned char* raw, u32 bins)
{
    BkScatterSmem s;
    s.stage = (u64*)raw;
    s.gbase = (u64*)(raw + (size_t)BK_TILE * 8);
    s.cnt = (u32*)(raw + (size_t)BK_TILE * 8 + (size_t)bins * 8);
    s.sofs = s.cnt + bins;
    s.warp_tot = s.sofs + bins;
    s.sbin = (unsigned short*)(s.warp_tot + 8);
    return s;
}

static size_t bk_scatter_smem_bytes(u32 bins) { return (size_t)BK_TILE * 8 + (size_t)bins * 16 + 32 + (size_t)BK_TILE * 2; }

// After cnt[] holds the tile's bin counts: reserve global space, compute the bin-major staging offsets.
__device__ __forceinline__ void bk_reserve(const BkScatterSmem& s, u32 bins, unsigned long long* __restrict__ cursor)
{
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // exclusive scan of cnt over `bins` entries (bins <= 2048 = 8 per thread)
    const u32 per = (bins + BK_THREADS - 1) / BK_THREADS;
    u32 local = 0;
    for (u32 j = 0; j < per; ++j) {
        const u32 b = tid * per + j;
        if (b < bins) local += s.cnt[b];
    }
    u32 incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s.warp_tot[warp] = incl;
    __syncthreads();
    u32 add = 0;
    for (u32 w = 0; w < warp; ++w) add += s.warp_tot[w];
    u32 run = incl - local + add;
    for (u32 j = 0; j < per; ++j) {
        const u32 b = tid * per + j;
        if (b < bins) {
            const u32 c = s.cnt[b];
            s.sofs[b] = run;
            run += c;
            if (c) s.gbase[b] = atomicAdd(&cursor[b], (unsigned long long)c);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void bk_flush(const BkScatterSmem& s, u32 ntile, u64* __restrict__ out)
{
    for (u32 j = threadIdx.x; j < ntile; j += BK_THREADS) {
        const u32 b = s.sbin[j];
        out[s.gbase[b] + (j - s.sofs[b])] = s.stage[j];
    }
}

__global__ void __launch_bounds__(BK_THREADS, 3) bk_scatter1_kernel(const u32* __restrict__ g0, const u32* __restrict__ g1, BkPlan pl, SeedParams sp,
                                                                   const BkMeta* __restrict__ meta, unsigned long long* __restrict__ cursor1,
                                                                   u64* __restrict__ recs)
{
    extern __shared__ __align__(16) unsigned char raw[];
    const BkScatterSmem s = bk_carve(raw, pl.B1);
    const u32 tid = threadIdx.x;
    for (u32 i = tid; i < pl.B1; i += BK_THREADS) s.cnt[i] = 0;
    __syncthreads();
    const u64 idx0 = (u64)blockIdx.x * BK_TILE + (u64)tid * BK_IPT;  // 16 consecutive positions per thread
    u64 rec[BK_IPT];
    u32 br[BK_IPT];  // bin << 16 | rank   (rank < BK_TILE = 4096)
    {
        const u32 g = idx0 >= pl.npad0;
        const u64 pos = g ? idx0 - pl.npad0 : idx0, npos = g ? pl.npos1 : pl.npos0;
        BkWindow w;
        u32 prevbase = 0;
        if (idx0 < pl.nidx) {
            w.load(g ? g1 : g0, pos);
            if (pl.aux && pos > 0) prevbase = base_at(g ? g1 : g0, (i64)pos - 1);
        }
#pragma unroll
        for (int it = 0; it < BK_IPT; ++it) {
            br[it] = 0xffffffffu;
            if (idx0 < pl.nidx && pos + it < npos) {
                u64 canon;
                u32 strand;
                if (seed_canon32(w.mer(it), sp, pl.shard, pl.nshard, canon, strand)) {
                    const u32 b = (u32)(canon >> pl.rem1);
                    const u64 keyrem = canon & ((1ull << pl.rem1) - 1);
                    u64 aux = 0;
                    if (pl.aux) {  // base before the seed | base after the seed << 2 (solid seeds: what a shift by one position adds)
                        const u32 prev = it ? w.base(it - 1) : prevbase;
                        aux = prev | (w.base(it + pl.w) << 2);
                    }
                    rec[it] = (keyrem << pl.kshift) | (aux << (pl.pbits + 2)) | ((pos + it) << 2) | ((u64)strand << 1) | (u64)g;
                    br[it] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
                }
            }
        }
    }
    __syncthreads();
    bk_reserve(s, pl.B1, cursor1);
    u32 ntile = 0;
#pragma unroll
    for (int it = 0; it < BK_IPT; ++it) {
        if (br[it] != 0xffffffffu) {
            const u32 b = br[it] >> 16, slot = s.sofs[b] + (br[it] & 0xffffu);
            s.stage[slot] = rec[it];
            s.sbin[slot] = (unsigned short)b;
        }
    }
    // tile population = sofs of the last bin + its count
    ntile = s.sofs[pl.B1 - 1] + s.cnt[pl.B1 - 1];
    __syncthreads();
    bk_flush(s, ntile, recs);
}

// ---- level 2 ----------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 bk_segment_of(const u64* __restrict__ off1, u32 B1, u64 i)
{
    u32 lo = 0, hi = B1;  // off1[lo] <= i < off1[hi]
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (off1[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// level-1 bucket that holds the first record of every scatter tile (so the scatter does not search)
__global__ void bk_tileseg_kernel(const u64* __restrict__ off1, BkPlan pl, const BkMeta* __restrict__ meta, u32* __restrict__ tile_seg, u32 ntiles)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const u64 i = (u64)t * BK_TILE;
    tile_seg[t] = i < meta->nrec ? bk_segment_of(off1, pl.B1, i) : 0u;
}

__global__ void __launch_bounds__(BK_THREADS) bk_hist2_kernel(const u64* __restrict__ recs, BkPlan pl, const BkMeta* __restrict__ meta,
                                                             const u64* __restrict__ off1, unsigned long long* __restrict__ count2)
{
    extern __shared__ u32 sh[];
    const u64 nrec = meta->nrec;
    const u32 b_lo = meta->b_lo;
    const int shift = pl.kshift + pl.rem1 - pl.d2;
    for (u64 base = (u64)blockIdx.x * BK_SUPER; base < nrec; base += (u64)gridDim.x * BK_SUPER) {
        const u64 end = min(base + (u64)BK_SUPER, nrec);
        u64 i0 = base;
        while (i0 < end) {
            const u32 seg = bk_segment_of(off1, pl.B1, i0);
            const u64 i1 = min(end, off1[seg + 1]);
            for (u32 i = threadIdx.x; i < pl.B2; i += blockDim.x) sh[i] = 0;
            __syncthreads();
            for (u64 i = i0 + threadIdx.x; i < i1; i += blockDim.x) atomicAdd(&sh[(u32)(recs[i] >> shift) & (pl.B2 - 1)], 1u);
            __syncthreads();
            unsigned long long* dst = count2 + (u64)(seg - b_lo) * pl.B2;
            for (u32 i = threadIdx.x; i < pl.B2; i += blockDim.x)
                if (sh[i]) atomicAdd(&dst[i], (unsigned long long)sh[i]);
            __syncthreads();
            i0 = i1;
        }
    }
}

// exclusive scan of count2[0 .. nfinal) -> off2 (nfinal + 1 entries) and the scatter cursors; one block
__global__ void __launch_bounds__(1024) bk_scan2_kernel(const unsigned long long* __restrict__ count2, const BkMeta* __restrict__ meta,
                                                       u64* __restrict__ off2, unsigned long long* __restrict__ cursor2)
{
    __shared__ u64 s_tot[1024];
    const u64 n = meta->nfinal;
    const u64 per = (n + 1023) / 1024;
    const u32 t = threadIdx.x;
    const u64 lo = min(n, (u64)t * per), hi = min(n, lo + per);
    u64 local = 0;
    for (u64 i = lo; i < hi; ++i) local += count2[i];
    s_tot[t] = local;
    __syncthreads();
    if (t == 0) {
        u64 run = 0;
        for (u32 i = 0; i < 1024; ++i) { const u64 v = s_tot[i]; s_tot[i] = run; run += v; }
        off2[n] = run;
    }
    __syncthreads();
    u64 run = s_tot[t];
    for (u64 i = lo; i < hi; ++i) {
        off2[i] = run;
        cursor2[i] = run;
        run += count2[i];
    }
}

__global__ void __launch_bounds__(BK_THREADS, 3) bk_scatter2_kernel(const u64* __restrict__ src, BkPlan pl, const BkMeta* __restrict__ meta,
                                                                   const u64* __restrict__ off1, const u32* __restrict__ tile_seg,
                                                                   unsigned long long* __restrict__ cursor2, u64* __restrict__ dst)
{
    extern __shared__ __align__(16) unsigned char raw[];
    const BkScatterSmem s = bk_carve(raw, pl.B2);
    const u32 tid = threadIdx.x;
    const u64 nrec = meta->nrec;
    const u32 b_lo = meta->b_lo;
    const int shift = pl.kshift + pl.rem1 - pl.d2;
    const u64 base = (u64)blockIdx.x * BK_TILE;
    if (base >= nrec) return;
    const u64 end = min(base + (u64)BK_TILE, nrec);
    u64 i0 = base;
    u32 seg = tile_seg[blockIdx.x];
    while (i0 < end) {  // one iteration unless the tile straddles level-1 buckets
        while (off1[seg + 1] <= i0) ++seg;
        const u64 i1 = min(end, off1[seg + 1]);
        for (u32 i = tid; i < pl.B2; i += BK_THREADS) s.cnt[i] = 0;
        __syncthreads();
        u64 rec[BK_IPT];
        u32 br[BK_IPT];
#pragma unroll
        for (int it = 0; it < BK_IPT; ++it) {  // all loads in flight before the first shared-memory atomic
            const u64 i = i0 + (u64)it * BK_THREADS + tid;
            rec[it] = i < i1 ? __ldcs(src + i) : 0;
        }
#pragma unroll
        for (int it = 0; it < BK_IPT; ++it) {
            const u64 i = i0 + (u64)it * BK_THREADS + tid;
            br[it] = 0xffffffffu;
            if (i < i1) {
                const u32 b = (u32)(rec[it] >> shift) & (pl.B2 - 1);
                br[it] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
            }
        }
        __syncthreads();
        bk_reserve(s, pl.B2, cursor2 + (u64)(seg - b_lo) * pl.B2);
#pragma unroll
        for (int it = 0; it < BK_IPT; ++it) {
            if (br[it] != 0xffffffffu) {
                const u32 b = br[it] >> 16, slot = s.sofs[b] + (br[it] & 0xffffu);
                s.stage[slot] = rec[it];
                s.sbin[slot] = (unsigned short)b;
            }
        }
        __syncthreads();
        bk_flush(s, (u32)(i1 - i0), dst);
        __syncthreads();
        i0 = i1;
    }
}


// ---- histogram-free variant: fixed-capacity buckets --------------------------------------------------------------
// The mixed mers spread the seeds uniformly, so bucket sizes are known in advance up to Poisson noise: level-1 bucket b of
// this rank owns recs1[(b - b_lo) * cap1 ..) with cap1 = mean + 8 sigma, final bucket f owns recs2[f * BK_CAP ..).  That
// removes both counting passes (bk_hist1: a second evaluation of every seed; bk_hist2: a second read of every record) and
// both scans, and the rank's share is a fixed range of level-1 bins.  Records that find their bucket full (repeats: many
// copies of one mer) are written to the overflow arrays in the radix-sort path's (key, position) format, their final
// bucket is marked dirty, and bk_group sends dirty / overfull buckets to that path entirely, so all copies of a mer are
// always examined together.  If the overflow arrays themselves fill up, the caller reruns the exact (counting) path.
struct BkOvf {
    u64* keys;                  // mixed mer << 2 | genome << 1 | strand
    u32* vals;                  // position
    u64 cap;
    unsigned long long* cursor; // records appended (may run past cap: detected by the host)
    u32* dirty;                 // 1 bit per final bucket of this rank
};

__device__ __noinline__ void bk_overflow(u64* keys, u32* vals, u64 cap, unsigned long long* cursor, u32* dirty, u64 rec, u32 b1, u32 b1_rel, int rem1,
                                         int kshift, int pbits, int d2)
{
    const u64 mixed = ((u64)b1 << rem1) | (rec >> kshift);
    const u64 slot = atomicAdd(cursor, 1ull);
    if (slot < cap) {
        keys[slot] = (mixed << 2) | ((rec & 1) << 1) | ((rec >> 1) & 1);
        vals[slot] = (u32)((rec >> 2) & ((1ull << pbits) - 1));
    }
    const u64 f = ((u64)b1_rel << d2) + ((rec >> (kshift + rem1 - d2)) & ((1u << d2) - 1));
    atomicOr(&dirty[f >> 5], 1u << (f & 31));
}

// like bk_reserve, for buckets of fixed capacity: gbase[b] = absolute index of the tile's first record of bin b,
// cnt[b] becomes the number of the tile's records of bin b that still fit.  Returns the tile population.
template <typename BaseFn>
__device__ __forceinline__ u32 bkf_reserve(const BkScatterSmem& s, u32 bins, unsigned long long* __restrict__ cursor, u32 cstride, u64 cap, BaseFn bucket_base)
{
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 per = (bins + BK_THREADS - 1) / BK_THREADS;
    u32 local = 0;
    for (u32 j = 0; j < per; ++j) {
        const u32 b = tid * per + j;
        if (b < bins) local += s.cnt[b];
    }
    u32 incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s.warp_tot[warp] = incl;
    __syncthreads();
    u32 add = 0, total = 0;
#pragma unroll
    for (u32 w = 0; w < BK_THREADS / 32; ++w) {
        const u32 t = s.warp_tot[w];
        if (w < warp) add += t;
        total += t;
    }
    u32 run = incl - local + add;
    for (u32 j = 0; j < per; ++j) {
        const u32 b = tid * per + j;
        if (b < bins) {
            const u32 c = s.cnt[b];
            s.sofs[b] = run;
            run += c;
            if (c) {
                const u64 rel = atomicAdd(&cursor[(size_t)b * cstride], (unsigned long long)c);
                s.gbase[b] = bucket_base(b) + rel;
                s.cnt[b] = rel >= cap ? 0u : (u32)min((u64)c, cap - rel);
            }
        }
    }
    __syncthreads();
    return total;
}

template <bool SOLID, int MINB>
__global__ void __launch_bounds__(BK_THREADS, MINB) bkf_scatter1_kernel(const u32* __restrict__ g0, const u32* __restrict__ g1, BkPlan pl, SeedParams sp,
                                                                    unsigned long long* __restrict__ cursor1, u64* __restrict__ recs, BkOvf ovf, u32 tile_base)
{
    extern __shared__ __align__(16) unsigned char raw[];
    const BkScatterSmem s = bk_carve(raw, pl.B1);
    const u32 tid = threadIdx.x;
    const u32 b_lo = pl.b_lo;
    for (u32 i = tid; i < pl.B1; i += BK_THREADS) s.cnt[i] = 0;
    __syncthreads();
    const u64 idx0 = ((u64)tile_base + blockIdx.x) * BK_TILE + (u64)tid * BK_IPT;  // 16 consecutive positions per thread
    u64 rec[BK_IPT];
    u32 br[BK_IPT];  // bin << 16 | rank   (rank < BK_TILE = 4096)
#pragma unroll
    for (int it = 0; it < BK_IPT; ++it) br[it] = 0xffffffffu;
    // sharded solid-seed runs: ownership mask of this thread's 16 positions, consumed after the range test below so that
    // EVERY lane of the warp (also those beyond nidx, whose mask stays 0) takes part in the warp votes
    u32 own_mask = 0, own_nx = 0, own_prev = 0, own_g = 0;
    u64 own_pos = 0;
    BkWindow own_w;
    own_w.hi = 0; own_w.lo = 0;
    const bool compact = SOLID && pl.nshard > 1;
    if (idx0 < pl.nidx) {
        const u32 g = idx0 >= pl.npad0;
        const u64 pos = g ? idx0 - pl.npad0 : idx0, npos = g ? pl.npos1 : pl.npos0;
        const u32* __restrict__ gp = g ? g1 : g0;
        BkWindow w;
        w.load(gp, pos);
        const int kbits = 2 * sp.w;
        const u64 kmask = (1ull << kbits) - 1;   // kbits <= 62
        const int mshift = kbits / 2 + 1;
        if (compact) {
            // ownership of the 16 seeds from two sliding windows (seed_owned_x, common.cuh): the low word of the forward seed is its
            // last m = min(16, w) bases, the low word of its reverse complement the complemented reversal of its first m bases.
            const u32 hi_h = (u32)(w.hi >> 32), hi_l = (u32)w.hi;
            const u32 nx = kbits < 32 ? __funnelshift_l(hi_l, hi_h, kbits) : __funnelshift_l(w.lo, hi_l, kbits - 32);
            const int m = sp.w < 16 ? sp.w : 16, skip = sp.w - m;                       // skip <= 15
            const u64 last = skip ? (w.hi << (2 * skip)) | ((u64)w.lo >> (32 - 2 * skip)) : w.hi;   // bases [skip, skip + 32) of the window
            const u64 tail = last >> (34 - 2 * m);                                      // seed `it` ends at bit 30 - 2 it of `tail`
            u64 head = __brevll(~w.hi);                                                 // complemented reversal of bases [0, 32)
            head = ((head & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((head & 0x5555555555555555ull) << 1);
            const u32 xmask = m == 16 ? 0xffffffffu : (1u << (2 * m)) - 1u;
            const u32 nvalid = pos >= npos ? 0u : (npos - pos >= 16 ? 16u : (u32)(npos - pos));
#pragma unroll
            for (int it = 0; it < BK_IPT; ++it) {
                // first m bases of seed `it` = bases [it, it + m): `head` has the complement of base k at bits 2k+1, 2k
                const u32 x = ((u32)(tail >> (30 - 2 * it)) ^ (u32)(head >> (2 * it))) & xmask;
                if (seed_owned_x(x, pl.shard, pl.nshard)) own_mask |= 1u << it;
            }
            own_mask &= (1u << nvalid) - 1u;
            own_nx = nx; own_g = g; own_pos = pos; own_w = w;
            if (pl.aux && pos > 0) own_prev = base_at(gp, (i64)pos - 1);
        } else if (SOLID) {
            // rolling evaluation: a shift by one position drops one base and adds one (forward: at the low end; reverse
            // complement: the complement at the high end).  nx = the 16 bases [w, w+16) of the window.
            const u32 hi_h = (u32)(w.hi >> 32), hi_l = (u32)w.hi;
            const u32 nx = kbits < 32 ? __funnelshift_l(hi_l, hi_h, kbits) : __funnelshift_l(w.lo, hi_l, kbits - 32);
            u32 prevbase = 0;
            if (pl.aux && pos > 0) prevbase = base_at(gp, (i64)pos - 1);
            u64 f = w.hi >> (64 - kbits);
            u64 rc = revcomp_seed(f, sp.w);
#pragma unroll
            for (int it = 0; it < BK_IPT; ++it) {
                const u32 nb = (nx >> (30 - 2 * it)) & 3u;  // base w + it: enters the mer at the next position, follows it at this one
                if (pos + it < npos && seed_owned(f, rc, pl.shard, pl.nshard)) {
                    const u32 strand = rc < f;  // GetDnaSeedMer: forward wins ties
                    u64 x = strand ? rc : f;
                    x ^= x >> mshift;
                    const u64 canon = (x * 0x9E3779B97F4A7C15ull) & kmask;  // == bk_mix
                    const u32 b = (u32)(canon >> pl.rem1);
                    {
                        const u64 keyrem = canon & ((1ull << pl.rem1) - 1);
                        u64 aux = 0;
                        if (pl.aux) {
                            const u32 prev = it ? (u32)(w.hi >> (64 - 2 * it)) & 3u : prevbase;
                            aux = prev | (nb << 2);
                        }
                        rec[it] = (keyrem << pl.kshift) | (aux << (pl.pbits + 2)) | ((pos + it) << 2) | ((u64)strand << 1) | (u64)g;
                        br[it] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
                    }
                }
                f = ((f << 2) | nb) & kmask;
                rc = (rc >> 2) | ((u64)(3u - nb) << (kbits - 2));
            }
        } else {
#pragma unroll
            for (int it = 0; it < BK_IPT; ++it) {
                u64 canon;
                u32 strand;
                if (pos + it < npos && seed_canon32(w.mer(it), sp, pl.shard, pl.nshard, canon, strand)) {
                    const u32 b = (u32)(canon >> pl.rem1);
                    {
                        const u64 keyrem = canon & ((1ull << pl.rem1) - 1);
                        rec[it] = (keyrem << pl.kshift) | ((pos + it) << 2) | ((u64)strand << 1) | (u64)g;
                        br[it] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
                    }
                }
            }
        }
    }
    if (compact) {
        // as many rounds of record building as the busiest lane of the warp owns seeds (mean 16 / nshard)
        const int kbits = 2 * sp.w;
        const u64 kmask = (1ull << kbits) - 1;
        const int mshift = kbits / 2 + 1;
#pragma unroll
        for (int k = 0; k < BK_IPT; ++k) {
            if (!__any_sync(0xffffffffu, own_mask != 0)) break;   // uniform: all 32 lanes are here
            if (own_mask) {
                const int it = __ffs(own_mask) - 1;
                own_mask &= own_mask - 1;
                const u64 mer = (own_w.hi << (2 * it)) | ((u64)own_w.lo >> (32 - 2 * it));
                const u64 ff = mer >> (64 - kbits), rr = revcomp_seed(ff, sp.w);
                const u32 strand = rr < ff;
                u64 x = strand ? rr : ff;
                x ^= x >> mshift;
                const u64 canon = (x * 0x9E3779B97F4A7C15ull) & kmask;  // == bk_mix
                const u32 b = (u32)(canon >> pl.rem1);
                const u64 keyrem = canon & ((1ull << pl.rem1) - 1);
                u64 aux = 0;
                if (pl.aux) {
                    const u32 prev = it ? (u32)(own_w.hi >> (64 - 2 * it)) & 3u : own_prev;
                    aux = prev | (((own_nx >> (30 - 2 * it)) & 3u) << 2);
                }
                rec[k] = (keyrem << pl.kshift) | (aux << (pl.pbits + 2)) | ((own_pos + it) << 2) | ((u64)strand << 1) | (u64)own_g;
                br[k] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
            }
        }
    }
    __syncthreads();
    // the tile's genome (npad0 is a multiple of BK_TILE): its records go to that genome's segment of every level-1 bucket
    const u32 gt = ((u64)tile_base + blockIdx.x) * BK_TILE >= pl.npad0 ? 1u : 0u;
    const u64 cap1 = gt ? pl.cap1g[1] : pl.cap1g[0], segw = (u64)pl.cap1g[0] + pl.cap1g[1], goff = gt ? pl.cap1g[0] : 0;
    const u32 ntile = bkf_reserve(s, pl.B1, cursor1 + (size_t)gt * pl.B1, 1, cap1, [=](u32 b) { return (u64)(b - b_lo) * segw + goff; });
#pragma unroll
    for (int it = 0; it < BK_IPT; ++it) {
        if (br[it] != 0xffffffffu) {
            const u32 b = br[it] >> 16, slot = s.sofs[b] + (br[it] & 0xffffu);
            s.stage[slot] = rec[it];
            s.sbin[slot] = (unsigned short)b;
        }
    }
    __syncthreads();
    for (u32 j = tid; j < ntile; j += BK_THREADS) {
        const u32 b = s.sbin[j], k = j - s.sofs[b];
        if (k < s.cnt[b]) recs[s.gbase[b] + k] = s.stage[j];
        else bk_overflow(ovf.keys, ovf.vals, ovf.cap, ovf.cursor, ovf.dirty, s.stage[j], b, b - b_lo, pl.rem1, pl.kshift, pl.pbits, pl.d2);
    }
}

// Sharded runs with solid seeds (nshard > 1): a rank owns 1 / nshard of the seeds, so a 4096-position tile yields a fraction of a tile of
// records while the per-tile costs (clearing and scanning the bin counters, one reservation per non-empty bin, the barriers) stay.
// Here a CTA walks `sub` consecutive tiles of one genome (sub ~ 0.875 nshard: the expected yield stays eight standard deviations below
// one tile) and parks what it owns -- the ownership scan, the record format and the output are bkf_scatter1_kernel's -- in shared
// memory as (record, bin << 16 | rank in bin) in arrival order, reserves once, and writes every record to its reserved slot.  A
// record that finds the parking area full (never, statistically) takes the overflow route like a record whose bucket is full.
constexpr size_t bkf_multi_smem_bytes(u32 bins) { return (size_t)BK_TILE * 8 + (size_t)bins * 16 + 32 + (size_t)BK_TILE * 4 + 16; }

__global__ void __launch_bounds__(BK_THREADS, 3) bkf_scatter1_multi_kernel(const u32* __restrict__ g0, const u32* __restrict__ g1, BkPlan pl, SeedParams sp,
                                                                          unsigned long long* __restrict__ cursor1, u64* __restrict__ recs, BkOvf ovf,
                                                                          u32 tiles0, u32 tiles_total, u32 sub, u32 ctas0)
{
    extern __shared__ __align__(16) unsigned char raw[];
    const BkScatterSmem s = bk_carve(raw, pl.B1);
    u32* sbr = (u32*)(raw + (size_t)BK_TILE * 8 + (size_t)pl.B1 * 16 + 32);   // [BK_TILE] bin << 16 | rank (takes the place of sbin)
    u32* fill = sbr + BK_TILE;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 b_lo = pl.b_lo;
    for (u32 i = tid; i < pl.B1; i += BK_THREADS) s.cnt[i] = 0;
    if (tid == 0) *fill = 0;
    __syncthreads();
    const u32 gt = blockIdx.x >= ctas0 ? 1u : 0u;
    const u32 first = gt ? tiles0 + (blockIdx.x - ctas0) * sub : blockIdx.x * sub;
    const u32 last = min(first + sub, gt ? tiles_total : tiles0);
    const int kbits = 2 * sp.w;
    const u64 kmask = (1ull << kbits) - 1;   // kbits <= 62
    const int mshift = kbits / 2 + 1;
    const int m = sp.w < 16 ? sp.w : 16, skip = sp.w - m;                       // skip <= 15
    const u32 xmask = m == 16 ? 0xffffffffu : (1u << (2 * m)) - 1u;
    const u32* __restrict__ gp = gt ? g1 : g0;
    const u64 npos = gt ? pl.npos1 : pl.npos0;
    const u32 lt = lanemask_lt();
    for (u32 tile = first; tile < last; ++tile) {
        const u64 idx0 = (u64)tile * BK_TILE + (u64)tid * BK_IPT;
        u32 own_mask = 0, own_nx = 0, own_prev = 0;
        u64 own_pos = 0;
        BkWindow w;
        w.hi = 0; w.lo = 0;
        if (idx0 < pl.nidx) {
            const u64 pos = gt ? idx0 - pl.npad0 : idx0;
            w.load(gp, pos);
            // ownership of the 16 seeds from two sliding windows (seed_owned_x, common.cuh): see bkf_scatter1_kernel
            const u32 hi_h = (u32)(w.hi >> 32), hi_l = (u32)w.hi;
            own_nx = kbits < 32 ? __funnelshift_l(hi_l, hi_h, kbits) : __funnelshift_l(w.lo, hi_l, kbits - 32);
            const u64 lastb = skip ? (w.hi << (2 * skip)) | ((u64)w.lo >> (32 - 2 * skip)) : w.hi;
            const u64 tail = lastb >> (34 - 2 * m);
            u64 head = __brevll(~w.hi);
            head = ((head & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((head & 0x5555555555555555ull) << 1);
            const u32 nvalid = pos >= npos ? 0u : (npos - pos >= 16 ? 16u : (u32)(npos - pos));
#pragma unroll
            for (int it = 0; it < BK_IPT; ++it) {
                const u32 x = ((u32)(tail >> (30 - 2 * it)) ^ (u32)(head >> (2 * it))) & xmask;
                if (seed_owned_x(x, pl.shard, pl.nshard)) own_mask |= 1u << it;
            }
            own_mask &= (1u << nvalid) - 1u;
            own_pos = pos;
            if (pl.aux && pos > 0) own_prev = base_at(gp, (i64)pos - 1);
        }
        // as many rounds as the busiest lane of the warp owns seeds (mean 16 / nshard)
        for (int k = 0; k < BK_IPT; ++k) {
            const u32 have = __ballot_sync(0xffffffffu, own_mask != 0);
            if (!have) break;   // uniform
            u32 base = 0;
            if (lane == 0) base = atomicAdd(fill, (u32)__popc(have));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (own_mask) {
                const int it = __ffs(own_mask) - 1;
                own_mask &= own_mask - 1;
                const u64 mer = it ? (w.hi << (2 * it)) | ((u64)w.lo >> (32 - 2 * it)) : w.hi;
                const u64 ff = mer >> (64 - kbits), rr = revcomp_seed(ff, sp.w);
                const u32 strand = rr < ff;
                u64 x = strand ? rr : ff;
                x ^= x >> mshift;
                const u64 canon = (x * 0x9E3779B97F4A7C15ull) & kmask;  // == bk_mix
                const u32 b = (u32)(canon >> pl.rem1);
                const u64 keyrem = canon & ((1ull << pl.rem1) - 1);
                u64 aux = 0;
                if (pl.aux) {
                    const u32 prev = it ? (u32)(w.hi >> (64 - 2 * it)) & 3u : own_prev;
                    aux = prev | (((own_nx >> (30 - 2 * it)) & 3u) << 2);
                }
                const u64 rec = (keyrem << pl.kshift) | (aux << (pl.pbits + 2)) | ((own_pos + it) << 2) | ((u64)strand << 1) | (u64)gt;
                const u32 p = base + (u32)__popc(have & lt);
                if (p < (u32)BK_TILE) {
                    s.stage[p] = rec;
                    sbr[p] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
                } else
                    bk_overflow(ovf.keys, ovf.vals, ovf.cap, ovf.cursor, ovf.dirty, rec, b, b - b_lo, pl.rem1, pl.kshift, pl.pbits, pl.d2);
            }
        }
    }
    __syncthreads();
    const u64 cap1 = gt ? pl.cap1g[1] : pl.cap1g[0], segw = (u64)pl.cap1g[0] + pl.cap1g[1], goff = gt ? pl.cap1g[0] : 0;
    bkf_reserve(s, pl.B1, cursor1 + (size_t)gt * pl.B1, 1, cap1, [=](u32 b) { return (u64)(b - b_lo) * segw + goff; });
    const u32 ntile = min(*fill, (u32)BK_TILE);
    for (u32 j = tid; j < ntile; j += BK_THREADS) {
        const u32 b = sbr[j] >> 16, k = sbr[j] & 0xffffu;
        if (k < s.cnt[b]) recs[s.gbase[b] + k] = s.stage[j];
        else bk_overflow(ovf.keys, ovf.vals, ovf.cap, ovf.cursor, ovf.dirty, s.stage[j], b, b - b_lo, pl.rem1, pl.kshift, pl.pbits, pl.d2);
    }
}

// level-2 partition of one tile of level-1 bucket b_lo + blockIdx.y into its B2 final buckets
template <int MINB>
__global__ void __launch_bounds__(BK_THREADS, MINB) bkf_scatter2_kernel(const u64* __restrict__ src, BkPlan pl, const unsigned long long* __restrict__ cursor1,
                                                                    unsigned long long* __restrict__ cursor2, u64* __restrict__ dst, BkOvf ovf, u32 gsel)
{
    extern __shared__ __align__(16) unsigned char raw[];
    // level-1 bucket relative to b_lo, genome: gsel = 2 covers both genomes' segments with one grid, gsel = 0 / 1 one genome's (the
    // segments of genome 0 are complete while genome 1 is still being uploaded)
    const u32 seg = gsel == 2 ? blockIdx.y >> 1 : blockIdx.y, g = gsel == 2 ? blockIdx.y & 1 : gsel;
    const u64 cnt1 = min((u64)cursor1[(size_t)g * pl.B1 + pl.b_lo + seg], (u64)(g ? pl.cap1g[1] : pl.cap1g[0]));
    const u64 base = (u64)blockIdx.x * BK_TILE;
    if (base >= cnt1) return;
    const BkScatterSmem s = bk_carve(raw, pl.B2);
    const u32 tid = threadIdx.x;
    const int shift = pl.kshift + pl.rem1 - pl.d2;
    const u32 n = (u32)min((u64)BK_TILE, cnt1 - base);
    const u64* __restrict__ in = src + (u64)seg * ((u64)pl.cap1g[0] + pl.cap1g[1]) + (g ? pl.cap1g[0] : 0) + base;
    for (u32 i = tid; i < pl.B2; i += BK_THREADS) s.cnt[i] = 0;
    __syncthreads();
    u64 rec[BK_IPT];
    u32 br[BK_IPT];
#pragma unroll
    for (int it = 0; it < BK_IPT; ++it) {  // all loads in flight before the first shared-memory atomic
        const u32 i = it * BK_THREADS + tid;
        rec[it] = i < n ? __ldcs(in + i) : 0;
    }
#pragma unroll
    for (int it = 0; it < BK_IPT; ++it) {
        const u32 i = it * BK_THREADS + tid;
        br[it] = 0xffffffffu;
        if (i < n) {
            const u32 b = (u32)(rec[it] >> shift) & (pl.B2 - 1);
            br[it] = (b << 16) | atomicAdd(&s.cnt[b], 1u);
        }
    }
    __syncthreads();
    const u64 fbase = (u64)seg * pl.B2;
    bkf_reserve(s, pl.B2, cursor2 + fbase * 2 + g, 2, (u64)BK_HALF, [=](u32 b) { return (fbase + b) * (u64)BK_CAP + (u64)g * BK_HALF; });
#pragma unroll
    for (int it = 0; it < BK_IPT; ++it) {
        if (br[it] != 0xffffffffu) {
            const u32 b = br[it] >> 16, slot = s.sofs[b] + (br[it] & 0xffffu);
            s.stage[slot] = rec[it];
            s.sbin[slot] = (unsigned short)b;
        }
    }
    __syncthreads();
    for (u32 j = tid; j < n; j += BK_THREADS) {
        const u32 b = s.sbin[j], k = j - s.sofs[b];
        if (k < s.cnt[b]) dst[s.gbase[b] + k] = s.stage[j];
        else bk_overflow(ovf.keys, ovf.vals, ovf.cap, ovf.cursor, ovf.dirty, s.stage[j], pl.b_lo + seg, seg, pl.rem1, pl.kshift, pl.pbits, pl.d2);
    }
}

// records this rank holds = sum of its level-1 cursors (statistics only)
__global__ void bkf_total_kernel(const unsigned long long* __restrict__ cursor1, BkPlan pl, BkMeta* __restrict__ meta)
{
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    unsigned long long t = 0;
    for (u32 b = pl.b_lo + threadIdx.x; b < pl.b_hi; b += blockDim.x) t += cursor1[b] + cursor1[(size_t)pl.B1 + b];
    atomicAdd(&s_sum, t);
    __syncthreads();
    if (threadIdx.x == 0) {
        meta->b_lo = pl.b_lo;
        meta->b_hi = pl.b_hi;
        meta->nrec = s_sum;
        meta->nfinal = (u64)(pl.b_hi - pl.b_lo) * pl.B2;
    }
}

// ---- final buckets ------------------------------------------------------------------------------------------
struct BkGroupArgs {
    const u64* recs;
    const u64* off2;
    const BkMeta* meta;
    u32* uniq;
    u64* pairs;
    u64 pair_cap;
    unsigned long long* counters;  // [0] fwd pairs, [1] repeat flag, [6] rev pairs
    u64* cand;                     // candidate list (aux mode fills it directly with the pairs whose left neighbour is no hit)
    u64 cand_cap;
    u32* spill_list;               // final buckets larger than BK_CAP
    unsigned long long* spill;     // [0] buckets, [1] records, [2] convert cursor, [3] direct candidates
    // fixed-capacity layout (cnt2 != nullptr): bucket f = recs[f * BK_CAP ...], genome g's records in [g * BK_HALF, ...): cnt2[2 f + g]
    // arrived (those beyond BK_HALF went to the overflow arrays), dirty bit f = some of its records overflowed
    const unsigned long long* cnt2;
    const u32* dirty;
    u64 nfinal;
};

constexpr int BK_GIPT = BK_CAP / BK_THREADS;  // records per thread in bk_group
constexpr int BK_DIRECT = 64;                 // per-bucket buffer for pairs that are candidates for certain

__global__ void __launch_bounds__(BK_THREADS, 6) bk_group_kernel(BkGroupArgs a, BkPlan pl)
{
    __shared__ u64 stage[BK_CAP];
    __shared__ u64 outp[BK_CAP / 2];        // pairs of this bucket: forward from the front, reverse from the back
    __shared__ u32 sofs[(1 << BK_D3) + 1];  // bin counts first, exclusive offsets after the scan
    __shared__ u32 s_warp[8], s_np[2], s_nc[2];
    __shared__ unsigned long long s_base[4];
    __shared__ u64 candp[2][BK_DIRECT];     // pairs that are certainly candidates (aux mode)
    const u64 f = blockIdx.x;
    u64 beg;
    u32 nb, n0 = 0xffffffffu;   // split layout: records [n0, nb) start at beg + BK_HALF
    bool spill_it;
    if (a.cnt2) {
        const unsigned long long t0 = a.cnt2[2 * f], t1 = a.cnt2[2 * f + 1];
        beg = f * BK_CAP;
        n0 = (u32)min(t0, (unsigned long long)BK_HALF);
        nb = n0 + (u32)min(t1, (unsigned long long)BK_HALF);
        spill_it = t0 > BK_HALF || t1 > BK_HALF || ((a.dirty[f >> 5] >> (f & 31)) & 1u);
    } else {
        if (f >= a.meta->nfinal) return;
        beg = a.off2[f];
        const u64 cnt = a.off2[f + 1] - beg;
        nb = (u32)min(cnt, (u64)BK_CAP + 1);
        spill_it = cnt > BK_CAP;
    }
    if (nb == 0) return;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (spill_it) {
        if (tid == 0) {
            a.spill_list[atomicAdd(&a.spill[0], 1ull)] = (u32)f;
            atomicAdd(&a.spill[1], a.cnt2 ? (unsigned long long)nb : (unsigned long long)(a.off2[f + 1] - beg));
        }
        return;
    }
    const int kshift = pl.kshift;
    const u32 nsub = 1u << pl.d3;
    const int dshift = kshift + pl.rem1 - pl.d2 - pl.d3;
    const u64 samekey = 1ull << kshift;  // two records carry the same mer iff (x ^ y) < samekey
    const u32 posmask = (u32)((1ull << pl.pbits) - 1);
    for (u32 i = tid; i <= nsub; i += BK_THREADS) sofs[i] = 0;
    if (tid < 2) { s_np[tid] = 0; s_nc[tid] = 0; }
    __syncthreads();
    {   // counting split on the next d3 key bits: sub-groups of ~1 record, all records of one mer in one sub-group
        u64 rec[BK_GIPT];
        u32 br[BK_GIPT];
#pragma unroll
        for (int it = 0; it < BK_GIPT; ++it) {
            const u32 i = it * BK_THREADS + tid;
            rec[it] = i < nb ? __ldcs(a.recs + beg + (i < n0 ? i : i - n0 + BK_HALF)) : 0;
        }
#pragma unroll
        for (int it = 0; it < BK_GIPT; ++it) {
            const u32 i = it * BK_THREADS + tid;
            br[it] = 0xffffffffu;
            if (i < nb) {
                const u32 b = (u32)(rec[it] >> dshift) & (nsub - 1);
                br[it] = (b << 16) | atomicAdd(&sofs[b], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of the counts in place: thread t owns bins [t*per, (t+1)*per)
        const u32 per = (nsub + BK_THREADS - 1) / BK_THREADS;
        u32 local = 0;
        for (u32 j = 0; j < per; ++j) {
            const u32 b = tid * per + j;
            if (b < nsub) local += sofs[b];
        }
        u32 incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (u32)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        u32 run = incl - local;
        for (u32 w = 0; w < warp; ++w) run += s_warp[w];
        for (u32 j = 0; j < per; ++j) {
            const u32 b = tid * per + j;
            if (b < nsub) { const u32 c = sofs[b]; sofs[b] = run; run += c; }
        }
        if (tid == BK_THREADS - 1) sofs[nsub] = run;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < BK_GIPT; ++it)
            if (br[it] != 0xffffffffu) stage[sofs[br[it] >> 16] + (br[it] & 0xffffu)] = rec[it];
    }
    __syncthreads();
    // every genome-0 record scans its own sub-group: mer unique in both genomes -> pair
    u32 repeat = 0;
#pragma unroll 1
    for (u32 i = tid; i < nb; i += BK_THREADS) {
        const u64 r = stage[i];
        if (r & 1) continue;
        const u32 b = (u32)(r >> dshift) & (nsub - 1);
        const u32 s0 = sofs[b], s1 = sofs[b + 1];
        if (s1 - s0 < 2) continue;
        u32 c0 = 0, c1 = 0;
        u64 r1 = 0;
#pragma unroll 1
        for (u32 j = s0; j < s1; ++j) {
            const u64 q = stage[j];
            if ((q ^ r) < samekey) {
                if (q & 1) { ++c1; r1 = q; } else ++c0;
            }
        }
        if (c0 + c1 > 1000) repeat = 1;
        if (c0 == 1 && c1 == 1) {
            const u32 p0 = (u32)(r >> 2) & posmask, p1 = (u32)(r1 >> 2) & posmask;
            const u64 e = (u64)p0 | ((u64)p1 << 32);
            const u32 rev = (u32)((r ^ r1) >> 1) & 1u;
            bool direct = false;
            if (pl.aux) {
                // solid seed: the left neighbour on the diagonal is a hit iff the one new base agrees; if it does not,
                // this pair is certainly the leftmost seed of its hit run -> straight to the candidate list
                const u32 a0 = (u32)(r >> (pl.pbits + 2)) & 15u, a1 = (u32)(r1 >> (pl.pbits + 2)) & 15u;
                bool agree;
                if (!rev) agree = p0 > 0 && p1 > 0 && (a0 & 3u) == (a1 & 3u);
                else agree = p0 > 0 && (u64)p1 + 1 < pl.npos1 && (a0 & 3u) == 3u - (a1 >> 2);
                direct = !agree;
            }
            if (direct) {
                const u32 k = atomicAdd(&s_nc[rev], 1u);
                if (k < BK_DIRECT) candp[rev][k] = e;
                else if (!rev) a.cand[atomicAdd(&a.counters[2], 1ull)] = e;   // overflow of the small shared buffer: rare
                else a.cand[a.cand_cap - 1 - atomicAdd(&a.counters[7], 1ull)] = e;
                atomicOr(&a.uniq[p0 >> 5], 1u << (p0 & 31));
            } else if (rev) outp[BK_CAP / 2 - 1 - atomicAdd(&s_np[1], 1u)] = e;
            else outp[atomicAdd(&s_np[0], 1u)] = e;
        }
    }
    if (__any_sync(0xffffffffu, repeat) && lane == 0) atomicMax(&a.counters[1], 1ull);
    __syncthreads();
    const u32 nf = s_np[0], nr = s_np[1];
    const u32 ncf = min(s_nc[0], (u32)BK_DIRECT), ncr = min(s_nc[1], (u32)BK_DIRECT);
    if (tid == 0) {
        s_base[0] = nf ? atomicAdd(&a.counters[0], (unsigned long long)nf) : 0ull;
        s_base[1] = nr ? atomicAdd(&a.counters[6], (unsigned long long)nr) : 0ull;
        s_base[2] = ncf ? atomicAdd(&a.counters[2], (unsigned long long)ncf) : 0ull;
        s_base[3] = ncr ? atomicAdd(&a.counters[7], (unsigned long long)ncr) : 0ull;
        if (s_nc[0] + s_nc[1]) atomicAdd(&a.spill[3], (unsigned long long)(s_nc[0] + s_nc[1]));
    }
    __syncthreads();
    if (tid < ncf) a.cand[s_base[2] + tid] = candp[0][tid];
    if (tid < ncr) a.cand[a.cand_cap - 1 - (s_base[3] + tid)] = candp[1][tid];
    const u64 bf = s_base[0], br2 = s_base[1];
    for (u32 j = tid; j < nf; j += BK_THREADS) {
        const u64 e = outp[j];
        const u32 p0 = (u32)e;
        atomicOr(&a.uniq[p0 >> 5], 1u << (p0 & 31));
        a.pairs[bf + j] = e;
    }
    for (u32 j = tid; j < nr; j += BK_THREADS) {
        const u64 e = outp[BK_CAP / 2 - 1 - j];
        const u32 p0 = (u32)e;
        atomicOr(&a.uniq[p0 >> 5], 1u << (p0 & 31));
        a.pairs[a.pair_cap - 1 - (br2 + j)] = e;
    }
}

// ---- final buckets, TMA-fed version: a hash join per bucket -----------------------------------------------------------------
// bk_group_kernel above ranks every record with a returning shared-memory atomic (counting split), re-stages the 8-byte
// records and then lets every genome-0 record scan its sub-group: ncu (profiles/r01_ncu_bucket_enumeration_summary.txt) has it
// issue-bound at 340 thread-instructions per record with 120 M shared-memory bank conflicts.  This version:
//   * persistent CTAs; the two halves of a final bucket (genome 0's and genome 1's records: contiguous runs of <= 1024 8-byte
//     records at 8 KB-aligned addresses of the fixed-capacity layout) are brought into shared memory by bulk asynchronous copies
//     (cp.async.bulk + mbarrier complete_tx), double buffered: bucket k+1 is in flight while bucket k is joined; records are
//     never moved again;
//   * build: every genome-1 record claims a slot of an open-addressing table (32-bit words: key bits << 4 | dup0 dup1 seen0 seen1,
//     addressed by the low key bits) with ONE atomicCAS -- all of a thread's records in flight together -- and leaves its index
//     there; a record that finds its key already present marks it dup1;
//   * probe: every genome-0 record walks the table with plain loads; on a hit it ORs seen0 in (dup0 if it was set already);
//   * a genome-0 record whose slot reads seen0 | seen1 and nothing else is a seed pair; pairs are classified in registers and
//     go straight to the global lists (no staging): a warp prefix over packed per-thread counts and one reservation per bucket
//     and list give every thread its output ranges.
// Buckets with >= 999 duplicate records (a mer with more than MER_REPEAT_LIMIT copies may be inside, which only the sorted path
// counts exactly: with c1 >= 1 copies in genome 1 the duplicates number c0 + c1 - 2, and more than 1000 copies in genome 0 alone
// overflow its half) go to the radix-sort + join path like dirty / overfull buckets do.  Output is identical to
// bk_group_kernel's (pair order inside the lists is unspecified in both).
constexpr int G3_IPT = BK_HALF / BK_THREADS;   // records per thread and genome
template <int LOGS, int NBUF>
struct G3Smem {
    u64 raw[NBUF][BK_CAP];                // TMA destinations: [buffer][genome 0: 0.., genome 1: BK_HALF..]
    u32 tab[1 << LOGS];                   // key << 4 | flags; 0 = empty
    unsigned short idx1[1 << LOGS];       // index of the genome-1 record that claimed the slot
    unsigned long long bar[2];            // mbarriers of the two buffers
    unsigned long long base[4];           // global reservations: forward pairs, reverse pairs, forward / reverse direct candidates
    u32 wtot[BK_THREADS / 32][2];         // per-warp totals: [0] forward | reverse << 16, [1] forward direct | reverse direct << 16
    u32 losers;
};
constexpr int G3_MAX_KEY_BITS = 28;
constexpr u32 G3_SEEN1 = 1u, G3_SEEN0 = 2u, G3_DUP1 = 4u, G3_DUP0 = 8u;

__device__ __forceinline__ u32 g3_smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g3_mbar_init(unsigned long long* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g3_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void g3_mbar_expect_tx(unsigned long long* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g3_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g3_bulk_load(void* dst, const void* src, u32 bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(g3_smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(g3_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void g3_mbar_wait(unsigned long long* bar, u32 parity)
{
    u32 done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(g3_smem_addr(bar)), "r"(parity)
            : "memory");
    }
}

// seed pair (p0 | p1 << 32) of two records with the same mer, and where it goes: 0 forward pair, 1 reverse pair,
// 2 / 3 forward / reverse pair that is certainly a candidate (solid seeds: the base left of the seed differs on the diagonal,
// or there is no such base).  Branch-free: every lane of a warp classifies a pair of its own.
__device__ __forceinline__ u32 g3_classify(u64 r, u64 r1, bool aux, int auxshift, u32 posmask, u32 last1, u64& e)
{
    const u32 p0 = (u32)(r >> 2) & posmask, p1 = (u32)(r1 >> 2) & posmask;
    e = (u64)p0 | ((u64)p1 << 32);
    const u32 rev = (u32)((r ^ r1) >> 1) & 1u;
    u32 direct = 0;
    if (aux) {   // uniform
        const u32 a0 = (u32)(r >> auxshift), a1 = (u32)(r1 >> auxshift);      // bits 0-1: base before the seed, bits 2-3: base after it
        const u32 want = rev ? 3u - ((a1 >> 2) & 3u) : a1 & 3u;
        const bool inside = p0 > 0 && (rev ? p1 < last1 : p1 > 0);            // the left neighbour on the diagonal exists
        direct = (inside && (a0 & 3u) == want) ? 0u : 2u;
    }
    return rev | direct;
}

// NBUF = 2: the next bucket is copied while this one is joined; NBUF = 1: the next bucket's copy starts when this one's records have
// been read for the last time (it overlaps the reservation, the table clean-up and the previous bucket's write-out only, but the
// smaller footprint lets more CTAs share an SM)
template <int LOGS, int NBUF, int CTAS>
__global__ void __launch_bounds__(BK_THREADS, CTAS) bk_group3_kernel(BkGroupArgs a, BkPlan pl)
{
    extern __shared__ __align__(128) unsigned char g3_raw[];
    typedef G3Smem<LOGS, NBUF> Smem;
    Smem& sm = *reinterpret_cast<Smem*>(g3_raw);
    constexpr u32 SMASK = (1u << LOGS) - 1u;
    constexpr u32 NONE = 0xffffffffu;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 nfinal = a.nfinal, stride = gridDim.x;
    const int kshift = pl.kshift;
    const u32 posmask = (u32)((1ull << pl.pbits) - 1);
    const int kb = pl.rem1 - pl.d2;   // <= G3_MAX_KEY_BITS (host)
    const u32 kmask = (1u << kb) - 1u;
    // slots are addressed by the TOP key bits: the keys are bits of x * odd constant, whose low bits depend on the low bits of x alone
    // (a few bases of the mer: as skewed as the base composition), while the high bits mix all of x
    const int hshift = kb > LOGS ? kb - LOGS : 0;
    const bool aux = pl.aux != 0;
    const int auxshift = pl.pbits + 2;
    const u32 last1 = (u32)(pl.npos1 - 1);   // positions fit 32 bits (pairs are p0 | p1 << 32)

    if (tid == 0) {
        g3_mbar_init(&sm.bar[0], 1);
        if (NBUF == 2) g3_mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.losers = 0;
    }
    {   // the table starts empty; every bucket leaves it empty again
        uint4* t = reinterpret_cast<uint4*>(sm.tab);
        for (u32 i = tid; i < (4u << LOGS) / 16; i += BK_THREADS) t[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    // bucket descriptor: what arrived of the two genomes and the word holding the bucket's dirty bit.  Loaded two rounds ahead,
    // decoded when needed, so that nobody waits for the loads.
    struct Desc { unsigned long long c0, c1; u32 dirty; };
    auto fetch = [&](u64 f) {
        Desc d;
        d.c0 = d.c1 = 0; d.dirty = 0;
        if (f < nfinal) {
            const ulonglong2 c = __ldg(reinterpret_cast<const ulonglong2*>(a.cnt2) + f);
            d.c0 = c.x; d.c1 = c.y;
            d.dirty = __ldg(a.dirty + (f >> 5));
        }
        return d;
    };
    auto decode = [&](const Desc& d, u64 f, u32& n0, u32& n1, bool& spill) {
        n0 = (u32)min(d.c0, (unsigned long long)BK_HALF);
        n1 = (u32)min(d.c1, (unsigned long long)BK_HALF);
        spill = d.c0 > BK_HALF || d.c1 > BK_HALF || ((d.dirty >> (f & 31)) & 1u);
    };
    auto issue = [&](int buf, u64 f, u32 n0, u32 n1) {  // one thread; an odd count reads one record of slack inside the half's own 8 KB
        const u32 b0 = (n0 * 8u + 15u) & ~15u, b1 = (n1 * 8u + 15u) & ~15u;
        g3_mbar_expect_tx(&sm.bar[buf], b0 + b1);
        g3_bulk_load(sm.raw[buf], a.recs + f * BK_CAP, b0, &sm.bar[buf]);
        g3_bulk_load(sm.raw[buf] + BK_HALF, a.recs + f * BK_CAP + BK_HALF, b1, &sm.bar[buf]);
    };

    u64 f = blockIdx.x;
    u32 n0a, n1a, n0b, n1b;   // this bucket, the next
    bool spa, spb;
    {
        const Desc da = fetch(f), db = fetch(f + stride);
        decode(da, f, n0a, n1a, spa);
        decode(db, f + stride, n0b, n1b, spb);
    }
    if (tid == 0 && n0a && n1a && !spa) issue(0, f, n0a, n1a);
    Desc dc = fetch(f + 2 * stride);
    u32 parity = 0;  // bit b: phase parity buffer b completes next

    // pairs of the previous bucket: written one round late, so that the global reservations (tid < 4) have a round to come back
    u64 pe[G3_IPT];
    u32 pkinds = 0x4444u, pA = 0, pB = 0;    // kinds, this thread's offsets inside the bucket's four lists
    bool pany = false;                       // the previous bucket has pairs (uniform)
    unsigned long long resv = 0;             // tid < 4: reservation of list tid
#pragma unroll
    for (int it = 0; it < G3_IPT; ++it) pe[it] = 0;

    auto write_out = [&]() {
        u64* q0 = a.pairs + (sm.base[0] + (pA & 0xffffu));
        u64* q1 = a.pairs + (a.pair_cap - 1 - (sm.base[1] + (pA >> 16)));
#pragma unroll
        for (int it = 0; it < G3_IPT; ++it) {
            const u32 kind = (pkinds >> (4 * it)) & 15u;
            const u64 ee = pe[it];
            u64* q = kind ? q1 : q0;
            if (kind < 2) {
                *q = ee;
                atomicOr(&a.uniq[(u32)ee >> 5], 1u << ((u32)ee & 31));
            }
            q0 += kind == 0 ? 1 : 0;
            q1 -= kind == 1 ? 1 : 0;
        }
        if (pkinds & 0x2222u) {   // direct candidates: about one pair in a hundred
            u64* q2 = a.cand + (sm.base[2] + (pB & 0xffffu));
            u64* q3 = a.cand + (a.cand_cap - 1 - (sm.base[3] + (pB >> 16)));
#pragma unroll
            for (int it = 0; it < G3_IPT; ++it) {
                const u32 kind = (pkinds >> (4 * it)) & 15u;
                const u64 ee = pe[it];
                if (kind == 2) *q2++ = ee;
                if (kind == 3) *q3-- = ee;
                if ((kind & 6u) == 2u) atomicOr(&a.uniq[(u32)ee >> 5], 1u << ((u32)ee & 31));
            }
        }
    };

    for (u32 k = 0; f < nfinal; ++k, f += stride) {
        const int buf = NBUF == 2 ? (int)(k & 1) : 0;
        if (NBUF == 2 && tid == 0 && n0b && n1b && !spb) issue(buf ^ 1, f + stride, n0b, n1b);  // the other buffer was released by the barrier ending the previous round
        const Desc dn = fetch(f + 3 * stride);
        bool spill_it = spa;
        const u32 n0 = n0a, n1 = n1a;
        const bool work = n0 && n1 && !spill_it;   // a bucket with records of one genome only has no pair
        u32 losers = 0;
        const u64* __restrict__ raw = sm.raw[buf];
        if (work) {
            g3_mbar_wait(&sm.bar[buf], (parity >> buf) & 1u);
            parity ^= 1u << buf;
            // ---- build: genome 1 ----
            u32 old[G3_IPT], val[G3_IPT];
#pragma unroll
            for (int it = 0; it < G3_IPT; ++it) {
                const u32 i = it * BK_THREADS + tid;
                old[it] = 0;
                val[it] = 0;
                if (i < n1) {
                    const u32 kk = (u32)(raw[BK_HALF + i] >> kshift) & kmask;
                    val[it] = (kk << 4) | G3_SEEN1;
                    old[it] = atomicCAS(&sm.tab[(kk >> hshift) & SMASK], 0u, val[it]);
                }
            }
#pragma unroll
            for (int it = 0; it < G3_IPT; ++it) {
                const u32 i = it * BK_THREADS + tid;
                if (i < n1) {
                    u32 o = old[it], s = (val[it] >> (4 + hshift)) & SMASK;
                    while (o != 0 && ((o ^ val[it]) >> 4) != 0) {   // another key lives here: next slot
                        s = (s + 1) & SMASK;
                        o = atomicCAS(&sm.tab[s], 0u, val[it]);
                    }
                    if (o == 0) sm.idx1[s] = (unsigned short)i;
                    else { atomicOr(&sm.tab[s], G3_DUP1); ++losers; }
                }
                __syncwarp();   // the probing lanes rejoin here: without it the warp stays split for the records that follow
            }
        }
        if (pany && tid < 4) sm.base[tid] = resv;   // the previous bucket's reservations have had the build phase to arrive
        __syncthreads();
        u64 rec[G3_IPT];
        u32 slot[G3_IPT];
        if (work) {
            // ---- probe: genome 0 ----
            u32 t[G3_IPT];
#pragma unroll
            for (int it = 0; it < G3_IPT; ++it) {
                const u32 i = it * BK_THREADS + tid;
                rec[it] = i < n0 ? raw[i] : 0;
                slot[it] = (((u32)(rec[it] >> kshift) & kmask) >> hshift) & SMASK;   // (& kmask: the level-2 bin bits sit above the key bits)
                t[it] = i < n0 ? sm.tab[slot[it]] : 0u;
            }
#pragma unroll
            for (int it = 0; it < G3_IPT; ++it) {
                const u32 kk = (u32)(rec[it] >> kshift) & kmask;
                u32 v = t[it], s = slot[it];
                while (v != 0 && (v >> 4) != kk) {
                    s = (s + 1) & SMASK;
                    v = sm.tab[s];
                }
                slot[it] = NONE;
                if (v != 0 && !(v & G3_DUP1)) {   // (an empty slot: the mer does not occur in genome 1; also every i >= n0)
                    const u32 o = atomicOr(&sm.tab[s], G3_SEEN0);
                    if (o & G3_SEEN0) { atomicOr(&sm.tab[s], G3_DUP0); ++losers; }
                    slot[it] = s;
                } else if (v != 0) ++losers;      // one of many copies in genome 1: counted like a duplicate
                __syncwarp();
            }
            if (losers) atomicAdd(&sm.losers, losers);
        }
        if (pany) write_out();   // previous bucket
        pany = false;
        pkinds = 0x4444u;
        __syncthreads();
        if (work) {
            spill_it = sm.losers >= 999u;
            // ---- pairing: a genome-0 record whose key was seen once in each genome ----
            u32 cA = 0, cB = 0;       // forward | reverse << 16, forward direct | reverse direct << 16
            if (!spill_it) {
                pkinds = 0;           // 4 bits per record: 0 forward, 1 reverse, 2 forward direct, 3 reverse direct, 4 nothing
#pragma unroll
                for (int it = 0; it < G3_IPT; ++it) {
                    u32 kind = 4;
                    const u32 s = slot[it];
                    if (s != NONE && (sm.tab[s] & 15u) == (G3_SEEN0 | G3_SEEN1)) {
                        kind = g3_classify(rec[it], raw[BK_HALF + sm.idx1[s]], aux, auxshift, posmask, last1, pe[it]);
                        const u32 inc = 1u << ((kind & 1u) << 4);
                        cA += kind < 2 ? inc : 0u;
                        cB += kind < 2 ? 0u : inc;
                    }
                    pkinds |= kind << (4 * it);
                }
            }
            // warp prefix of the packed counts
            u32 iA = cA, iB = cB;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 tA = __shfl_up_sync(0xffffffffu, iA, o), tB = __shfl_up_sync(0xffffffffu, iB, o);
                if (lane >= (u32)o) { iA += tA; iB += tB; }
            }
            if (lane == 31) { sm.wtot[warp][0] = iA; sm.wtot[warp][1] = iB; }
            pA = iA - cA;
            pB = iB - cB;
        }
        __syncthreads();   // every read of the table and of raw[buf] is done
        if (NBUF == 1 && tid == 0 && n0b && n1b && !spb) issue(0, f + stride, n0b, n1b);
        if (work) {
            {   // ---- empty the table for the next bucket ----
                uint4* t = reinterpret_cast<uint4*>(sm.tab);
#pragma unroll
                for (u32 i = tid; i < (4u << LOGS) / 16; i += BK_THREADS) t[i] = make_uint4(0u, 0u, 0u, 0u);
                if (tid == 0) sm.losers = 0;
            }
            if (!spill_it) {
                u32 tA = 0, tB = 0;
#pragma unroll
                for (u32 w = 0; w < BK_THREADS / 32; ++w) {
                    const u32 xA = sm.wtot[w][0], xB = sm.wtot[w][1];
                    if (w < warp) { pA += xA; pB += xB; }
                    tA += xA; tB += xB;
                }
                pany = (tA | tB) != 0;
                if (tid < 4) {   // one reservation per list; consumed in the next round
                    const u32 tot = tid == 0 ? tA & 0xffffu : tid == 1 ? tA >> 16 : tid == 2 ? tB & 0xffffu : tB >> 16;
                    unsigned long long* ctr = tid == 0 ? &a.counters[0] : tid == 1 ? &a.counters[6] : tid == 2 ? &a.counters[2] : &a.counters[7];
                    resv = tot ? atomicAdd(ctr, (unsigned long long)tot) : 0ull;
                    if (tid >= 2 && tot) atomicAdd(&a.spill[3], (unsigned long long)tot);
                }
            }
        }
        if (spill_it && tid == 0) {
            a.spill_list[atomicAdd(&a.spill[0], 1ull)] = (u32)f;
            atomicAdd(&a.spill[1], (unsigned long long)(n0 + n1));
        }
        __syncthreads();  // the table is empty, the per-warp totals are read: the next round may begin
        n0a = n0b; n1a = n1b; spa = spb;
        decode(dc, f + 2 * stride, n0b, n1b, spb);
        dc = dn;
    }
    if (pany) {   // the last bucket's pairs
        if (tid < 4) sm.base[tid] = resv;
        __syncthreads();
        write_out();
    }
}

// spilled buckets -> (key, position) arrays of the radix-sort path: key = mixed mer << 2 | genome << 1 | strand
__global__ void __launch_bounds__(BK_THREADS) bk_spill_kernel(const u64* __restrict__ recs, const u64* __restrict__ off2, const BkMeta* __restrict__ meta,
                                                             const u32* __restrict__ spill_list, BkPlan pl, u64* __restrict__ keys,
                                                             u32* __restrict__ vals, unsigned long long* __restrict__ cursor,
                                                             const unsigned long long* __restrict__ cnt2)
{
    __shared__ unsigned long long s_base;
    const u32 f = spill_list[blockIdx.x];
    const u64 beg = cnt2 ? (u64)f * BK_CAP : off2[f];
    const u64 n0 = cnt2 ? min((u64)cnt2[2 * (u64)f], (u64)BK_HALF) : ~0ull;   // split layout: records [n0, nb) start at beg + BK_HALF
    const u64 nb = cnt2 ? n0 + min((u64)cnt2[2 * (u64)f + 1], (u64)BK_HALF) : off2[f + 1] - beg;
    if (threadIdx.x == 0) s_base = atomicAdd(cursor, (unsigned long long)nb);
    __syncthreads();
    const u64 b1 = meta->b_lo + f / pl.B2;
    const u64 posmask = (1ull << pl.pbits) - 1;
    for (u64 i = threadIdx.x; i < nb; i += blockDim.x) {
        const u64 r = recs[beg + (i < n0 ? i : i - n0 + BK_HALF)];
        const u64 mixed = (b1 << pl.rem1) | (r >> pl.kshift);  // the mixed mer: equal exactly when the mers are equal
        keys[s_base + i] = (mixed << 2) | ((r & 1) << 1) | ((r >> 1) & 1);
        vals[s_base + i] = (u32)((r >> 2) & posmask);
    }
}

// ---- host ---------------------------------------------------------------------------------------------------
static int bit_len(u64 x)
{
    int b = 0;
    while (x) { ++b; x >>= 1; }
    return b;
}

// `share`: 1 / fraction of the seeds this rank owns (shard count)
static bool make_plan(const SeedParams& sp, u64 npos0, u64 npos1, u64 share, BkPlan* out)
{
    BkPlan p;
    p.kbits = 2 * sp.w;
    p.npos0 = npos0; p.npos1 = npos1; p.ntot = npos0 + npos1;
    p.pbits = bit_len((npos0 > npos1 ? npos0 : npos1));
    if (p.pbits < 1) p.pbits = 1;
    // records this rank will hold, counting the smaller genome like the larger one: final buckets keep the two genomes' records
    // in separate halves, so the bucket count follows the larger genome
    const u64 expect = 2 * (npos0 > npos1 ? npos0 : npos1) / share;
    if (expect < 65536 || p.ntot >= (1ull << 40)) return false;
    int T = bit_len(expect / 1536);          // 768..1536 records per final bucket: BK_CAP is > 13 sigma away
    if (T > p.kbits - 2) T = p.kbits - 2;    // keep key bits for the in-bucket comparison
    if (T < 2 || T > 22) return false;
    p.d1 = (T + 1) / 2;
    p.w = sp.w;
    const bool solid = sp.nruns == 1 && sp.L == sp.w && (sp.w & 1) && getenv("MAUVE_CUDA_NO_SOLID") == nullptr;
    p.aux = solid && p.kbits - p.d1 + 2 + 4 + p.pbits <= 64 ? 4 : 0;
    // the record must fit 64 bits: (kbits - d1) + aux + strand + genome + pbits
    while (p.kbits - p.d1 + 2 + p.aux + p.pbits > 64 && p.d1 < 11 && p.d1 < T) ++p.d1;
    if (p.kbits - p.d1 + 2 + p.aux + p.pbits > 64) return false;
    p.kshift = p.pbits + 2 + p.aux;
    p.d2 = T - p.d1;
    if (p.d1 > 11 || p.d2 > 11 || p.d2 < 0) return false;
    p.rem1 = p.kbits - p.d1;
    p.d3 = p.rem1 - p.d2 < BK_D3 ? p.rem1 - p.d2 : BK_D3;
    if (p.d3 < 0) return false;
    p.B1 = 1u << p.d1;
    p.B2 = 1u << p.d2;
    p.npad0 = (npos0 + BK_TILE - 1) / BK_TILE * BK_TILE;
    p.nidx = p.npad0 + ((npos1 + 15) & ~15ull);
    *out = p;
    return true;
}

// Histogram-free run (see "fixed-capacity buckets" above).  *fell_back is set when the overflow arrays filled up: nothing
// usable was produced and the caller must reset uniq / counters and take the exact path.
static int bucket_group_fixed(Session& s, const SeedParams& sp, BkPlan pl, int shard_index, int shard_count, u64 pair_cap, cudaEvent_t ev_scatter1,
                              cudaEvent_t ev_scatter2, bool* fell_back, u64* nrecords)
{
    *fell_back = false;
    cudaStream_t st = s.stream;
    unsigned long long* ctr = s.counters.as<unsigned long long>();
    pl.b_lo = 0;   // seeds are owned by hash (seed_owned), so every bucket holds this rank's 1/shard_count share
    pl.b_hi = pl.B1;
    const u32 nb1 = pl.b_hi - pl.b_lo;
    for (int g = 0; g < 2; ++g) {
        const double mean1 = (double)(g ? pl.npos1 : pl.npos0) / (double)shard_count / pl.B1;
        pl.cap1g[g] = (u32)(((u64)(mean1 + 8.0 * sqrt(mean1) + 64.0) + 31) & ~31ull);
    }
    pl.cap1 = pl.cap1g[0] > pl.cap1g[1] ? pl.cap1g[0] : pl.cap1g[1];
    const u64 nfinal = (u64)nb1 * pl.B2;
    u64 ovf_cap = pl.ntot / 8 / (u64)shard_count + (1ull << 20);
    if (const char* e = getenv("MAUVE_CUDA_OVF_CAP")) ovf_cap = strtoull(e, nullptr, 10);  // tests: force the fallback
    const u64 dirty_words = nfinal / 32 + 1;
    MCU_TRY(s.bk_a.reserve(((u64)nb1 * ((u64)pl.cap1g[0] + pl.cap1g[1]) + 1) * 8));
    MCU_TRY(s.bk_b.reserve((nfinal * BK_CAP + 1) * 8));
    MCU_TRY(s.bk_tab1.reserve((size_t)(pl.B1 + 1) * 8 * 3 + 64));
    MCU_TRY(s.bk_tab2.reserve((nfinal + 1) * 8 * 3 + 64 + dirty_words * 4));
    MCU_TRY(s.bk_spill.reserve(nfinal * 4 + 64));
    MCU_TRY(s.keys_a.reserve((ovf_cap + 1) * 8));
    MCU_TRY(s.vals_a.reserve((ovf_cap + 1) * 4));
    unsigned long long* cursor1 = s.bk_tab1.as<unsigned long long>();
    BkMeta* meta = (BkMeta*)(cursor1 + (size_t)(pl.B1 + 1) * 3);  // inside the +64 bytes slack, as in the exact path
    unsigned long long* cursor2 = s.bk_tab2.as<unsigned long long>();
    unsigned long long* spill = cursor2 + 2 * nfinal + 2;  // [0] buckets [1] records [2] overflow / convert cursor [3] direct candidates
    u32* dirty = (u32*)(spill + 4);
    MCU_CUDA(cudaMemsetAsync(cursor1, 0, (size_t)(2 * pl.B1 + 1) * 8, st));
    MCU_CUDA(cudaMemsetAsync(cursor2, 0, (2 * nfinal + 2) * 8 + 32 + dirty_words * 4, st));
    const u32* g0 = s.packed[0].as<u32>();
    const u32* g1 = s.packed[1].as<u32>();
    static bool attr_done = false;
    if (!attr_done) {
        MCU_CUDA(cudaFuncSetAttribute(bkf_scatter1_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        MCU_CUDA(cudaFuncSetAttribute(bkf_scatter1_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        MCU_CUDA(cudaFuncSetAttribute(bkf_scatter1_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        MCU_CUDA(cudaFuncSetAttribute(bkf_scatter2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        MCU_CUDA(cudaFuncSetAttribute(bkf_scatter2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        MCU_CUDA(cudaFuncSetAttribute(bkf_scatter2_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        attr_done = true;
    }
    BkOvf ovf;
    ovf.keys = s.keys_a.as<u64>(); ovf.vals = s.vals_a.as<u32>(); ovf.cap = ovf_cap; ovf.cursor = spill + 2; ovf.dirty = dirty;
    const bool solid_pattern = sp.nruns == 1 && sp.L == sp.w && getenv("MAUVE_CUDA_NO_SOLID") == nullptr;
    const unsigned tiles1 = (unsigned)div_up(pl.nidx, BK_TILE);
    MCU_CUDA(cudaEventRecord(s.kev[0], st));
    MCU_CUDA(cudaEventRecord(s.kev[1], st));
    const u32 multi_sub = (solid_pattern && shard_count > 1 && getenv("MAUVE_CUDA_NO_MULTI") == nullptr) ? (u32)((7 * shard_count) / 8) : 0;
    auto scatter1 = [&](unsigned t0, unsigned t1) {
        if (t1 <= t0) return;
        if (multi_sub > 1 && t0 == 0 && t1 == tiles1) {   // the whole scan in one launch: `multi_sub` tiles of one genome per CTA
            static bool attr_multi = false;
            if (!attr_multi) {
                cudaFuncSetAttribute(bkf_scatter1_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bkf_multi_smem_bytes(BK_MAXB));
                attr_multi = true;
            }
            const u32 tiles0 = (u32)(pl.npad0 / BK_TILE);
            const u32 ctas0 = (tiles0 + multi_sub - 1) / multi_sub, ctas1 = (tiles1 - tiles0 + multi_sub - 1) / multi_sub;
            bkf_scatter1_multi_kernel<<<ctas0 + ctas1, BK_THREADS, bkf_multi_smem_bytes(pl.B1), st>>>(g0, g1, pl, sp, cursor1, s.bk_a.as<u64>(), ovf, tiles0,
                                                                                                        tiles1, multi_sub, ctas0);
            s.launches++;
            return;
        }
        static const int s1_minb = getenv("MAUVE_CUDA_S1_MINB") ? atoi(getenv("MAUVE_CUDA_S1_MINB")) : 3;   // A/B: registers vs occupancy
        if (solid_pattern && s1_minb == 4) bkf_scatter1_kernel<true, 4><<<t1 - t0, BK_THREADS, bk_scatter_smem_bytes(pl.B1), st>>>(g0, g1, pl, sp, cursor1, s.bk_a.as<u64>(), ovf, t0);
        else if (solid_pattern) bkf_scatter1_kernel<true, 3><<<t1 - t0, BK_THREADS, bk_scatter_smem_bytes(pl.B1), st>>>(g0, g1, pl, sp, cursor1, s.bk_a.as<u64>(), ovf, t0);
        else bkf_scatter1_kernel<false, 3><<<t1 - t0, BK_THREADS, bk_scatter_smem_bytes(pl.B1), st>>>(g0, g1, pl, sp, cursor1, s.bk_a.as<u64>(), ovf, t0);
        s.launches++;
    };
    bool scatter2_done = false;
    if (s.up_chunks > 0 && s.pack_world == 1) {
        // chunked upload in flight (session_upload_begin): pack every piece as it lands and scatter the tiles whose seeds it
        // completes.  A seed window (and the neighbour bases of aux records) reads at most 64 bases past its position.
        unsigned done = 0;
        for (int g = 0; g < 2; ++g) MCU_TRY(s.packed[g].reserve((div_up(s.n[g], 16) + 2) * sizeof(u32)));
        g0 = s.packed[0].as<u32>();
        g1 = s.packed[1].as<u32>();
        for (int g = 0; g < 2; ++g)
            for (int c = 0; c < s.up_chunks; ++c) {
                u64 ready = 0;
                MCU_TRY(run_pack_piece(s, g, c, (u32*)(ctr + 4), &ready));
                const bool last = c == s.up_chunks - 1;
                const u64 npos_g = g ? pl.npos1 : pl.npos0;
                u64 pos_ok = last ? npos_g : (ready > 64 ? ready - 64 : 0);  // positions of genome g whose windows are packed
                if (pos_ok > npos_g) pos_ok = npos_g;
                u64 idx_ok = g ? pl.npad0 + pos_ok : pos_ok;
                if (last) idx_ok = g ? pl.nidx : pl.npad0;  // padding positions produce no record: the tile may run
                // tiles wholly below idx_ok (genome 0's last tile usually straddles into genome 1: it waits for genome 1's first piece)
                unsigned upto = (g == 1 && last) ? tiles1 : (unsigned)(idx_ok / BK_TILE);
                if (upto > tiles1) upto = tiles1;
                if (nb1) scatter1(done, upto);
                if (upto > done) done = upto;
                if (nb1 && last) {   // all tiles of genome g are out: its level-1 segments can be partitioned while the rest still arrives
                    const u32 capg = g ? pl.cap1g[1] : pl.cap1g[0];
                    bkf_scatter2_kernel<3><<<dim3((unsigned)div_up(capg, BK_TILE), nb1), BK_THREADS, bk_scatter_smem_bytes(pl.B2), st>>>(
                        s.bk_a.as<u64>(), pl, cursor1, cursor2, s.bk_b.as<u64>(), ovf, (u32)g);
                    s.launches++;
                    scatter2_done = true;
                }
            }
        s.up_chunks = 0;
    } else if (nb1) scatter1(0, tiles1);
    MCU_CUDA(cudaEventRecord(ev_scatter1, st));
    MCU_CUDA(cudaEventRecord(s.kev[2], st));
    bkf_total_kernel<<<1, 256, 0, st>>>(cursor1, pl, meta);
    MCU_CUDA(cudaEventRecord(s.kev[3], st));
    if (nb1 && !scatter2_done) {
        const dim3 grid2((unsigned)div_up(pl.cap1, BK_TILE), 2 * nb1);
        static const int s2_minb = getenv("MAUVE_CUDA_S2_MINB") ? atoi(getenv("MAUVE_CUDA_S2_MINB")) : 4;   // measured (100 Mbp pair): 3 CTAs/SM (72 registers) 0.986 ms, 4 (64) 0.945
        if (s2_minb == 5) bkf_scatter2_kernel<5><<<grid2, BK_THREADS, bk_scatter_smem_bytes(pl.B2), st>>>(s.bk_a.as<u64>(), pl, cursor1, cursor2, s.bk_b.as<u64>(), ovf, 2u);
        else if (s2_minb == 4) bkf_scatter2_kernel<4><<<grid2, BK_THREADS, bk_scatter_smem_bytes(pl.B2), st>>>(s.bk_a.as<u64>(), pl, cursor1, cursor2, s.bk_b.as<u64>(), ovf, 2u);
        else bkf_scatter2_kernel<3><<<grid2, BK_THREADS, bk_scatter_smem_bytes(pl.B2), st>>>(s.bk_a.as<u64>(), pl, cursor1, cursor2, s.bk_b.as<u64>(), ovf, 2u);
    }
    MCU_CUDA(cudaEventRecord(ev_scatter2, st));
    MCU_CUDA(cudaEventRecord(s.kev[4], st));
    BkGroupArgs ga;
    ga.recs = s.bk_b.as<u64>(); ga.off2 = nullptr; ga.meta = meta; ga.uniq = s.uniq.as<u32>(); ga.pairs = s.pairs.as<u64>(); ga.pair_cap = pair_cap;
    ga.counters = ctr; ga.spill_list = s.bk_spill.as<u32>(); ga.spill = spill;
    ga.cand = s.cand.as<u64>(); ga.cand_cap = pair_cap;
    ga.cnt2 = cursor2; ga.dirty = dirty; ga.nfinal = nfinal;
    // bk_group3 keeps <= 28 key bits that differ inside a final bucket next to four flag bits; wider keys (not reached by the seed
    // tables at sizes that take this path) keep bk_group.  MAUVE_CUDA_GROUP_V1 / MAUVE_CUDA_GROUP_VARIANT: A/B switches.
    static const bool group_v1_env = getenv("MAUVE_CUDA_GROUP_V1") != nullptr;
    static const int variant = getenv("MAUVE_CUDA_GROUP_VARIANT") ? atoi(getenv("MAUVE_CUDA_GROUP_VARIANT")) : 3;   // measured (100 Mbp pair): 3: 1.64 ms, 5: 1.65, 1: 1.83, 4: 1.91, 0: 2.34
    const bool group_v1 = group_v1_env || pl.rem1 - pl.d2 > G3_MAX_KEY_BITS;
    if (nfinal && group_v1) bk_group_kernel<<<(unsigned)nfinal, BK_THREADS, 0, st>>>(ga, pl);
    else if (nfinal) {
        // persistent: as many CTAs as fit, each walks the buckets with stride gridDim.x
        auto launch = [&](auto kernel, size_t smem, int ctas) -> int {
            MCU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const u64 want = (u64)sm_count() * ctas;
            kernel<<<(unsigned)(nfinal < want ? nfinal : want), BK_THREADS, smem, st>>>(ga, pl);
            return MCU_OK;
        };
        switch (variant) {
        case 1: MCU_TRY(launch(bk_group3_kernel<12, 2, 3>, sizeof(G3Smem<12, 2>), 3)); break;
        case 2: MCU_TRY(launch(bk_group3_kernel<11, 2, 5>, sizeof(G3Smem<11, 2>), 5)); break;
        case 0: MCU_TRY(launch(bk_group3_kernel<11, 2, 4>, sizeof(G3Smem<11, 2>), 4)); break;
        case 4: MCU_TRY(launch(bk_group3_kernel<11, 1, 6>, sizeof(G3Smem<11, 1>), 6)); break;
        case 5: MCU_TRY(launch(bk_group3_kernel<12, 1, 4>, sizeof(G3Smem<12, 1>), 4)); break;
        default: MCU_TRY(launch(bk_group3_kernel<12, 1, 5>, sizeof(G3Smem<12, 1>), 5)); break;   // one buffer, 4096-slot table, 5 CTAs per SM
        }
    }
    MCU_CUDA(cudaEventRecord(s.kev[5], st));
    s.launches += 3;
    MCU_CUDA(cudaGetLastError());
    struct { unsigned long long spill[4]; BkMeta meta; unsigned long long ctr[8]; } h;
    MCU_CUDA(cudaMemcpyAsync(h.spill, spill, 32, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaMemcpyAsync(&h.meta, meta, sizeof(BkMeta), cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaMemcpyAsync(h.ctr, ctr, 64, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    // pairs listed so far came from bk_group (neighbour bases already compared in aux mode); the sort path may append more
    s.bk_group_fwd = h.ctr[0];
    s.bk_group_rev = h.ctr[6];
    *nrecords = h.meta.nrec;
    const u64 novf = h.spill[2], ns = novf + h.spill[1];
    if (novf > ovf_cap || ns > ovf_cap) { *fell_back = true; return MCU_OK; }
    s.bk_spilled = ns;
    s.bk_direct = h.spill[3];
    s.bk_aux = pl.aux != 0;
    if (ns) {
        MCU_TRY(s.keys_b.reserve((ns + 1) * 8));
        MCU_TRY(s.vals_b.reserve((ns + 1) * 4));
        if (h.spill[0]) {
            bk_spill_kernel<<<(unsigned)h.spill[0], BK_THREADS, 0, st>>>(s.bk_b.as<u64>(), nullptr, meta, s.bk_spill.as<u32>(), pl, s.keys_a.as<u64>(),
                                                                         s.vals_a.as<u32>(), spill + 2, cursor2);
            s.launches++;
        }
        bool in_a = true;
        u64 before = s.radix.launches;
        MCU_TRY(radix_sort_pairs<u64>(s.radix, s.keys_a.as<u64>(), s.vals_a.as<u32>(), s.keys_b.as<u64>(), s.vals_b.as<u32>(), ns, pl.kbits + 2, false,
                                      st, &in_a, nullptr));
        s.launches += s.radix.launches - before;
        MCU_TRY(join_sorted_u64(s, in_a ? s.keys_a.as<u64>() : s.keys_b.as<u64>(), in_a ? s.vals_a.as<u32>() : s.vals_b.as<u32>(), ns, pair_cap));
    }
    return MCU_OK;
}

bool bucket_plan_applies(const SeedParams& sp, u64 npos0, u64 npos1, int shard_count)
{
    BkPlan pl;
    return npos0 && npos1 && getenv("MAUVE_CUDA_EXACT_BUCKETS") == nullptr && make_plan(sp, npos0, npos1, (u64)shard_count, &pl);
}

int bucket_group(Session& s, const SeedParams& sp, int shard_index, int shard_count, u64 pair_cap, cudaEvent_t ev_scatter1, cudaEvent_t ev_scatter2,
                 bool* used, u64* nrecords)
{
    *used = false;
    const u64 npos0 = s.n[0] >= (u64)sp.L ? s.n[0] - sp.L + 1 : 0;
    const u64 npos1 = s.n[1] >= (u64)sp.L ? s.n[1] - sp.L + 1 : 0;
    BkPlan pl;
    if (!npos0 || !npos1 || !make_plan(sp, npos0, npos1, (u64)shard_count, &pl)) return MCU_OK;
    pl.shard = (u32)shard_index;
    pl.nshard = (u32)shard_count;
    cudaStream_t st = s.stream;
    unsigned long long* ctr = s.counters.as<unsigned long long>();
    if (getenv("MAUVE_CUDA_EXACT_BUCKETS") == nullptr) {
        bool fell_back = false;
        MCU_TRY(bucket_group_fixed(s, sp, pl, shard_index, shard_count, pair_cap, ev_scatter1, ev_scatter2, &fell_back, nrecords));
        if (!fell_back) { *used = true; s.bk_exact = false; return MCU_OK; }
        // too many copies of too many mers for the overflow arrays: start over with exact bucket sizes
        const u64 uniq_words = div_up(npos0 + 1, 32) + 1;
        u32 gap_flag = 0;
        MCU_CUDA(cudaMemcpyAsync(&gap_flag, ctr + 4, 4, cudaMemcpyDeviceToHost, st));
        MCU_CUDA(cudaStreamSynchronize(st));
        MCU_CUDA(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), st));
        MCU_CUDA(cudaMemcpyAsync(ctr + 4, &gap_flag, 4, cudaMemcpyHostToDevice, st));
        MCU_CUDA(cudaMemsetAsync(s.uniq.p, 0, uniq_words * sizeof(u32), st));
        s.bk_fallbacks++;
    }
    pl.b_lo = pl.b_hi = pl.cap1 = 0;
    s.bk_exact = true;
    const u64 nfinal_max = (u64)pl.B1 * pl.B2;
    MCU_TRY(s.bk_a.reserve((pl.ntot + 1) * 8));
    MCU_TRY(s.bk_b.reserve((pl.ntot + 1) * 8));
    MCU_TRY(s.bk_tab1.reserve((size_t)(pl.B1 + 1) * 8 * 3 + 64));
    MCU_TRY(s.bk_tab2.reserve((nfinal_max + 1) * 8 * 3 + 64));
    MCU_TRY(s.bk_spill.reserve(nfinal_max * 4 + 64));
    MCU_TRY(s.bk_tileseg.reserve((div_up(pl.ntot, BK_TILE) + 1) * 4));
    unsigned long long* count1 = s.bk_tab1.as<unsigned long long>();
    u64* off1 = (u64*)(count1 + pl.B1 + 1);
    unsigned long long* cursor1 = (unsigned long long*)(off1 + pl.B1 + 1);
    BkMeta* meta = (BkMeta*)(cursor1 + pl.B1);  // inside the +64 bytes slack
    unsigned long long* count2 = s.bk_tab2.as<unsigned long long>();
    u64* off2 = (u64*)(count2 + nfinal_max + 1);
    unsigned long long* cursor2 = (unsigned long long*)(off2 + nfinal_max + 1);
    unsigned long long* spill = (unsigned long long*)(cursor2 + nfinal_max);  // [0] buckets [1] records [2] convert cursor
    MCU_CUDA(cudaMemsetAsync(count1, 0, (size_t)(pl.B1 + 1) * 8, st));
    MCU_CUDA(cudaMemsetAsync(count2, 0, (nfinal_max + 1) * 8, st));
    MCU_CUDA(cudaMemsetAsync(spill, 0, 32, st));
    const u32* g0 = s.packed[0].as<u32>();
    const u32* g1 = s.packed[1].as<u32>();
    static bool attr_done = false;
    if (!attr_done) {
        MCU_CUDA(cudaFuncSetAttribute(bk_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        MCU_CUDA(cudaFuncSetAttribute(bk_scatter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem_bytes(BK_MAXB)));
        attr_done = true;
    }
    const int hgrid = sm_count() * 8;
    MCU_CUDA(cudaEventRecord(s.kev[0], st));
    bk_hist1_kernel<<<hgrid, BK_THREADS, pl.B1 * 4, st>>>(g0, g1, pl, sp, count1);
    MCU_CUDA(cudaEventRecord(s.kev[1], st));
    bk_scan1_kernel<<<1, 32, 0, st>>>(count1, pl, shard_index, shard_count, off1, cursor1, meta);
    const unsigned tiles1 = (unsigned)div_up(pl.nidx, BK_TILE);
    const unsigned tiles = (unsigned)div_up(pl.ntot, BK_TILE);
    bk_scatter1_kernel<<<tiles1, BK_THREADS, bk_scatter_smem_bytes(pl.B1), st>>>(g0, g1, pl, sp, meta, cursor1, s.bk_a.as<u64>());
    MCU_CUDA(cudaEventRecord(ev_scatter1, st));
    MCU_CUDA(cudaEventRecord(s.kev[2], st));
    bk_tileseg_kernel<<<(tiles + 255) / 256, 256, 0, st>>>(off1, pl, meta, s.bk_tileseg.as<u32>(), tiles);
    bk_hist2_kernel<<<hgrid, BK_THREADS, pl.B2 * 4, st>>>(s.bk_a.as<u64>(), pl, meta, off1, count2);
    MCU_CUDA(cudaEventRecord(s.kev[3], st));
    bk_scan2_kernel<<<1, 1024, 0, st>>>(count2, meta, off2, cursor2);
    bk_scatter2_kernel<<<tiles, BK_THREADS, bk_scatter_smem_bytes(pl.B2), st>>>(s.bk_a.as<u64>(), pl, meta, off1, s.bk_tileseg.as<u32>(), cursor2, s.bk_b.as<u64>());
    MCU_CUDA(cudaEventRecord(ev_scatter2, st));
    MCU_CUDA(cudaEventRecord(s.kev[4], st));
    BkGroupArgs ga;
    ga.recs = s.bk_b.as<u64>(); ga.off2 = off2; ga.meta = meta; ga.uniq = s.uniq.as<u32>(); ga.pairs = s.pairs.as<u64>(); ga.pair_cap = pair_cap;
    ga.counters = ctr; ga.spill_list = s.bk_spill.as<u32>(); ga.spill = spill;
    ga.cand = s.cand.as<u64>(); ga.cand_cap = pair_cap;
    ga.cnt2 = nullptr; ga.dirty = nullptr; ga.nfinal = nfinal_max;
    bk_group_kernel<<<(unsigned)nfinal_max, BK_THREADS, 0, st>>>(ga, pl);
    MCU_CUDA(cudaEventRecord(s.kev[5], st));
    s.launches += 8;
    MCU_CUDA(cudaGetLastError());
    // spilled buckets (if any) go through the radix-sort + join path
    struct { unsigned long long spill[4]; BkMeta meta; unsigned long long ctr[8]; } h;
    MCU_CUDA(cudaMemcpyAsync(h.spill, spill, 32, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaMemcpyAsync(&h.meta, meta, sizeof(BkMeta), cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaMemcpyAsync(h.ctr, ctr, 64, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    // pairs listed so far came from bk_group (neighbour bases already compared in aux mode); the sort path may append more
    s.bk_group_fwd = h.ctr[0];
    s.bk_group_rev = h.ctr[6];
    *nrecords = h.meta.nrec;
    s.bk_spilled = h.spill[1];
    s.bk_direct = h.spill[3];
    s.bk_aux = pl.aux != 0;
    if (h.spill[0]) {
        const u64 ns = h.spill[1];
        MCU_TRY(s.keys_a.reserve((ns + 1) * 8));
        MCU_TRY(s.keys_b.reserve((ns + 1) * 8));
        MCU_TRY(s.vals_a.reserve((ns + 1) * 4));
        MCU_TRY(s.vals_b.reserve((ns + 1) * 4));
        bk_spill_kernel<<<(unsigned)h.spill[0], BK_THREADS, 0, st>>>(s.bk_b.as<u64>(), off2, meta, s.bk_spill.as<u32>(), pl, s.keys_a.as<u64>(),
                                                                     s.vals_a.as<u32>(), spill + 2, nullptr);
        s.launches++;
        bool in_a = true;
        u64 before = s.radix.launches;
        MCU_TRY(radix_sort_pairs<u64>(s.radix, s.keys_a.as<u64>(), s.vals_a.as<u32>(), s.keys_b.as<u64>(), s.vals_b.as<u32>(), ns, pl.kbits + 2, false,
                                      st, &in_a, nullptr));
        s.launches += s.radix.launches - before;
        MCU_TRY(join_sorted_u64(s, in_a ? s.keys_a.as<u64>() : s.keys_b.as<u64>(), in_a ? s.vals_a.as<u32>() : s.vals_b.as<u32>(), ns, pair_cap));
    }
    *used = true;
    return MCU_OK;
}

}  // namespace mcu
