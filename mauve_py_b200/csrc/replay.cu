// Bucket replay: the part of MemHash::AddHashEntry (LM/MemHash.cpp:209-251) that depends on insertion order.
//
// The reference keeps every hash bucket (generalized offset mod 40000) as a vector sorted by
// MheCompare (LM/MatchHashEntry.h:121-143) and looks a new seed up with std::lower_bound.  MheCompare
// mixes containment ("equal") with start order, so it is not a strict weak order once a bucket
// holds matches of different diagonals whose genome-0 spans interleave: the binary search can
// then walk past the match that contains the seed, the seed is extended again and the SAME match
// is stored twice (observed on MDS42 with seed 0x7954f: 77,056 rows, one of them a duplicate).
// Which seeds trip this depends on the bucket content at that moment, i.e. on the mer-sorted
// insertion order of MatchFinder::SearchRange.
//
// A bucket is "clean" when, in genome-0 start order, every match starts strictly after the end
// of its predecessor (start0 > prev.start0 + prev.len).  In a clean bucket every comparison of
// the binary search is decided by start0 alone and agrees with the vector order, so the search
// always finds the containing match: the bucket ends up as the distinct matches in start order,
// which is what the parallel closed form produces.  Matches of one diagonal never violate the
// condition (consecutive chains are more than L apart), so unclean buckets need two diagonals
// that collide mod 40000 with interleaved spans: a handful per genome pair.
//
// Unclean buckets are replayed exactly: (1) collect every unique seed pair inside their matches,
// (2) sort them by canonical mer (= the reference's insertion order; unique seeds have distinct
// mers), stable-sort by bucket, (3) one thread per bucket runs the reference's lower_bound /
// equality test / sorted insert on the bucket's vector of row indices -- the extension result of a
// seed is the closed-form match that contains it, so no genome access is needed here --, (4) the
// final list is reassembled with the replayed buckets in place.
#include "anchor.cuh"

#include <utility>

namespace mcu {

constexpr u32 RP_TABLE = 40000;  // DEFAULT_MEM_TABLE_SIZE, LM/MemHash.h:30
constexpr u32 RP_WORDS = (RP_TABLE + 31) / 32;

struct RpEnt {
    i64 len, s0, s1, off, mersize;
};

__device__ __forceinline__ i64 rp_offset(i64 len, i64 s0, i64 s1) { return s1 - s0 - (s1 < 0 ? len : 0); }  // CalculateOffset, LM/MatchHashEntry.cpp:141-160

__device__ __forceinline__ u32 rp_bucket(const mcu_match& m)
{
    i64 o = rp_offset(m.len, m.start0, m.start1) % (i64)RP_TABLE;
    if (o < 0) o += RP_TABLE;
    return (u32)o;
}

__device__ __forceinline__ RpEnt rp_from_row(const mcu_match& m)
{
    RpEnt e;
    e.len = m.len; e.s0 = m.start0; e.s1 = m.start1;
    e.off = rp_offset(m.len, m.start0, m.start1);
    e.mersize = 0;  // stored entries went through operator=, which zeroes m_mersize (LM/MatchHashEntry.cpp:112-120)
    return e;
}

// MatchHashEntry::Contains, LM/MatchHashEntry.cpp:164-200 (two sequences, both defined)
__device__ __forceinline__ bool rp_contains(const RpEnt& x, const RpEnt& y)
{
    if (x.off != y.off) return false;
    const i64 diff = y.s0 - x.s0;
    if (x.s0 == 0) return false;
    if (diff < 0 || x.len < y.len + diff) return false;
    const i64 diff_rc = y.len - x.len + diff, diff_i = y.s1 - x.s1;
    if (y.s1 < 0 && diff_rc == diff_i) return true;
    return diff == diff_i;
}

// MatchHashEntry::strict_start_lessthan_ptr, LM/MatchHashEntry.cpp:48-67
__device__ __forceinline__ bool rp_start_less(const RpEnt& a, const RpEnt& b)
{
    i64 d = a.s0 - b.s0;  // genome-0 starts are always positive
    if (d != 0) return d < 0;
    const i64 as = a.s1 < 0 ? -a.s1 + a.len - a.mersize : a.s1;
    const i64 bs = b.s1 < 0 ? -b.s1 + b.len - b.mersize : b.s1;
    d = as - bs;
    return d < 0;
}

// MheCompare, LM/MatchHashEntry.h:121-143
__device__ __forceinline__ bool rp_less(const RpEnt& a, const RpEnt& b)
{
    if (rp_contains(a, b) || rp_contains(b, a)) return false;
    return rp_start_less(a, b);
}

// ---- detection ------------------------------------------------------------------------------
// ctr: [0] unclean buckets, [1] rows in unclean buckets, [2] seeds collected, [3] pool cursor, [4] extra rows
__global__ void rp_detect_kernel(const mcu_match* __restrict__ rows, u64 n, u32* __restrict__ bitmap, unsigned long long* __restrict__ ctr)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += (u64)gridDim.x * blockDim.x) {
        const mcu_match a = rows[i], b = rows[i + 1];
        const u32 bk = rp_bucket(a);
        if (bk == rp_bucket(b) && b.start0 <= a.start0 + a.len) {
            const u32 bit = 1u << (bk & 31);
            if (!(atomicOr(&bitmap[bk >> 5], bit) & bit)) atomicAdd(&ctr[0], 1ull);
        }
    }
}

__global__ void rp_list_kernel(const u32* __restrict__ bitmap, u32* __restrict__ list, unsigned long long* __restrict__ cursor)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < RP_TABLE && ((bitmap[b >> 5] >> (b & 31)) & 1u)) list[atomicAdd(cursor, 1ull)] = b;
}

// ---- seed collection: one warp per row of an unclean bucket -----------------------------------------
struct RpCollect {
    const mcu_match* rows;
    u64 n;
    const u32* bitmap;
    u64* canon;
    u32* p0;
    u32* row;
    u64 cap;
    unsigned long long* ctr;
};

__global__ void __launch_bounds__(256) rp_collect_kernel(RpCollect c, ExtendArgs a, SeedParams sp, int count_only)
{
    const u32 lane = threadIdx.x & 31;
    const u64 wid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 i = wid; i < c.n; i += nw) {
        const mcu_match m = c.rows[i];
        const u32 bk = rp_bucket(m);
        if (!((c.bitmap[bk >> 5] >> (bk & 31)) & 1u)) continue;
        if (lane == 0 && count_only) atomicAdd(&c.ctr[1], 1ull);
        const bool rev = m.start1 < 0;
        const i64 lo = m.start0 - 1, hi = lo + m.len - sp.L;
        const i64 d = rev ? hi + (-m.start1) - 1 : m.start1 - m.start0;
        for (i64 t0 = lo; t0 <= hi; t0 += 32) {
            const i64 t = t0 + lane;
            i64 other;
            const bool sd = t <= hi && uniq_bit(a.uniq, t) && probe_hit(a, sp, rev, d, t, other);
            const u32 msk = __ballot_sync(0xffffffffu, sd);
            if (!msk) continue;
            u64 base = 0;
            if (lane == 0) base = atomicAdd(&c.ctr[2], (unsigned long long)__popc(msk));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sd && !count_only) {
                const u64 slot = base + __popc(msk & lanemask_lt());
                if (slot < c.cap) {
                    const u64 f = extract_seed(load_mer32(a.g0, (u64)t), sp), rc = revcomp_seed(f, sp.w);
                    c.canon[slot] = rc < f ? rc : f;
                    c.p0[slot] = (u32)t;
                    c.row[slot] = (u32)i;
                }
            }
        }
    }
}

__global__ void rp_iota_kernel(u32* __restrict__ idx, u64 n)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) idx[i] = (u32)i;
}

__global__ void rp_bucket_keys_kernel(const mcu_match* __restrict__ rows, const u32* __restrict__ row, const u32* __restrict__ idx, u64 n,
                                      u64* __restrict__ keys)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) keys[i] = rp_bucket(rows[row[idx[i]]]);
}

// ---- replay: one thread per unclean bucket ------------------------------------------------------------
struct RpReplay {
    const mcu_match* rows;
    u64 n;
    const u32* list;
    u32 nlist;
    const u64* bkeys;   // bucket id of every collected seed, sorted (mer order inside a bucket)
    const u32* perm;    // seed index in that order
    const u32* p0;
    const u32* row;
    u64 nseeds;
    u32* pool;
    u64 pool_cap;
    u32* extra;         // [RP_TABLE] rows added per bucket
    u64* vinfo;         // per list entry: {pool base, count, first row}
    unsigned long long* ctr;
    int L;
};

__device__ __forceinline__ u64 rp_lower_rows(const mcu_match* rows, u64 n, u32 bucket)
{
    u64 lo = 0, hi = n;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (rp_bucket(rows[mid]) < bucket) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ u64 rp_lower_keys(const u64* keys, u64 n, u64 v)
{
    u64 lo = 0, hi = n;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// std::lower_bound (libstdc++ __lower_bound) over the bucket vector with MheCompare(element, value)
__device__ __forceinline__ u32 rp_lower_bound(const mcu_match* rows, const u32* V, u32 nv, const RpEnt& val)
{
    u32 first = 0, len = nv;
    while (len > 0) {
        const u32 half = len >> 1, mid = first + half;
        if (rp_less(rp_from_row(rows[V[mid]]), val)) { first = mid + 1; len = len - half - 1; }
        else len = half;
    }
    return first;
}

__global__ void __launch_bounds__(64) rp_replay_kernel(RpReplay r)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= r.nlist) return;
    const u32 b = r.list[k];
    const u64 rb = rp_lower_rows(r.rows, r.n, b), re = rp_lower_rows(r.rows, r.n, b + 1);
    const u64 sb = rp_lower_keys(r.bkeys, r.nseeds, b), se = rp_lower_keys(r.bkeys, r.nseeds, (u64)b + 1);
    const u64 cap = (re - rb) + (se - sb);
    const u64 base = atomicAdd(&r.ctr[3], (unsigned long long)cap);
    u32 nv = 0;
    if (base + cap <= r.pool_cap) {
        u32* V = r.pool + base;
        for (u64 q = sb; q < se; ++q) {
            const u32 si = r.perm[q];
            const u32 y = r.row[si];
            const mcu_match my = r.rows[y];
            const bool rev = my.start1 < 0;
            const i64 p0 = r.p0[si];
            const i64 lo = my.start0 - 1, hi = lo + my.len - r.L;
            const i64 d = rev ? hi + (-my.start1) - 1 : my.start1 - my.start0;
            const i64 p1 = rev ? d - p0 : p0 + d;
            RpEnt s;  // the seed as HashMatch + SetDirection build it (LM/MemHash.cpp:167-203)
            s.len = r.L; s.mersize = r.L;
            s.s0 = p0 + 1;
            s.s1 = rev ? -(p1 + 1) : p1 + 1;
            s.off = rp_offset(s.len, s.s0, s.s1);
            u32 at = rp_lower_bound(r.rows, V, nv, s);
            if (at < nv) {
                const RpEnt e = rp_from_row(r.rows[V[at]]);
                if (!rp_less(e, s) && !rp_less(s, e)) continue;  // ++m_collision_count
            }
            const RpEnt ext = rp_from_row(my);                    // ExtendMatch(mhe): the maximal match containing the seed
            at = rp_lower_bound(r.rows, V, nv, ext);
            for (u32 j = nv; j > at; --j) V[j] = V[j - 1];
            V[at] = y;
            ++nv;
        }
    }
    r.vinfo[3 * (u64)k] = base;
    r.vinfo[3 * (u64)k + 1] = nv;
    r.vinfo[3 * (u64)k + 2] = rb;
    const u32 ex = nv >= (u32)(re - rb) ? nv - (u32)(re - rb) : 0u;
    r.extra[b] = ex;
    if (ex) atomicAdd(&r.ctr[4], (unsigned long long)ex);
    if (nv < (u32)(re - rb)) atomicAdd(&r.ctr[5], 1ull);  // cannot happen: every row owns at least one unique seed
}

// exclusive scan of extra[RP_TABLE] in one block
__global__ void __launch_bounds__(1024) rp_scan_kernel(const u32* __restrict__ extra, u64* __restrict__ prefix)
{
    __shared__ u64 s_tot[1024];
    constexpr u32 PER = (RP_TABLE + 1023) / 1024;
    const u32 t = threadIdx.x;
    u64 local = 0;
    for (u32 j = 0; j < PER; ++j) {
        const u32 i = t * PER + j;
        if (i < RP_TABLE) local += extra[i];
    }
    s_tot[t] = local;
    __syncthreads();
    if (t == 0) {
        u64 run = 0;
        for (u32 i = 0; i < 1024; ++i) { u64 v = s_tot[i]; s_tot[i] = run; run += v; }
    }
    __syncthreads();
    u64 run = s_tot[t];
    for (u32 j = 0; j < PER; ++j) {
        const u32 i = t * PER + j;
        if (i < RP_TABLE) { prefix[i] = run; run += extra[i]; }
    }
}

__global__ void rp_assemble_clean_kernel(const mcu_match* __restrict__ rows, u64 n, const u32* __restrict__ bitmap, const u64* __restrict__ prefix,
                                         mcu_match* __restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const mcu_match m = rows[i];
        const u32 bk = rp_bucket(m);
        if (!((bitmap[bk >> 5] >> (bk & 31)) & 1u)) out[i + prefix[bk]] = m;
    }
}

__global__ void rp_assemble_replayed_kernel(const mcu_match* __restrict__ rows, const u32* __restrict__ list, u32 nlist, const u64* __restrict__ vinfo,
                                            const u32* __restrict__ pool, const u64* __restrict__ prefix, mcu_match* __restrict__ out)
{
    const u32 k = blockIdx.x;
    if (k >= nlist) return;
    const u64 base = vinfo[3 * (u64)k], nv = vinfo[3 * (u64)k + 1], rb = vinfo[3 * (u64)k + 2];
    const u64 dst = rb + prefix[list[k]];
    for (u64 j = threadIdx.x; j < nv; j += blockDim.x) out[dst + j] = rows[pool[base + j]];
}

static int rp_grid(u64 n, int block)
{
    u64 want = div_up(n, (u64)block), cap = (u64)sm_count() * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

int replay_unclean(Session& s, const SeedParams* sp, bool can_replay, u64* unclean_buckets, u64* duplicate_rows)
{
    *unclean_buckets = 0;
    *duplicate_rows = 0;
    const u64 n = s.match_count;
    if (n < 2) return MCU_OK;
    cudaStream_t st = s.stream;
    MCU_TRY(s.rp_ctr.reserve(8 * sizeof(unsigned long long)));
    MCU_TRY(s.rp_bitmap.reserve(RP_WORDS * sizeof(u32)));
    unsigned long long* ctr = s.rp_ctr.as<unsigned long long>();
    MCU_CUDA(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), st));
    MCU_CUDA(cudaMemsetAsync(s.rp_bitmap.p, 0, RP_WORDS * sizeof(u32), st));
    const mcu_match* rows = s.matches.as<mcu_match>();
    rp_detect_kernel<<<rp_grid(n, 256), 256, 0, st>>>(rows, n, s.rp_bitmap.as<u32>(), ctr);
    s.launches++;
    MCU_CUDA(cudaMemcpyAsync(s.h_replay, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    const u32 nlist = (u32)s.h_replay[0];
    *unclean_buckets = nlist;
    if (nlist == 0 || !can_replay || !sp) return MCU_OK;

    // ---- rare path ----
    ExtendArgs ea;
    ea.g0 = s.packed[0].as<u32>(); ea.g1 = s.packed[1].as<u32>();
    ea.npos0 = s.n[0] >= (u64)sp->L ? s.n[0] - sp->L + 1 : 0;
    ea.npos1 = s.n[1] >= (u64)sp->L ? s.n[1] - sp->L + 1 : 0;
    ea.uniq = s.uniq.as<u32>();
    ea.cand = nullptr; ea.nfwd = ea.nrev = ea.cap = 0; ea.out = nullptr; ea.counters = nullptr;
    RpCollect rc;
    rc.rows = rows; rc.n = n; rc.bitmap = s.rp_bitmap.as<u32>();
    rc.canon = nullptr; rc.p0 = nullptr; rc.row = nullptr; rc.cap = 0; rc.ctr = ctr;
    const int wgrid = rp_grid(n * 32, 256);
    rp_collect_kernel<<<wgrid, 256, 0, st>>>(rc, ea, *sp, 1);  // count
    s.launches++;
    MCU_CUDA(cudaMemcpyAsync(s.h_replay, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    const u64 urows = s.h_replay[1], nseeds = s.h_replay[2];
    if (nseeds >= 0xFFFFFFFFull) { set_error("bucket replay: too many seeds"); return MCU_EINVAL; }
    MCU_TRY(s.rp_list.reserve((u64)nlist * 4 + 16));
    MCU_TRY(s.rp_canon.reserve((nseeds + 1) * 8));
    MCU_TRY(s.rp_keys_b.reserve((nseeds + 1) * 8));
    MCU_TRY(s.rp_bkeys.reserve((nseeds + 1) * 8));
    MCU_TRY(s.rp_idx_a.reserve((nseeds + 1) * 4));
    MCU_TRY(s.rp_idx_b.reserve((nseeds + 1) * 4));
    MCU_TRY(s.rp_p0.reserve((nseeds + 1) * 4));
    MCU_TRY(s.rp_row.reserve((nseeds + 1) * 4));
    MCU_TRY(s.rp_pool.reserve((urows + nseeds + 1) * 4));
    MCU_TRY(s.rp_extra.reserve(RP_TABLE * 4));
    MCU_TRY(s.rp_prefix.reserve(RP_TABLE * 8));
    MCU_TRY(s.rp_vinfo.reserve((u64)nlist * 24 + 24));
    MCU_CUDA(cudaMemsetAsync(ctr + 2, 0, sizeof(unsigned long long), st));
    MCU_CUDA(cudaMemsetAsync(s.rp_extra.p, 0, RP_TABLE * 4, st));
    rc.canon = s.rp_canon.as<u64>(); rc.p0 = s.rp_p0.as<u32>(); rc.row = s.rp_row.as<u32>(); rc.cap = nseeds;
    rp_collect_kernel<<<wgrid, 256, 0, st>>>(rc, ea, *sp, 0);
    rp_list_kernel<<<(RP_TABLE + 255) / 256, 256, 0, st>>>(s.rp_bitmap.as<u32>(), s.rp_list.as<u32>(), ctr + 6);
    s.launches += 2;

    // insertion order = ascending canonical mer; then group by bucket (stable)
    const u32* perm = nullptr;
    const u64* bkeys = nullptr;
    if (nseeds) {
        rp_iota_kernel<<<rp_grid(nseeds, 256), 256, 0, st>>>(s.rp_idx_a.as<u32>(), nseeds);
        s.launches++;
        bool in_a = true;
        u64 before = s.radix.launches;
        MCU_TRY(radix_sort_pairs<u64>(s.radix, s.rp_canon.as<u64>(), s.rp_idx_a.as<u32>(), s.rp_keys_b.as<u64>(), s.rp_idx_b.as<u32>(), nseeds,
                                      2 * sp->w, false, st, &in_a, nullptr));
        u32* idx1 = in_a ? s.rp_idx_a.as<u32>() : s.rp_idx_b.as<u32>();
        u32* idx2 = in_a ? s.rp_idx_b.as<u32>() : s.rp_idx_a.as<u32>();
        rp_bucket_keys_kernel<<<rp_grid(nseeds, 256), 256, 0, st>>>(rows, s.rp_row.as<u32>(), idx1, nseeds, s.rp_bkeys.as<u64>());
        s.launches++;
        MCU_TRY(radix_sort_pairs<u64>(s.radix, s.rp_bkeys.as<u64>(), idx1, s.rp_keys_b.as<u64>(), idx2, nseeds, 16, false, st, &in_a, nullptr));
        s.launches += s.radix.launches - before;
        perm = in_a ? idx1 : idx2;
        bkeys = in_a ? s.rp_bkeys.as<u64>() : s.rp_keys_b.as<u64>();
    }
    RpReplay rr;
    rr.rows = rows; rr.n = n; rr.list = s.rp_list.as<u32>(); rr.nlist = nlist;
    rr.bkeys = bkeys; rr.perm = perm; rr.p0 = s.rp_p0.as<u32>(); rr.row = s.rp_row.as<u32>(); rr.nseeds = nseeds;
    rr.pool = s.rp_pool.as<u32>(); rr.pool_cap = urows + nseeds;
    rr.extra = s.rp_extra.as<u32>(); rr.vinfo = s.rp_vinfo.as<u64>(); rr.ctr = ctr; rr.L = sp->L;
    rp_replay_kernel<<<(nlist + 63) / 64, 64, 0, st>>>(rr);
    rp_scan_kernel<<<1, 1024, 0, st>>>(s.rp_extra.as<u32>(), s.rp_prefix.as<u64>());
    s.launches += 2;
    MCU_CUDA(cudaMemcpyAsync(s.h_replay, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    if (s.h_replay[5] || s.h_replay[3] > urows + nseeds) { set_error("bucket replay: inconsistent bucket state"); return MCU_ECUDA; }
    const u64 extra = s.h_replay[4];
    MCU_TRY(s.rp_out.reserve((n + extra + 1) * sizeof(mcu_match)));
    rp_assemble_clean_kernel<<<rp_grid(n, 256), 256, 0, st>>>(rows, n, s.rp_bitmap.as<u32>(), s.rp_prefix.as<u64>(), s.rp_out.as<mcu_match>());
    rp_assemble_replayed_kernel<<<nlist, 128, 0, st>>>(rows, s.rp_list.as<u32>(), nlist, s.rp_vinfo.as<u64>(), s.rp_pool.as<u32>(),
                                                       s.rp_prefix.as<u64>(), s.rp_out.as<mcu_match>());
    s.launches += 2;
    MCU_CUDA(cudaGetLastError());
    std::swap(s.matches, s.rp_out);
    s.match_count = n + extra;
    *duplicate_rows = extra;
    return MCU_OK;
}

}  // namespace mcu
