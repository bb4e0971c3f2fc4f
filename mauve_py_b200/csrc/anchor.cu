// Anchoring pipeline kernels (sm_100a).  Reference semantics cited per kernel; LM = libMems/libMems.
//
// Data layout in HBM (per session):
//   ascii[g]   n_g bytes                      input genome g
//   packed[g]  ceil(n_g/16)+2 u32             2-bit, MSB-first per word, 2 zero pad words (SortedMerList::sequence)
//   keys/vals  (n_0-L+1)+(n_1-L+1) pairs x2   key = canon(2w bits)<<2 | genome<<1 | strand ; val = position
//   uniq       1 bit per genome-0 position    the seed starting here is unique in both genomes
//   pairs      u64 list p0 | p1<<32           unique seed pairs, forward-strand from the front / reverse from the back
//   cand       u64 list, same layout          pairs that may be the leftmost unique seed of their match
//   matches    mcu_match rows                 raw, then ordered into the reference list order
#include "anchor.cuh"

namespace mcu {

// =========================================================================================
// pack2bit: SortedMerList::translate32 (LM/SortedMerList.cpp:425-460) with the BasicDNATable
// (:29-47): A=0, C/B/Y=1, G/S/K=2, T=3, anything else 0; '-' is an error (:433-437).
// One thread -> one 32-bit word (16 bases), 128-bit coalesced loads.  HBM-bound: 1 B read +
// 0.25 B written per base.
// =========================================================================================
__device__ __forceinline__ u32 base_code(u32 c, u32& gap)
{
    // 2-bit code per letter index (c & 31): C=3,B=2,Y=25 -> 1 ; G=7,S=19,K=11 -> 2 ; T=20 -> 3
    const u64 TABLE = (1ull << 6) | (1ull << 4) | (1ull << 50) | (2ull << 14) | (2ull << 38) | (2ull << 22) | (3ull << 40);
    gap |= (c == (u32)'-');
    u32 u = c & 0xDFu;
    return (u - 65u) < 26u ? (u32)((TABLE >> (2 * (c & 31u))) & 3ull) : 0u;
}

__global__ void __launch_bounds__(256) pack_kernel(const u8* __restrict__ seq, u64 n, u32* __restrict__ packed, u64 word_lo, u64 word_hi, u32* __restrict__ err)
{
    u32 gap = 0;
    for (u64 w = word_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; w < word_hi; w += (u64)gridDim.x * blockDim.x) {
        u64 base = w * 16;
        u32 v = 0;
        if (base + 16 <= n) {
            uint4 q = __ldg((const uint4*)(seq + base));
            u32 qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    u32 c = (qq[k] >> (8 * b)) & 255u;
                    v |= base_code(c, gap) << (30 - 2 * (4 * k + b));
                }
            }
        } else if (base < n) {
            for (u32 j = 0; base + j < n; ++j) v |= base_code(seq[base + j], gap) << (30 - 2 * j);
        }
        packed[w] = v;
    }
    if (gap) atomicOr(err, 1u);
}

// =========================================================================================
// seedgen: for every position p in [0, n-L+1): canonical spaced seed + strand flag =
// SortedMerList::GetSeedMer (:726-762), RevCompMer (:597-614), GetDnaSeedMer (:764-769),
// FillDnaSeedSML (:771-783) / FillDnaSML (:617-723, same values for solid seeds).
// Emits key = canon<<2 | genome<<1 | strand (order-isomorphic to bmer::mer per genome, with the
// genome bit placed so that one sort of both genomes also groups the join key `canon`), and
// accumulates the digit histograms of all radix passes on the fly (so the sort never re-reads
// the keys for counting).  Coalesced: thread t of a block handles position base+t.
// HBM traffic: 0.25 B/base packed read (L1/L2 served) + sizeof(K)+4 written.
// =========================================================================================
// Sharded sorted-mer-list build: rank r owns the seeds whose key lies in the r-th of n RANGES of the key space, so
// that the ranks' sorted lists, one after the other, are the sorted list.  The key is the canonical mer = min(forward, reverse
// complement), whose density over the key range is ~2 (1 - x) rather than flat: the splitters sit at the quantiles
// 1 - sqrt(1 - k / n) of that density (top 16 key bits), which balances random sequence to a few percent.
struct SmlSplit {
    u32 n = 0;       // 0: ownership by hash (match finding)
    int shift = 0;   // key >> shift = the bits the splitters are compared with
    u32 s[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

template <typename K>
__global__ void __launch_bounds__(256) seedgen_kernel(const u32* __restrict__ packed, u64 npos, SeedParams sp, u32 genome,
                                                     K* __restrict__ keys, u32* __restrict__ vals, u64 out_base,
                                                     u64* __restrict__ hist, int passes, u32 shard, u32 nshard,
                                                     unsigned long long* __restrict__ out_counter, SmlSplit split = SmlSplit())
{
    extern __shared__ u32 sh_hist[];  // passes*256
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const bool sharded = nshard > 1;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (npos + stride - 1) / stride;
    for (u64 r = 0; r < rounds; ++r) {
        u64 p = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool live = p < npos;
        u64 key = 0;
        if (live) {
            u64 f = extract_seed(load_mer32(packed, p), sp);
            u64 rc = revcomp_seed(f, sp.w);
            u32 strand = rc < f;  // GetDnaSeedMer: forward wins ties (f < rc|1)
            u64 canon = strand ? rc : f;
            key = (canon << 2) | (genome << 1) | strand;
            if (split.n) {   // sorted-mer-list shards: ranges of the key, so that the ranks' sorted lists concatenate
                const u32 top = (u32)(key >> split.shift);
                u32 owner = 0;
                for (u32 k = 0; k + 1 < split.n; ++k) owner += top >= split.s[k];
                live = owner == shard;
            } else
                live = seed_owned(f, rc, shard, nshard);
        }
        u64 slot = out_base + p;
        if (sharded) {  // unordered compaction (tie order is irrelevant for match finding)
            u32 m = __ballot_sync(0xffffffffu, live);
            u64 wbase = 0;
            if (lane == 0 && m) wbase = atomicAdd(out_counter, (unsigned long long)__popc(m));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            slot = out_base + wbase + __popc(m & lanemask_lt());
        }
        if (live) {
            keys[slot] = (K)key;
            vals[slot] = (u32)p;
            for (int q = 0; q < passes; ++q) atomicAdd(&sh_hist[q * 256 + (u32)((key >> (8 * q)) & 255)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x)
        if (sh_hist[i]) atomicAdd((unsigned long long*)&hist[i], (unsigned long long)sh_hist[i]);
}

// keys -> bmer::mer values (canonical seed left-aligned in 64 bits | strand bit) for MemorySML::Read
template <typename K>
__global__ void keys_to_mers_kernel(const K* __restrict__ keys, u64 n, int w, u64* __restrict__ mers)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 k = keys[i];
        mers[i] = ((k >> 2) << (64 - 2 * w)) | (k & 1);
    }
}

// =========================================================================================
// join: MatchFinder::SearchRange (LM/MatchFinder.cpp:172-340) merges the sorted lists on
// mer & seed_mask and hands every equal-key run to EnumerateMatches; for two genomes both
// PairwiseMatchFinder::EnumerateMatches (LM/PairwiseMatchFinder.cpp:37-71) and MemHash's
// (LM/MemHash.cpp:139-162, tolerances 0/1) keep exactly the runs with one occurrence in each
// genome.  In the combined sorted array such a run is "g0 entry immediately followed by a g1
// entry with the same canon, different canon on both sides" -- a purely local test.
// Output = HashMatch + SetDirection (LM/MemHash.cpp:167-203) as
//   uniq   bitmap over genome-0 positions: the seed starting here is unique in both genomes
//          (12.5 MB at 100 Mbp: L2 resident, set with fire-and-forget atomics)
//   pairs  u64 list p0 | p1 << 32: forward-strand pairs are filled from the front, reverse-strand
//          pairs from the back, so the strand needs no storage (one reservation per block and strand).
// HBM traffic: sizeof(K) per sorted entry + 8 B per seed pair read (vals) + 8 B written.
// counters: [0] forward pairs, [1] repeat-limit flag, [6] reverse pairs.
// =========================================================================================
template <typename K>
__global__ void __launch_bounds__(256) join_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, u64 n, u32* __restrict__ uniq,
                                                  u64* __restrict__ pairs, u64 pair_cap, unsigned long long* __restrict__ counters)
{
    constexpr int IPT = 8;
    __shared__ u32 s_warp[2][8];
    __shared__ unsigned long long s_base[2];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 tile = (u64)blockIdx.x * (256 * IPT);
    u32 repeat = 0, fmask = 0, rmask = 0;
    u64 ent[IPT];
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
        const u64 i = tile + (u64)it * 256 + threadIdx.x;
        ent[it] = 0;
        if (i < n) {
            const K k = keys[i];
            const K canon = k >> 2;
            const bool next2 = i + 2 < n && (keys[i + 2] >> 2) == canon;
            // a run longer than MER_REPEAT_LIMIT contains an entry whose +2 and +1000 neighbours are both in it
            if (next2 && i + 1000 < n && (keys[i + 1000] >> 2) == canon) repeat = 1;
            if (!((k >> 1) & 1) && i + 1 < n && !next2) {  // a pair starts at its genome-0 entry
                const K k1 = keys[i + 1];
                if ((k1 >> 2) == canon && ((k1 >> 1) & 1) && !(i > 0 && (keys[i - 1] >> 2) == canon)) {
                    const u32 p0 = vals[i], p1 = vals[i + 1];
                    ent[it] = (u64)p0 | ((u64)p1 << 32);
                    if ((k ^ k1) & 1) rmask |= 1u << it; else fmask |= 1u << it;
                    atomicOr(&uniq[p0 >> 5], 1u << (p0 & 31));
                }
            }
        }
    }
    // block-wide exclusive scan of the per-thread counts: one reservation per block and strand
    const u32 cf = __popc(fmask), cr = __popc(rmask);
    u32 inf = cf, inr = cr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 tf = __shfl_up_sync(0xffffffffu, inf, o), tr = __shfl_up_sync(0xffffffffu, inr, o);
        if (lane >= (u32)o) { inf += tf; inr += tr; }
    }
    if (lane == 31) { s_warp[0][warp] = inf; s_warp[1][warp] = inr; }
    __syncthreads();
    u32 wf = 0, wr = 0, totf = 0, totr = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if ((u32)w < warp) { wf += s_warp[0][w]; wr += s_warp[1][w]; }
        totf += s_warp[0][w];
        totr += s_warp[1][w];
    }
    if (threadIdx.x == 0) {
        s_base[0] = totf ? atomicAdd(&counters[0], (unsigned long long)totf) : 0ull;
        s_base[1] = totr ? atomicAdd(&counters[6], (unsigned long long)totr) : 0ull;
    }
    __syncthreads();
    u64 of = s_base[0] + wf + inf - cf, orv = s_base[1] + wr + inr - cr;
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
        if (fmask & (1u << it)) pairs[of++] = ent[it];
        if (rmask & (1u << it)) pairs[pair_cap - 1 - (orv++)] = ent[it];
    }
    if (__any_sync(0xffffffffu, repeat) && lane == 0) atomicMax(&counters[1], 1ull);
}

// candidate filter: a seed pair whose left neighbour on the same diagonal is also a unique seed
// pair cannot be the leftmost unique seed of its match.  This removes the bulk of AddHashEntry's
// "already contained" rejections (LM/MemHash.cpp:215-220): 2.7 M seed pairs -> ~30 k candidates
// on MDS42.  counters[2] / [7] = forward / reverse candidates.
__global__ void __launch_bounds__(256) candidate_kernel(ExtendArgs a, SeedParams sp, const u64* __restrict__ pairs, u64 pfwd, u64 prev_, u64 pair_cap, bool solid,
                                                       u64 agree_fwd, u64 agree_rev)
{
    const u32 lane = threadIdx.x & 31;
    const u64 total = pfwd + prev_;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (total + stride - 1) / stride;
    for (u64 r = 0; r < rounds; ++r) {
        const u64 idx = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool is_cand = false, rev = false;
        u64 e = 0;
        if (idx < total) {
            rev = idx >= pfwd;
            e = rev ? pairs[pair_cap - 1 - (idx - pfwd)] : pairs[idx];
            const i64 p0 = (i64)(e & 0xffffffffu), p1 = (i64)(e >> 32);
            const i64 d = rev ? p0 + p1 : p1 - p0;
            i64 other;
            if (rev ? idx - pfwd < agree_rev : idx < agree_fwd) {
                // the bucket kernel already compared the neighbour bases (p0 > 0 and the partner exists): only the bitmap is left
                is_cand = !uniq_bit(a.uniq, p0 - 1);
            } else if (solid) {
                // hit(p0) is known: the left neighbour is a hit iff the one new base agrees
                bool left_hit = false;
                if (p0 > 0 && uniq_bit(a.uniq, p0 - 1)) {
                    if (!rev) left_hit = p1 > 0 && base_at(a.g0, p0 - 1) == base_at(a.g1, p1 - 1);
                    else left_hit = p1 + 1 < (i64)a.npos1 && base_at(a.g0, p0 - 1) == 3u - base_at(a.g1, p1 + sp.w);
                }
                is_cand = !left_hit;
            } else
                is_cand = !(p0 > 0 && uniq_bit(a.uniq, p0 - 1) && probe_hit(a, sp, rev, d, p0 - 1, other));
        }
        const u32 mf = __ballot_sync(0xffffffffu, is_cand && !rev), mr = __ballot_sync(0xffffffffu, is_cand && rev);
        if (mf | mr) {
            u64 bf = 0, br = 0;
            if (lane == 0) {
                if (mf) bf = atomicAdd(&a.counters[2], (unsigned long long)__popc(mf));
                if (mr) br = atomicAdd(&a.counters[7], (unsigned long long)__popc(mr));
            }
            bf = __shfl_sync(0xffffffffu, bf, 0);
            br = __shfl_sync(0xffffffffu, br, 0);
            if (is_cand) {
                if (!rev) a.cand[bf + __popc(mf & lanemask_lt())] = e;
                else a.cand[a.cap - 1 - (br + __popc(mr & lanemask_lt()))] = e;
            }
        }
    }
}

// =========================================================================================
// extend: MatchFinder::ExtendMatch (LM/MatchFinder.h:218-374) + the containment dedupe of
// MemHash::AddHashEntry (LM/MemHash.cpp:209-251), in closed form:
// on the seed's diagonal a "hit" at offset t is spaced-seed equality of the two genomes with
// the strand relation of the seed (forward: f0 == f1; reverse: f0 == revcomp(f1) and f0 not
// its own reverse complement -- the parity test of :281-303); the match is the maximal chain
// of hits with gaps <= L containing the seed, clipped to valid seed positions (:248-257).
// A hit at a genome-0 position whose uniq bit is set IS a unique seed pair of this diagonal
// (its only genome-1 occurrence is the one the hit compares against).
// =========================================================================================
// ---- solid seeds of odd weight (the default for genomes > ~40 Mbp: LM/SeedMasks.h has no spaced pattern in the
// CODING_SEED slot above weight 16) ---------------------------------------------------------------------------
// With a solid pattern a hit at t means bases [t, t+w) agree on the diagonal, two consecutive hits are at most one
// position apart, and a single mismatching base separates hits by w+1 > L: a match is exactly a maximal run of
// agreeing bases.  (Odd w: a solid mer is never its own reverse complement, so the parity test of
// MatchFinder.h:281-303 never fires.)  Runs are measured 32 bases per XOR instead of probing seed by seed.

// reverse complement of 32 packed bases
__device__ __forceinline__ u64 revcomp32(u64 v)
{
    u64 x = __brevll(~v);
    return ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

struct SolidDiag {
    const u32* g0;
    const u32* g1;
    i64 n0, n1;   // sequence lengths in bases
    bool rev;
    i64 d;        // forward: g1 index = g0 index + d ; reverse: g1 index = d - g0 index
    // do base x of genome 0 and its partner base agree?  (both indices must be in range)
    __device__ __forceinline__ bool agree(i64 x) const
    {
        const u32 a = base_at(g0, x);
        return rev ? a == 3u - base_at(g1, d - x) : a == base_at(g1, x + d);
    }
    // XOR of the 32 bases [x, x+32) of genome 0 with their partners; caller guarantees both blocks are in range
    __device__ __forceinline__ u64 diff32(i64 x) const
    {
        const u64 a = load_mer32(g0, (u64)x);
        const u64 b = rev ? revcomp32(load_mer32(g1, (u64)(d - x - 31))) : load_mer32(g1, (u64)(x + d));
        return a ^ b;
    }
};

// number of agreeing bases going right from base x0 (inclusive), at most `maxlen`
__device__ __forceinline__ i64 solid_run_right(const SolidDiag& D, i64 x0, i64 maxlen, u32 lane)
{
    i64 done = 0;
    while (done + 32 <= maxlen) {  // whole blocks, 32 lanes x 32 bases per round
        const i64 blk = done + 32 * (i64)lane;
        u64 x = 0;
        const bool in = blk + 32 <= maxlen;
        if (in) x = D.diff32(x0 + blk);
        const u32 bad = __ballot_sync(0xffffffffu, in && x != 0), live = __ballot_sync(0xffffffffu, in);
        if (bad) {
            const int l = __ffs(bad) - 1;
            const u64 xl = __shfl_sync(0xffffffffu, x, l);
            return done + 32 * (i64)l + (__clzll(xl) >> 1);
        }
        done += 32 * (i64)__popc(live);
    }
    while (done < maxlen && D.agree(x0 + done)) ++done;  // tail shorter than a block (uniform across the warp)
    return done;
}

// number of agreeing bases going left from base x0 (inclusive), at most `maxlen`
__device__ __forceinline__ i64 solid_run_left(const SolidDiag& D, i64 x0, i64 maxlen, u32 lane)
{
    i64 done = 0;
    while (done + 32 <= maxlen) {
        const i64 blk = done + 32 * (i64)lane;
        u64 x = 0;
        const bool in = blk + 32 <= maxlen;
        if (in) x = D.diff32(x0 - blk - 31);
        const u32 bad = __ballot_sync(0xffffffffu, in && x != 0), live = __ballot_sync(0xffffffffu, in);
        if (bad) {
            const int l = __ffs(bad) - 1;
            const u64 xl = __shfl_sync(0xffffffffu, x, l);
            return done + 32 * (i64)l + (__ffsll((long long)xl) - 1) / 2;
        }
        done += 32 * (i64)__popc(live);
    }
    while (done < maxlen && D.agree(x0 - done)) ++done;
    return done;
}

__global__ void __launch_bounds__(256) extend_solid_kernel(ExtendArgs a, SeedParams sp, i64 n0, i64 n1)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp_global = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const i64 w = sp.w;
    const u64 total = a.nfwd + a.nrev;
    for (u64 c = warp_global; c < total; c += nwarps) {
        SolidDiag D;
        D.g0 = a.g0; D.g1 = a.g1; D.n0 = n0; D.n1 = n1;
        D.rev = c >= a.nfwd;
        const u64 e = D.rev ? a.cand[a.cap - 1 - (c - a.nfwd)] : a.cand[c];
        const i64 t0 = (i64)(e & 0xffffffffu), p1 = (i64)(e >> 32);
        // base x of genome 0 pairs with base x + (p1 - t0) (forward) or (t0 + p1 + w - 1) - x (reverse) of genome 1
        D.d = D.rev ? t0 + p1 + w - 1 : p1 - t0;
        // ---- left: bases t0-1, t0-2, ... ----
        const i64 maxl = D.rev ? min(t0, n1 - 1 - (D.d - t0)) : min(t0, t0 + D.d);
        const i64 left = maxl > 0 ? solid_run_left(D, t0 - 1, maxl, lane) : 0;
        const i64 lo = t0 - left;
        // another unique seed pair at a seed position in [lo, t0): not the leftmost seed of this match
        bool other = false;
        for (i64 wd = (lo >> 5) + lane; wd <= ((t0 - 1) >> 5) && left > 0; wd += 32) {
            u32 bits = __ldg(a.uniq + wd);
            if (wd == (lo >> 5)) bits &= 0xffffffffu << (lo & 31);
            if (wd == ((t0 - 1) >> 5)) bits &= 0xffffffffu >> (31 - ((t0 - 1) & 31));
            other |= bits != 0;
        }
        if (__any_sync(0xffffffffu, other)) continue;
        // ---- right: bases t0+w, t0+w+1, ... ----
        const i64 xr = t0 + w;
        const i64 maxr = D.rev ? min(n0 - xr, D.d - xr + 1) : min(n0 - xr, n1 - (xr + D.d));
        const i64 right = maxr > 0 ? solid_run_right(D, xr, maxr, lane) : 0;
        if (lane == 0) {
            const u64 slot = atomicAdd(&a.counters[3], 1ull);
            mcu_match m;
            m.len = left + w + right;
            m.start0 = lo + 1;
            // forward: partner of base lo; reverse: the match covers genome-1 bases [d - (lo+len-1), d - lo], reported negative
            m.start1 = D.rev ? -((D.d - (lo + m.len - 1)) + 1) : lo + D.d + 1;
            a.out[slot] = m;
        }
    }
}

// One warp per candidate: lanes probe the L offsets beyond the current end, ballot, jump to
// the farthest hit.  While walking left, meeting another unique seed pair of the same diagonal
// means this candidate is not the leftmost one of its match -> abandon (exactly one emitter
// per match).  counters[3] = number of matches.
__global__ void __launch_bounds__(256) extend_kernel(ExtendArgs a, SeedParams sp)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp_global = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const i64 L = sp.L;
    const u64 total = a.nfwd + a.nrev;
    for (u64 c = warp_global; c < total; c += nwarps) {
        const bool rev = c >= a.nfwd;
        const u64 e = rev ? a.cand[a.cap - 1 - (c - a.nfwd)] : a.cand[c];
        const i64 t0 = (i64)(e & 0xffffffffu), p1 = (i64)(e >> 32);
        const i64 d = rev ? t0 + p1 : p1 - t0;
        // ---- walk left ----
        i64 cur = t0;
        bool abandoned = false;
        while (true) {
            i64 t = cur - 1 - (i64)lane, other = 0;
            bool h = (i64)lane < L && probe_hit(a, sp, rev, d, t, other);
            bool uq = h && uniq_bit(a.uniq, t);
            if (__any_sync(0xffffffffu, uq)) { abandoned = true; break; }
            u32 hits = __ballot_sync(0xffffffffu, h);
            if (!hits) break;
            cur -= 32 - __clz(hits);  // farthest hit: highest lane = largest distance
        }
        if (abandoned) continue;
        const i64 lo = cur;
        // ---- walk right ----
        cur = t0;
        while (true) {
            i64 t = cur + 1 + (i64)lane, other = 0;
            bool h = (i64)lane < L && probe_hit(a, sp, rev, d, t, other);
            u32 hits = __ballot_sync(0xffffffffu, h);
            if (!hits) break;
            cur += 32 - __clz(hits);
        }
        const i64 hi = cur;
        if (lane == 0) {
            u64 slot = atomicAdd(&a.counters[3], 1ull);
            mcu_match m;
            m.len = hi - lo + L;
            m.start0 = lo + 1;
            m.start1 = rev ? -((d - hi) + 1) : lo + d + 1;
            a.out[slot] = m;
        }
    }
}

// =========================================================================================
// order: MemHash::GetMatchList (LM/MemHash.h:183-203) walks buckets 0..39999
// (bucket = generalized offset mod 40000, LM/MemHash.cpp:213, offset per
// MatchHashEntry::CalculateOffset LM/MatchHashEntry.cpp:141-160); inside a bucket entries are
// kept sorted by strict_start_lessthan_ptr (LM/MatchHashEntry.cpp:48-67; stored entries have
// m_mersize == 0 because the copy goes through operator=, :112-120).
// Two stable LSD sorts: secondary key (genome-1 start as compared there), then bucket|start0.
// =========================================================================================
__global__ void order_keys_kernel(const mcu_match* __restrict__ rows, u64 n, int s0_bits, u64* __restrict__ primary, u64* __restrict__ secondary,
                                  u32* __restrict__ idx)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        mcu_match m = rows[i];
        i64 off = m.start1 - m.start0 - (m.start1 < 0 ? m.len : 0);
        i64 b = ((off % 40000) + 40000) % 40000;
        u64 s1 = m.start1 < 0 ? (u64)(-m.start1 + m.len) : (u64)m.start1;
        primary[i] = ((u64)b << s0_bits) | (u64)m.start0;
        secondary[i] = (s1 << 1) | (m.start1 < 0 ? 1u : 0u);
        idx[i] = (u32)i;
    }
}

__global__ void gather_u64_kernel(const u64* __restrict__ src, const u32* __restrict__ idx, u64 n, u64* __restrict__ dst)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void gather_rows_kernel(const mcu_match* __restrict__ src, const u32* __restrict__ idx, u64 n, mcu_match* __restrict__ dst)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

// =========================================================================================
// host side
// =========================================================================================
static int grid_for(u64 n, int block, int per_sm)
{
    u64 want = div_up(n, (u64)block);
    u64 cap = (u64)sm_count() * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

static int bit_length(u64 x)
{
    int b = 0;
    while (x) { ++b; x >>= 1; }
    return b;
}

int session_init(Session& s)
{
    MCU_TRY(ensure_device());
    MCU_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) MCU_CUDA(cudaEventCreate(&s.ev[i]));
    for (int i = 0; i < 12; ++i) MCU_CUDA(cudaEventCreate(&s.kev[i]));
    MCU_CUDA(cudaEventCreateWithFlags(&s.ev_aux, cudaEventDisableTiming));
    MCU_CUDA(cudaHostAlloc((void**)&s.h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    MCU_TRY(s.counters.reserve(8 * sizeof(unsigned long long)));
    MCU_CUDA(cudaHostAlloc((void**)&s.h_replay, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    s.use_buckets = getenv("MAUVE_CUDA_SORT_PATH") == nullptr;
    s.ok = true;
    return MCU_OK;
}

void session_destroy(Session& s)
{
    if (!s.ok) return;
    cudaStreamSynchronize(s.stream);
    DevBuf* bufs[] = {&s.ascii[0], &s.ascii[1], &s.packed[0], &s.packed[1], &s.keys_a, &s.keys_b, &s.vals_a, &s.vals_b,
                      &s.uniq, &s.pairs, &s.cand, &s.raw_matches, &s.ord_keys_a, &s.ord_keys_b, &s.ord_vals_a, &s.ord_vals_b,
                      &s.matches, &s.ord_primary, &s.counters, &s.radix.hist, &s.radix.status, &s.radix.counters,
                      &s.rp_ctr, &s.rp_bitmap, &s.rp_list, &s.rp_canon, &s.rp_keys_b, &s.rp_idx_a, &s.rp_idx_b, &s.rp_p0, &s.rp_row, &s.rp_bkeys,
                      &s.rp_pool, &s.rp_extra, &s.rp_prefix, &s.rp_vinfo, &s.rp_out,
                      &s.bk_a, &s.bk_b, &s.bk_tab1, &s.bk_tab2, &s.bk_spill, &s.bk_tileseg, &s.bt_tab, &s.bt_seg_a, &s.bt_seg_b,
                      &s.sol_raw, &s.sol_freq[0], &s.sol_freq[1], &s.as_rows, &s.as_match, &s.as_off, &s.as_lcb, &s.gathered, &s.comm_small};
    if (s.copy_stream) { cudaStreamSynchronize(s.copy_stream); cudaStreamDestroy(s.copy_stream); s.copy_stream = nullptr; }
    for (cudaEvent_t e : s.up_events) cudaEventDestroy(e);
    s.up_events.clear();
    if (s.ev_aux) cudaEventDestroy(s.ev_aux);
    if (s.h_comm) cudaFreeHost(s.h_comm);
    s.h_comm = nullptr;
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 8; ++i) cudaEventDestroy(s.ev[i]);
    for (int i = 0; i < 12; ++i) cudaEventDestroy(s.kev[i]);
    if (s.h_counters) cudaFreeHost(s.h_counters);
    if (s.h_replay) cudaFreeHost(s.h_replay);
    cudaStreamDestroy(s.stream);
    s.ok = false;
}

int session_upload(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1)
{
    const char* seq[2] = {seq0, seq1};
    u64 n[2] = {n0, n1};
    s.up_chunks = 0;
    for (int g = 0; g < 2; ++g) {
        if (n[g] && !seq[g]) { set_error("session_upload: NULL sequence"); return MCU_EINVAL; }
        if (n[g] >= 0xFFFFFFFFull) { set_error("sequence longer than the reference's 32-bit position limit"); return MCU_EINVAL; }
        MCU_TRY(s.ascii[g].reserve(n[g] + 16));
        if (n[g]) MCU_CUDA(cudaMemcpyAsync(s.ascii[g].p, seq[g], n[g], cudaMemcpyHostToDevice, s.stream));
        s.n[g] = n[g];
    }
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    return MCU_OK;
}

// packs words [w0, w1) of genome g (words past the end of the sequence become the zero padding)
int session_upload_slice(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1, int rank, int world)
{
    const char* seq[2] = {seq0, seq1};
    u64 n[2] = {n0, n1};
    s.up_chunks = 0;
    for (int g = 0; g < 2; ++g) {
        if (n[g] && !seq[g]) { set_error("session_upload_slice: NULL sequence"); return MCU_EINVAL; }
        if (n[g] >= 0xFFFFFFFFull) { set_error("sequence longer than the reference's 32-bit position limit"); return MCU_EINVAL; }
        MCU_TRY(s.ascii[g].reserve(n[g] + 16));
        const u64 chunk = pack_chunk_words(n[g], world);
        const u64 b0 = chunk * 16 * (u64)rank, b1 = chunk * 16 * (u64)(rank + 1);
        const u64 lo = b0 < n[g] ? b0 : n[g], hi = b1 < n[g] ? b1 : n[g];
        if (hi > lo) MCU_CUDA(cudaMemcpyAsync(s.ascii[g].as<char>() + lo, seq[g] + lo, hi - lo, cudaMemcpyHostToDevice, s.stream));
        s.n[g] = n[g];
    }
    return MCU_OK;
}

int session_upload_begin(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1, int chunks)
{
    const char* seq[2] = {seq0, seq1};
    u64 n[2] = {n0, n1};
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    if (!s.copy_stream) MCU_CUDA(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    while (s.up_events.size() < (size_t)2 * chunks) {
        cudaEvent_t e;
        MCU_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s.up_events.push_back(e);
    }
    // the copies must not overtake work of an earlier run that still reads the buffers
    MCU_CUDA(cudaEventRecord(s.ev_aux, s.stream));
    MCU_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev_aux, 0));
    for (int g = 0; g < 2; ++g) {
        if (n[g] && !seq[g]) { set_error("session_upload_begin: NULL sequence"); return MCU_EINVAL; }
        if (n[g] >= 0xFFFFFFFFull) { set_error("sequence longer than the reference's 32-bit position limit"); return MCU_EINVAL; }
        MCU_TRY(s.ascii[g].reserve(n[g] + 16));
        s.n[g] = n[g];
        const u64 cb = (div_up(div_up(n[g], (u64)chunks), 4096) * 4096);  // whole scatter tiles per piece
        s.up_chunk_bases[g] = cb ? cb : 4096;
    }
    s.up_chunks = chunks;
    for (int g = 0; g < 2; ++g)
        for (int c = 0; c < chunks; ++c) {
            const u64 cb = s.up_chunk_bases[g];
            const u64 b0 = (u64)c * cb < n[g] ? (u64)c * cb : n[g];
            const u64 b1 = c == chunks - 1 ? n[g] : ((u64)(c + 1) * cb < n[g] ? (u64)(c + 1) * cb : n[g]);
            if (b1 > b0) MCU_CUDA(cudaMemcpyAsync(s.ascii[g].as<char>() + b0, seq[g] + b0, b1 - b0, cudaMemcpyHostToDevice, s.copy_stream));
            MCU_CUDA(cudaEventRecord(s.up_events[(size_t)g * chunks + c], s.copy_stream));
        }
    return MCU_OK;
}

static int run_pack_words(Session& s, int g, u64 w0, u64 w1, u32* err_flag)
{
    if (w1 <= w0) return MCU_OK;
    pack_kernel<<<grid_for(w1 - w0, 256, 8), 256, 0, s.stream>>>(s.ascii[g].as<u8>(), s.n[g], s.packed[g].as<u32>(), w0, w1, err_flag);
    s.launches++;
    return MCU_OK;
}

static int run_pack(Session& s, int g, u32* err_flag)
{
    const u64 words = div_up(s.n[g], 16) + 2;
    if (s.pack_world > 1) {  // this rank's chunk; the other chunks arrive through after_pack (NCCL all-gather, comm.cu)
        const u64 chunk = pack_chunk_words(s.n[g], s.pack_world);
        MCU_TRY(s.packed[g].reserve(chunk * (u64)s.pack_world * sizeof(u32)));
        return run_pack_words(s, g, chunk * (u64)s.pack_rank, chunk * (u64)(s.pack_rank + 1), err_flag);
    }
    MCU_TRY(s.packed[g].reserve(words * sizeof(u32)));
    return run_pack_words(s, g, 0, words, err_flag);
}

int run_pack_genome(Session& s, int g, u32* err_flag) { return run_pack(s, g, err_flag); }

// chunked upload: makes the session stream wait for piece c of genome g and packs the words that piece completes.
// Returns through *bases_ready how many leading bases of the genome are packed afterwards.
int run_pack_piece(Session& s, int g, int c, u32* err_flag, u64* bases_ready)
{
    const u64 cb = s.up_chunk_bases[g], n = s.n[g];
    const u64 words = div_up(n, 16) + 2;
    MCU_CUDA(cudaStreamWaitEvent(s.stream, s.up_events[(size_t)g * s.up_chunks + c], 0));
    const bool last = c == s.up_chunks - 1;
    const u64 b0 = (u64)c * cb, b1 = last ? n : (u64)(c + 1) * cb;
    const u64 w0 = b0 / 16 < words ? b0 / 16 : words, w1 = last ? words : (b1 / 16 < words ? b1 / 16 : words);
    if (c == 0) MCU_TRY(s.packed[g].reserve(words * sizeof(u32)));
    MCU_TRY(run_pack_words(s, g, w0, w1, err_flag));
    *bases_ready = last ? n : (b1 < n ? b1 : n);
    return MCU_OK;
}

// whatever is still in flight of a chunked upload: wait for it and pack everything (paths that do not overlap)
static int finish_chunked_upload(Session& s, u32* err_flag)
{
    for (int g = 0; g < 2; ++g) {
        u64 ready = 0;
        for (int c = 0; c < s.up_chunks; ++c) MCU_TRY(run_pack_piece(s, g, c, err_flag, &ready));
    }
    s.up_chunks = 0;
    return MCU_OK;
}

// Phase 1 of a run: pack + enumeration of the unique seed pairs of this rank's key range (uniq bitmap, pair list).
template <typename K>
static int run_enumerate(Session& s, const SeedParams& sp, int shard_index, int shard_count)
{
    const u64 npos0 = s.n[0] >= (u64)sp.L ? s.n[0] - sp.L + 1 : 0;
    const u64 npos1 = s.n[1] >= (u64)sp.L ? s.n[1] - sp.L + 1 : 0;
    const u64 ntot = npos0 + npos1;
    const int key_bits = 2 * sp.w + 2;
    const int passes = (key_bits + 7) / 8;
    const bool sharded = shard_count > 1;
    unsigned long long* ctr = s.counters.as<unsigned long long>();
    // counters: [0] fwd pairs [1] repeat flag [2] fwd candidates [3] matches [4] gap error (u32) [5] shard element count [6] rev pairs [7] rev candidates
    MCU_CUDA(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), s.stream));
    MCU_CUDA(cudaEventRecord(s.ev[0], s.stream));

    // ---- pack ----  (a chunked upload in flight: the bucketed path packs piece by piece under the copies, bucket.cu)
    const bool chunked = s.up_chunks > 0 && s.pack_world == 1;
    const bool overlap = chunked && s.use_buckets && bucket_plan_applies(sp, npos0, npos1, shard_count);
    if (chunked && !overlap) MCU_TRY(finish_chunked_upload(s, (u32*)(ctr + 4)));
    else if (!chunked) {
        for (int g = 0; g < 2; ++g) MCU_TRY(run_pack(s, g, (u32*)(ctr + 4)));
        if (s.after_pack) MCU_TRY(s.after_pack(s));
    }
    MCU_CUDA(cudaEventRecord(s.ev[1], s.stream));

    // ---- enumerate unique seed pairs: uniq bitmap + pair list ----
    const u64 pair_cap = (npos0 < npos1 ? npos0 : npos1) + 1;
    const u64 uniq_words = div_up(npos0 + 1, 32) + 1;
    MCU_TRY(s.uniq.reserve(uniq_words * sizeof(u32)));
    MCU_TRY(s.pairs.reserve(pair_cap * sizeof(u64)));
    MCU_TRY(s.cand.reserve(pair_cap * sizeof(u64)));
    MCU_CUDA(cudaMemsetAsync(s.uniq.p, 0, uniq_words * sizeof(u32), s.stream));
    bool bucketed = false;
    s.bk_direct = 0;
    s.bk_aux = false;
    u64 nsort = ntot;
    int passes_run = 0;
    s.bk_spilled = 0;
    if (s.use_buckets) MCU_TRY(bucket_group(s, sp, shard_index, shard_count, pair_cap, s.ev[2], s.ev[3], &bucketed, &nsort));
    if (s.up_chunks > 0 && s.pack_world == 1) MCU_TRY(finish_chunked_upload(s, (u32*)(ctr + 4)));  // no-op when the bucketed path consumed the pieces
    if (!bucketed) {
        // ---- seedgen ----
        MCU_TRY(s.keys_a.reserve((ntot + 1) * sizeof(K)));
        MCU_TRY(s.keys_b.reserve((ntot + 1) * sizeof(K)));
        MCU_TRY(s.vals_a.reserve((ntot + 1) * sizeof(u32)));
        MCU_TRY(s.vals_b.reserve((ntot + 1) * sizeof(u32)));
        MCU_TRY(radix_clear_hist(s.radix, s.stream));
        const u64 npos[2] = {npos0, npos1};
        u64 base = 0;
        for (int g = 0; g < 2; ++g) {
            if (npos[g]) {
                seedgen_kernel<K><<<grid_for(npos[g], 256, 8), 256, passes * 256 * sizeof(u32), s.stream>>>(
                    s.packed[g].as<u32>(), npos[g], sp, (u32)g, s.keys_a.as<K>(), s.vals_a.as<u32>(), sharded ? 0 : base,
                    s.radix.hist.as<u64>(), passes, (u32)shard_index, (u32)shard_count, ctr + 5);
                s.launches++;
            }
            base += npos[g];
        }
        if (sharded) {
            MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
            MCU_CUDA(cudaStreamSynchronize(s.stream));
            nsort = s.h_counters[5];
        }
        MCU_CUDA(cudaEventRecord(s.ev[2], s.stream));

        // ---- sort ----
        bool in_a = true;
        u64 before = s.radix.launches;
        MCU_TRY(radix_sort_pairs<K>(s.radix, s.keys_a.as<K>(), s.vals_a.as<u32>(), s.keys_b.as<K>(), s.vals_b.as<u32>(), nsort, key_bits, true,
                                    s.stream, &in_a, &passes_run));
        s.launches += s.radix.launches - before;
        const K* skeys = in_a ? s.keys_a.as<K>() : s.keys_b.as<K>();
        const u32* svals = in_a ? s.vals_a.as<u32>() : s.vals_b.as<u32>();
        MCU_CUDA(cudaEventRecord(s.ev[3], s.stream));

        // ---- join ----
        if (nsort) {
            join_kernel<K><<<(unsigned)div_up(nsort, 256 * 8), 256, 0, s.stream>>>(skeys, svals, nsort, s.uniq.as<u32>(), s.pairs.as<u64>(), pair_cap, ctr);
            s.launches++;
        }
    }
    MCU_CUDA(cudaEventRecord(s.ev[4], s.stream));
    MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    MCU_CUDA(cudaGetLastError());
    s.gap_seen = ((u32*)(s.h_counters + 4))[0] != 0;
    if (s.gap_seen && !s.defer_gap_error) { set_error("gap character '-' in a genome sequence (input must be unaligned)"); return MCU_EGAP; }
    s.run.sp = sp;
    s.run.sharded = sharded;
    s.run.bucketed = bucketed;
    s.run.passes = passes_run;
    s.run.nsort = nsort;
    s.run.pair_cap = pair_cap;
    s.run.uniq_words = uniq_words;
    s.run.enumerated = true;
    return MCU_OK;
}

// Phase 2: candidates -> extension -> reference list order (+ bucket replay when the bitmap covers the whole key space).
static int run_finish(Session& s, bool uniq_is_global, float* stage_ms, u64* stats)
{
    if (!s.run.enumerated) { set_error("mcu_session_finish without mcu_session_enumerate"); return MCU_EINVAL; }
    s.run.enumerated = false;
    const SeedParams& sp = s.run.sp;
    const u64 npos0 = s.n[0] >= (u64)sp.L ? s.n[0] - sp.L + 1 : 0;
    const u64 npos1 = s.n[1] >= (u64)sp.L ? s.n[1] - sp.L + 1 : 0;
    const bool sharded = s.run.sharded, bucketed = s.run.bucketed;
    const int passes_run = s.run.passes;
    const u64 nsort = s.run.nsort, pair_cap = s.run.pair_cap;
    unsigned long long* ctr = s.counters.as<unsigned long long>();
    s.run.uniq_global = uniq_is_global || !sharded;

    // ---- candidates + extend ----
    const u64 pfwd = s.h_counters[0], prev_ = s.h_counters[6];
    const u64 npairs = pfwd + prev_ + s.bk_direct;  // bk_direct: pairs the bucket kernel already classified as candidates
    const u64 repeat_flag = s.h_counters[1];
    u64 ncand = 0, nmatch = 0;
    if (npairs) {
        ExtendArgs ea;
        ea.g0 = s.packed[0].as<u32>(); ea.g1 = s.packed[1].as<u32>();
        ea.npos0 = npos0; ea.npos1 = npos1;
        ea.uniq = s.uniq.as<u32>();
        ea.cand = s.cand.as<u64>(); ea.nfwd = 0; ea.nrev = 0; ea.cap = pair_cap;
        ea.out = nullptr; ea.counters = ctr;
        MCU_CUDA(cudaEventRecord(s.kev[6], s.stream));
        const bool solid = sp.nruns == 1 && sp.L == sp.w && (sp.w & 1) && getenv("MAUVE_CUDA_NO_SOLID") == nullptr;
        candidate_kernel<<<grid_for(pfwd + prev_ + 1, 256, 8), 256, 0, s.stream>>>(ea, sp, s.pairs.as<u64>(), pfwd, prev_, pair_cap, solid,
                                                                                     s.bk_aux ? s.bk_group_fwd : 0, s.bk_aux ? s.bk_group_rev : 0);
        s.launches++;
        MCU_CUDA(cudaEventRecord(s.kev[7], s.stream));
        MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        MCU_CUDA(cudaStreamSynchronize(s.stream));
        ea.nfwd = s.h_counters[2]; ea.nrev = s.h_counters[7];
        ncand = ea.nfwd + ea.nrev;
        MCU_TRY(s.raw_matches.reserve((ncand + 1) * sizeof(mcu_match)));
        ea.out = s.raw_matches.as<mcu_match>();
        if (solid) extend_solid_kernel<<<grid_for(ncand * 32, 256, 8), 256, 0, s.stream>>>(ea, sp, (i64)s.n[0], (i64)s.n[1]);
        else extend_kernel<<<grid_for(ncand * 32, 256, 8), 256, 0, s.stream>>>(ea, sp);
        s.launches++;
        MCU_CUDA(cudaEventRecord(s.kev[8], s.stream));
        MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        MCU_CUDA(cudaStreamSynchronize(s.stream));
        nmatch = s.h_counters[3];
    }
    MCU_CUDA(cudaEventRecord(s.ev[5], s.stream));

    // ---- order ----
    u64 unclean = 0, dup_rows = 0;
    if (sharded && uniq_is_global) {
        // rank 0 orders the gathered rows of all ranks (session_merge): this rank's list goes out as it is
        s.match_count = nmatch;
        MCU_TRY(s.matches.reserve((nmatch + 1) * sizeof(mcu_match)));
        if (nmatch) MCU_CUDA(cudaMemcpyAsync(s.matches.p, s.raw_matches.p, nmatch * sizeof(mcu_match), cudaMemcpyDeviceToDevice, s.stream));
    } else {
        MCU_TRY(order_matches(s, s.raw_matches.as<mcu_match>(), nmatch, s.n[0]));
        MCU_TRY(replay_unclean(s, &sp, !sharded, &unclean, &dup_rows));  // sharded runs replay after the merge
        nmatch = s.match_count;
    }
    MCU_CUDA(cudaEventRecord(s.ev[6], s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    MCU_CUDA(cudaGetLastError());

    if (stage_ms) {
        for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&stage_ms[i], s.ev[i], s.ev[i + 1]);
        cudaEventElapsedTime(&stage_ms[6], s.ev[0], s.ev[6]);
        stage_ms[7] = bucketed ? (s.bk_exact ? -2.0f : -1.0f) : (float)passes_run;  // < 0: bucketed enumeration (-2: exact bucket sizes)
        for (int i = 8; i < 16; ++i) stage_ms[i] = 0.f;
        stage_ms[15] = (float)s.bk_spilled;
        if (bucketed)
            for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&stage_ms[8 + i], s.kev[i], s.kev[i + 1]);
        if (npairs) {
            cudaEventElapsedTime(&stage_ms[13], s.kev[6], s.kev[7]);
            if (ncand) cudaEventElapsedTime(&stage_ms[14], s.kev[7], s.kev[8]);
        }
    }
    if (stats) {
        stats[0] = npairs; stats[1] = nmatch; stats[2] = npairs - nmatch; stats[3] = repeat_flag;
        stats[4] = ncand; stats[5] = nsort; stats[6] = unclean; stats[7] = dup_rows;
    }
    return MCU_OK;
}

int join_sorted_u64(Session& s, const u64* keys, const u32* vals, u64 n, u64 pair_cap)
{
    if (!n) return MCU_OK;
    join_kernel<u64><<<(unsigned)div_up(n, 256 * 8), 256, 0, s.stream>>>(keys, vals, n, s.uniq.as<u32>(), s.pairs.as<u64>(), pair_cap,
                                                                        s.counters.as<unsigned long long>());
    s.launches++;
    MCU_CUDA(cudaGetLastError());
    return MCU_OK;
}

// rows with equal primary key (same hash bucket AND same genome-0 start: next to never, duplicates of the order-dependent buckets
// apart) are adjacent after the stable sort, still in their original order: the head of such a run orders it by the secondary key
// (stable insertion: what a second LSD sort underneath would have given)
__global__ void order_ties_kernel(const u64* __restrict__ pk, u32* __restrict__ idx, const u64* __restrict__ secondary, u64 n)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = pk[i];
        if (i > 0 && pk[i - 1] == k) continue;
        if (i + 1 >= n || pk[i + 1] != k) continue;
        u64 j = i + 2;
        while (j < n && pk[j] == k) ++j;
        for (u64 a = i + 1; a < j; ++a) {
            const u32 v = idx[a];
            const u64 kv = secondary[v];
            u64 b = a;
            while (b > i && secondary[idx[b - 1]] > kv) {
                idx[b] = idx[b - 1];
                --b;
            }
            idx[b] = v;
        }
    }
}

static int bits_for(u64 v)
{
    int b = 1;
    while (b < 64 && (v >> b) != 0) ++b;
    return b;
}

// max_start0: an upper bound of the rows' genome-0 starts (the genome's length) when the caller knows one, else 0
int order_matches(Session& s, const mcu_match* rows_dev, u64 n, u64 max_start0)
{
    s.match_count = n;
    MCU_TRY(s.matches.reserve((n + 1) * sizeof(mcu_match)));
    if (n == 0) return MCU_OK;
    if (n >= 0xFFFFFFFFull) { set_error("too many matches"); return MCU_EINVAL; }
    MCU_TRY(s.ord_keys_a.reserve(n * sizeof(u64)));
    MCU_TRY(s.ord_keys_b.reserve(n * sizeof(u64)));
    MCU_TRY(s.ord_vals_a.reserve(n * sizeof(u32)));
    MCU_TRY(s.ord_vals_b.reserve(n * sizeof(u32)));
    MCU_TRY(s.ord_primary.reserve(n * sizeof(u64)));
    u64* secondary = s.ord_primary.as<u64>();
    u64* primary = s.ord_keys_a.as<u64>();
    const int s0_bits = max_start0 ? bits_for(max_start0 + 1) : 33;  // starts < 2^32 (+ length)
    const int g = grid_for(n, 256, 8);
    order_keys_kernel<<<g, 256, 0, s.stream>>>(rows_dev, n, s0_bits, primary, secondary, s.ord_vals_a.as<u32>());
    s.launches++;
    bool in_a = true;
    u64 before = s.radix.launches;
    // ONE stable sort on bucket | start0 (16 + s0_bits bits); the secondary key only decides between rows that tie on it
    MCU_TRY(radix_sort_pairs<u64>(s.radix, s.ord_keys_a.as<u64>(), s.ord_vals_a.as<u32>(), s.ord_keys_b.as<u64>(), s.ord_vals_b.as<u32>(), n,
                                  s0_bits + 16, false, s.stream, &in_a, nullptr));
    s.launches += s.radix.launches - before;
    u32* idx_final = in_a ? s.ord_vals_a.as<u32>() : s.ord_vals_b.as<u32>();
    const u64* pk = in_a ? s.ord_keys_a.as<u64>() : s.ord_keys_b.as<u64>();
    order_ties_kernel<<<g, 256, 0, s.stream>>>(pk, idx_final, secondary, n);
    s.launches++;
    gather_rows_kernel<<<g, 256, 0, s.stream>>>(rows_dev, idx_final, n, s.matches.as<mcu_match>());
    s.launches++;
    MCU_CUDA(cudaGetLastError());
    return MCU_OK;
}

int session_run(Session& s, u64 seed, int shard_index, int shard_count, float* stage_ms, u64* stats)
{
    MCU_TRY(session_enumerate(s, seed, shard_index, shard_count));
    return run_finish(s, false, stage_ms, stats);
}

int session_enumerate(Session& s, u64 seed, int shard_index, int shard_count)
{
    SeedParams sp;
    MCU_TRY(make_seed_params(seed, &sp));
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) { set_error("bad shard spec"); return MCU_EINVAL; }
    if (2 * sp.w + 2 <= 32) return run_enumerate<u32>(s, sp, shard_index, shard_count);
    return run_enumerate<u64>(s, sp, shard_index, shard_count);
}

int session_finish(Session& s, bool uniq_is_global, float* stage_ms, u64* stats) { return run_finish(s, uniq_is_global, stage_ms, stats); }

// Rank-0 merge of the per-rank lists of one sharded run whose unique-seed bitmap was combined across ranks: every match
// has exactly one emitter, so there is nothing to dedupe; order the rows and replay order-dependent buckets exactly.
int session_merge(Session& s, const mcu_match* rows_dev, u64 n, u64* unclean, u64* dup_rows)
{
    *unclean = 0;
    *dup_rows = 0;
    MCU_TRY(order_matches(s, rows_dev, n, s.n[0]));
    MCU_TRY(replay_unclean(s, &s.run.sp, s.run.uniq_global, unclean, dup_rows));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    return MCU_OK;
}

// ---- single-genome SML --------------------------------------------------------------------
template <typename K>
static int sml_build_t(Session& s, const SeedParams& sp, u32* pos_out, u64* mer_out, u32* packed_out, u64* len_out, int shard = 0, int nshard = 1)
{
    const u64 n = s.n[0];
    const u64 npos_all = n >= (u64)sp.L ? n - sp.L + 1 : 0;
    u64 npos = npos_all;
    SmlSplit split;
    if (nshard > 1) {
        const int kb = 2 * sp.w + 2;
        const int tb = kb < 16 ? kb : 16;
        split.n = (u32)nshard;
        split.shift = kb - tb;
        for (int k = 1; k < nshard; ++k) split.s[k - 1] = (u32)((double)(1u << tb) * (1.0 - sqrt(1.0 - (double)k / nshard)));
    }
    const int key_bits = 2 * sp.w + 2;
    const int passes = (key_bits + 7) / 8;
    unsigned long long* ctr = s.counters.as<unsigned long long>();
    MCU_CUDA(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), s.stream));
    MCU_CUDA(cudaEventRecord(s.ev[0], s.stream));
    MCU_TRY(run_pack(s, 0, (u32*)(ctr + 4)));
    MCU_CUDA(cudaEventRecord(s.ev[1], s.stream));
    MCU_TRY(s.keys_a.reserve((npos + 1) * sizeof(u64)));  // u64-sized: the mer conversion reuses keys_b
    MCU_TRY(s.keys_b.reserve((npos + 1) * sizeof(u64)));
    MCU_TRY(s.vals_a.reserve((npos + 1) * sizeof(u32)));
    MCU_TRY(s.vals_b.reserve((npos + 1) * sizeof(u32)));
    MCU_TRY(radix_clear_hist(s.radix, s.stream));
    bool in_a = true;
    if (npos) {
        seedgen_kernel<K><<<grid_for(npos, 256, 8), 256, passes * 256 * sizeof(u32), s.stream>>>(
            s.packed[0].as<u32>(), npos, sp, 0u, s.keys_a.as<K>(), s.vals_a.as<u32>(), 0, s.radix.hist.as<u64>(), passes, (u32)shard, (u32)nshard, ctr + 5,
            split);
        s.launches++;
        if (nshard > 1) {   // how many seeds this shard owns: sizes the sort
            MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
            MCU_CUDA(cudaStreamSynchronize(s.stream));
            npos = s.h_counters[5];
        }
        MCU_CUDA(cudaEventRecord(s.ev[2], s.stream));
        u64 before = s.radix.launches;
        MCU_TRY(radix_sort_pairs<K>(s.radix, s.keys_a.as<K>(), s.vals_a.as<u32>(), s.keys_b.as<K>(), s.vals_b.as<u32>(), npos, key_bits, true,
                                    s.stream, &in_a, &s.sml_passes));
        s.launches += s.radix.launches - before;
    } else
        MCU_CUDA(cudaEventRecord(s.ev[2], s.stream));
    MCU_CUDA(cudaEventRecord(s.ev[3], s.stream));
    MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&s.sml_ms[i], s.ev[i], s.ev[i + 1]);
    s.sml_key_bytes_last = (int)sizeof(K);
    if (((u32*)(s.h_counters + 4))[0]) { set_error("gap character '-' in a genome sequence (input must be unaligned)"); return MCU_EGAP; }
    const K* skeys = in_a ? s.keys_a.as<K>() : s.keys_b.as<K>();
    const u32* svals = in_a ? s.vals_a.as<u32>() : s.vals_b.as<u32>();
    s.sml_keys = skeys;  // consumers of the sorted list on the device (sol.cu)
    s.sml_vals = svals;
    s.sml_npos = npos;
    s.sml_key_bytes = (int)sizeof(K);
    if (pos_out && npos) MCU_CUDA(cudaMemcpyAsync(pos_out, svals, npos * sizeof(u32), cudaMemcpyDeviceToHost, s.stream));
    if (mer_out && npos) {
        u64* mers = in_a ? s.keys_b.as<u64>() : s.keys_a.as<u64>();  // the buffer not holding the result
        keys_to_mers_kernel<K><<<grid_for(npos, 256, 8), 256, 0, s.stream>>>(skeys, npos, sp.w, mers);
        s.launches++;
        MCU_CUDA(cudaMemcpyAsync(mer_out, mers, npos * sizeof(u64), cudaMemcpyDeviceToHost, s.stream));
    }
    if (packed_out) MCU_CUDA(cudaMemcpyAsync(packed_out, s.packed[0].p, (div_up(n, 16) + 2) * sizeof(u32), cudaMemcpyDeviceToHost, s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    MCU_CUDA(cudaGetLastError());
    if (len_out) *len_out = npos;
    return MCU_OK;
}

int sml_build_device(Session& s, const char* seq, u64 n, u64 seed, u32* pos_out, u64* mer_out, u32* packed_out, u64* len_out, int shard, int nshard)
{
    SeedParams sp;
    MCU_TRY(make_seed_params(seed, &sp));
    if (nshard < 1 || nshard > 16 || shard < 0 || shard >= nshard) { set_error("sorted mer list: bad shard %d of %d (at most 16)", shard, nshard); return MCU_EINVAL; }
    MCU_TRY(session_upload(s, seq, n, nullptr, 0));
    if (2 * sp.w + 2 <= 32) return sml_build_t<u32>(s, sp, pos_out, mer_out, packed_out, len_out, shard, nshard);
    return sml_build_t<u64>(s, sp, pos_out, mer_out, packed_out, len_out, shard, nshard);
}

}  // namespace mcu
