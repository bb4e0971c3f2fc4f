// Batched gap search (sm_100a): many small sequence pairs in one pass.
//
// Recursive anchoring calls the whole seed-match path once per inter-anchor gap: pairwiseAnchorSearch
// (LM/ProgressiveAligner.cpp:590-679) builds two DNAMemorySMLs on the gap substrings with the weight
// getDefaultSeedWeight(average gap length) and pattern rank 0 (:617-626, LM/SeedMasks.h:298-401), runs MemHash with the MUM
// settings (:649-651) and shifts the coordinates back; recurseOnPairs (:681-924) does that for every gap of every round,
// SearchLCBGaps (LM/Aligner.cpp:784-930) for the inter-LCB regions: thousands of pairs of tens of bp to a few kbp.  One
// launch sequence per pair would be all launch latency, so the pairs that share a seed pattern are processed together:
//
//   cat0 / cat1   the genome-0 / genome-1 sides of all pairs, concatenated (segment s = [off[s], off[s+1]))
//   keys          segment << (2w+2) | canonical mer << 2 | genome << 1 | strand     one radix sort for the whole batch; the
//                 segment bits keep equal mers of different pairs apart, so join_kernel's local test is per pair
//   candidates / extension: as in anchor.cu, with the diagonal walk confined to the segment pair (DiagBounds)
//   order         MemHash::GetMatchList order (LM/MemHash.h:183-203) inside every segment, segments ascending
//
// Hash buckets whose content depends on the reference's insertion order (csrc/replay.cu) need two diagonals 40000 apart
// in one pair; such pairs are reported back and the caller runs them through the single-pair path, which replays them.
#include "anchor.cuh"

namespace mcu {

__device__ __forceinline__ u32 bt_find_seg(const u64* __restrict__ off, u32 n, u64 p)
{
    u32 lo = 0, hi = n;  // off[lo] <= p < off[hi]
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) bt_seedgen_kernel(const u32* __restrict__ packed, const u64* __restrict__ off, u32 nseg, u64 nbases, SeedParams sp,
                                                        u32 genome, int seg_shift, u64* __restrict__ keys, u32* __restrict__ vals,
                                                        unsigned long long* __restrict__ counter)
{
    const u32 lane = threadIdx.x & 31;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (nbases + stride - 1) / stride;
    for (u64 r = 0; r < rounds; ++r) {
        const u64 p = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool live = p < nbases;
        u64 key = 0;
        if (live) {
            const u32 seg = bt_find_seg(off, nseg, p);
            live = p + (u64)sp.L <= __ldg(off + seg + 1);  // the seed window stays inside its segment (SMLLength = n - L + 1 per sequence)
            if (live) {
                const u64 f = extract_seed(load_mer32(packed, p), sp);
                const u64 rc = revcomp_seed(f, sp.w);
                const u32 strand = rc < f;
                key = ((u64)seg << seg_shift) | ((strand ? rc : f) << 2) | (genome << 1) | strand;
            }
        }
        const u32 m = __ballot_sync(0xffffffffu, live);
        if (m) {
            u64 base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) {
                const u64 slot = base + __popc(m & lanemask_lt());
                keys[slot] = key;
                vals[slot] = (u32)p;
            }
        }
    }
}

struct BtSegs {
    const u64* off0;
    const u64* off1;
    u32 nseg;
};

__device__ __forceinline__ DiagBounds bt_bounds(const BtSegs& sg, u32 seg, int L)
{
    DiagBounds b;
    b.lo0 = (i64)__ldg(sg.off0 + seg);
    b.hi0 = (i64)__ldg(sg.off0 + seg + 1) - L + 1;
    b.lo1 = (i64)__ldg(sg.off1 + seg);
    b.hi1 = (i64)__ldg(sg.off1 + seg + 1) - L + 1;
    return b;
}

// candidate filter of anchor.cu (left neighbour on the diagonal is a unique seed pair -> not the leftmost), per segment
__global__ void __launch_bounds__(256) bt_candidate_kernel(ExtendArgs a, SeedParams sp, BtSegs sg, const u64* __restrict__ pairs, u64 pfwd, u64 prev_, u64 pair_cap)
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
            const DiagBounds b = bt_bounds(sg, bt_find_seg(sg.off0, sg.nseg, (u64)p0), sp.L);
            i64 other;
            is_cand = !(p0 > b.lo0 && uniq_bit(a.uniq, p0 - 1) && probe_hit_in(a, sp, rev, d, p0 - 1, other, b));
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

// extend_kernel of anchor.cu confined to the segment pair; rows come out in segment-local 1-based coordinates
__global__ void __launch_bounds__(256) bt_extend_kernel(ExtendArgs a, SeedParams sp, BtSegs sg, u32* __restrict__ row_seg)
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
        const u32 seg = bt_find_seg(sg.off0, sg.nseg, (u64)t0);
        const DiagBounds b = bt_bounds(sg, seg, sp.L);
        i64 cur = t0;
        bool abandoned = false;
        while (true) {  // walk left
            i64 t = cur - 1 - (i64)lane, other = 0;
            const bool h = (i64)lane < L && probe_hit_in(a, sp, rev, d, t, other, b);
            const bool uq = h && uniq_bit(a.uniq, t);
            if (__any_sync(0xffffffffu, uq)) { abandoned = true; break; }
            const u32 hits = __ballot_sync(0xffffffffu, h);
            if (!hits) break;
            cur -= 32 - __clz(hits);
        }
        if (abandoned) continue;
        const i64 lo = cur;
        cur = t0;
        while (true) {  // walk right
            i64 t = cur + 1 + (i64)lane, other = 0;
            const bool h = (i64)lane < L && probe_hit_in(a, sp, rev, d, t, other, b);
            const u32 hits = __ballot_sync(0xffffffffu, h);
            if (!hits) break;
            cur += 32 - __clz(hits);
        }
        const i64 hi = cur;
        if (lane == 0) {
            const u64 slot = atomicAdd(&a.counters[3], 1ull);
            mcu_match m;
            m.len = hi - lo + L;
            m.start0 = lo - b.lo0 + 1;
            m.start1 = rev ? -((d - hi) - b.lo1 + 1) : lo + d - b.lo1 + 1;
            a.out[slot] = m;
            row_seg[slot] = seg;
        }
    }
}

// sort keys of order_keys_kernel (anchor.cu) + the segment as the most significant key
__global__ void bt_order_keys_kernel(const mcu_match* __restrict__ rows, u64 n, int s0_bits, u64* __restrict__ primary, u64* __restrict__ secondary,
                                     u32* __restrict__ idx)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const mcu_match m = rows[i];
        const i64 off = m.start1 - m.start0 - (m.start1 < 0 ? m.len : 0);
        const i64 b = ((off % 40000) + 40000) % 40000;
        const u64 s1 = m.start1 < 0 ? (u64)(-m.start1 + m.len) : (u64)m.start1;
        primary[i] = ((u64)b << s0_bits) | (u64)m.start0;
        secondary[i] = (s1 << 1) | (m.start1 < 0 ? 1u : 0u);
        idx[i] = (u32)i;
    }
}

__global__ void bt_gather_u64_kernel(const u64* __restrict__ src, const u32* __restrict__ idx, u64 n, u64* __restrict__ dst)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void bt_gather_seg_kernel(const u32* __restrict__ seg, const u32* __restrict__ idx, u64 n, u64* __restrict__ dst)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] = seg[idx[i]];
}

// final gather + detection of order-dependent buckets: in (segment, bucket, start0) order a row that starts inside or directly
// after its predecessor's span (same segment and bucket) makes the bucket "unclean" (csrc/replay.cu)
__global__ void bt_finish_kernel(const mcu_match* __restrict__ rows, const u32* __restrict__ seg, const u32* __restrict__ idx, u64 n,
                                 mcu_match* __restrict__ out_rows, u32* __restrict__ out_seg, u32* __restrict__ unclean_flag)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const mcu_match m = rows[idx[i]];
        const u32 sg = seg[idx[i]];
        out_rows[i] = m;
        out_seg[i] = sg;
        if (i > 0) {
            const mcu_match q = rows[idx[i - 1]];
            if (seg[idx[i - 1]] == sg) {
                const i64 om = m.start1 - m.start0 - (m.start1 < 0 ? m.len : 0), oq = q.start1 - q.start0 - (q.start1 < 0 ? q.len : 0);
                const i64 bm = ((om % 40000) + 40000) % 40000, bq = ((oq % 40000) + 40000) % 40000;
                if (bm == bq && m.start0 <= q.start0 + q.len) unclean_flag[sg] = 1u;
            }
        }
    }
}

static int bt_grid(u64 n, int per_sm)
{
    u64 want = div_up(n, 256), cap = (u64)sm_count() * per_sm;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

static int bt_bits(u64 x)
{
    int b = 0;
    while (x) { ++b; x >>= 1; }
    return b;
}

int batch_find_mums(Session& s, const char* cat0, const u64* off0, const char* cat1, const u64* off1, u32 n_seg, u64 seed,
                    std::vector<mcu_match>* rows_out, std::vector<u32>* seg_out, std::vector<u32>* unclean_segs, u64* seed_pairs)
{
    rows_out->clear();
    seg_out->clear();
    unclean_segs->clear();
    if (seed_pairs) *seed_pairs = 0;
    if (n_seg == 0) return MCU_OK;
    SeedParams sp;
    MCU_TRY(make_seed_params(seed, &sp));
    const u64 n0 = off0[n_seg], n1 = off1[n_seg];
    if (n0 < (u64)sp.L || n1 < (u64)sp.L) return MCU_OK;
    const int seg_bits = bt_bits(n_seg - 1) ? bt_bits(n_seg - 1) : 1, seg_shift = 2 * sp.w + 2;
    if (seg_shift + seg_bits > 64) { set_error("mcu_find_mums_batch: %u pairs with seed weight %d do not fit a 64-bit key", n_seg, sp.w); return MCU_EINVAL; }
    MCU_TRY(session_upload(s, cat0, n0, cat1, n1));  // also checks the 32-bit position limit
    cudaStream_t st = s.stream;
    unsigned long long* ctr = s.counters.as<unsigned long long>();
    MCU_CUDA(cudaMemsetAsync(ctr, 0, 8 * sizeof(unsigned long long), st));
    for (int g = 0; g < 2; ++g) MCU_TRY(run_pack_genome(s, g, (u32*)(ctr + 4)));
    // segment tables on the device: off0 | off1 (n_seg + 1 each), unclean flags
    MCU_TRY(s.bt_tab.reserve((size_t)(n_seg + 1) * 16 + (size_t)n_seg * 4 + 64));
    u64* d_off0 = s.bt_tab.as<u64>();
    u64* d_off1 = d_off0 + n_seg + 1;
    u32* d_unclean = (u32*)(d_off1 + n_seg + 1);
    MCU_CUDA(cudaMemcpyAsync(d_off0, off0, (size_t)(n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
    MCU_CUDA(cudaMemcpyAsync(d_off1, off1, (size_t)(n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
    MCU_CUDA(cudaMemsetAsync(d_unclean, 0, (size_t)n_seg * 4, st));
    const u64 ntot = n0 + n1;
    MCU_TRY(s.keys_a.reserve((ntot + 1) * 8));
    MCU_TRY(s.keys_b.reserve((ntot + 1) * 8));
    MCU_TRY(s.vals_a.reserve((ntot + 1) * 4));
    MCU_TRY(s.vals_b.reserve((ntot + 1) * 4));
    const u64 pair_cap = (n0 < n1 ? n0 : n1) + 1, uniq_words = div_up(n0 + 1, 32) + 1;
    MCU_TRY(s.uniq.reserve(uniq_words * 4));
    MCU_TRY(s.pairs.reserve(pair_cap * 8));
    MCU_TRY(s.cand.reserve(pair_cap * 8));
    MCU_CUDA(cudaMemsetAsync(s.uniq.p, 0, uniq_words * 4, st));
    bt_seedgen_kernel<<<bt_grid(n0, 8), 256, 0, st>>>(s.packed[0].as<u32>(), d_off0, n_seg, n0, sp, 0u, seg_shift, s.keys_a.as<u64>(), s.vals_a.as<u32>(), ctr + 5);
    bt_seedgen_kernel<<<bt_grid(n1, 8), 256, 0, st>>>(s.packed[1].as<u32>(), d_off1, n_seg, n1, sp, 1u, seg_shift, s.keys_a.as<u64>(), s.vals_a.as<u32>(), ctr + 5);
    s.launches += 2;
    MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 64, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    if (((u32*)(s.h_counters + 4))[0]) { set_error("gap character '-' in a sequence (input must be unaligned)"); return MCU_EGAP; }
    const u64 nkeys = s.h_counters[5];
    if (!nkeys) return MCU_OK;
    bool in_a = true;
    u64 before = s.radix.launches;
    MCU_TRY(radix_sort_pairs<u64>(s.radix, s.keys_a.as<u64>(), s.vals_a.as<u32>(), s.keys_b.as<u64>(), s.vals_b.as<u32>(), nkeys, seg_shift + seg_bits, false,
                                  st, &in_a, nullptr));
    s.launches += s.radix.launches - before;
    MCU_TRY(join_sorted_u64(s, in_a ? s.keys_a.as<u64>() : s.keys_b.as<u64>(), in_a ? s.vals_a.as<u32>() : s.vals_b.as<u32>(), nkeys, pair_cap));
    MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 64, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    const u64 pfwd = s.h_counters[0], prev_ = s.h_counters[6];
    if (seed_pairs) *seed_pairs = pfwd + prev_;
    if (pfwd + prev_ == 0) return MCU_OK;
    ExtendArgs ea;
    ea.g0 = s.packed[0].as<u32>(); ea.g1 = s.packed[1].as<u32>();
    ea.npos0 = n0; ea.npos1 = n1;
    ea.uniq = s.uniq.as<u32>();
    ea.cand = s.cand.as<u64>(); ea.nfwd = 0; ea.nrev = 0; ea.cap = pair_cap;
    ea.out = nullptr; ea.counters = ctr;
    BtSegs sg;
    sg.off0 = d_off0; sg.off1 = d_off1; sg.nseg = n_seg;
    bt_candidate_kernel<<<bt_grid(pfwd + prev_, 8), 256, 0, st>>>(ea, sp, sg, s.pairs.as<u64>(), pfwd, prev_, pair_cap);
    s.launches++;
    MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 64, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    ea.nfwd = s.h_counters[2]; ea.nrev = s.h_counters[7];
    const u64 ncand = ea.nfwd + ea.nrev;
    MCU_TRY(s.raw_matches.reserve((ncand + 1) * sizeof(mcu_match)));
    MCU_TRY(s.bt_seg_a.reserve((ncand + 1) * 4));
    MCU_TRY(s.bt_seg_b.reserve((ncand + 1) * 4));
    ea.out = s.raw_matches.as<mcu_match>();
    bt_extend_kernel<<<bt_grid(ncand * 32, 8), 256, 0, st>>>(ea, sp, sg, s.bt_seg_a.as<u32>());
    s.launches++;
    MCU_CUDA(cudaMemcpyAsync(s.h_counters, ctr, 64, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    const u64 n = s.h_counters[3];
    if (!n) return MCU_OK;
    if (n >= 0xFFFFFFFFull) { set_error("too many matches"); return MCU_EINVAL; }
    // ---- order: three stable sorts (secondary, bucket | start0, segment) ----
    MCU_TRY(s.matches.reserve((n + 1) * sizeof(mcu_match)));
    MCU_TRY(s.ord_keys_a.reserve(n * 8));
    MCU_TRY(s.ord_keys_b.reserve(n * 8));
    MCU_TRY(s.ord_vals_a.reserve(n * 4));
    MCU_TRY(s.ord_vals_b.reserve(n * 4));
    MCU_TRY(s.ord_primary.reserve(n * 8));
    const int s0_bits = 33, s1_bits = 36, g = bt_grid(n, 8);
    const mcu_match* raw = s.raw_matches.as<mcu_match>();
    bt_order_keys_kernel<<<g, 256, 0, st>>>(raw, n, s0_bits, s.ord_primary.as<u64>(), s.ord_keys_a.as<u64>(), s.ord_vals_a.as<u32>());
    before = s.radix.launches;
    MCU_TRY(radix_sort_pairs<u64>(s.radix, s.ord_keys_a.as<u64>(), s.ord_vals_a.as<u32>(), s.ord_keys_b.as<u64>(), s.ord_vals_b.as<u32>(), n, s1_bits, false, st,
                                  &in_a, nullptr));
    u32* idx = in_a ? s.ord_vals_a.as<u32>() : s.ord_vals_b.as<u32>();
    u32* idx_other = in_a ? s.ord_vals_b.as<u32>() : s.ord_vals_a.as<u32>();
    bt_gather_u64_kernel<<<g, 256, 0, st>>>(s.ord_primary.as<u64>(), idx, n, s.ord_keys_a.as<u64>());
    MCU_TRY(radix_sort_pairs<u64>(s.radix, s.ord_keys_a.as<u64>(), idx, s.ord_keys_b.as<u64>(), idx_other, n, s0_bits + 16, false, st, &in_a, nullptr));
    if (!in_a) { u32* t = idx; idx = idx_other; idx_other = t; }
    bt_gather_seg_kernel<<<g, 256, 0, st>>>(s.bt_seg_a.as<u32>(), idx, n, s.ord_keys_a.as<u64>());
    MCU_TRY(radix_sort_pairs<u64>(s.radix, s.ord_keys_a.as<u64>(), idx, s.ord_keys_b.as<u64>(), idx_other, n, seg_bits, false, st, &in_a, nullptr));
    if (!in_a) { u32* t = idx; idx = idx_other; idx_other = t; }
    bt_finish_kernel<<<g, 256, 0, st>>>(raw, s.bt_seg_a.as<u32>(), idx, n, s.matches.as<mcu_match>(), s.bt_seg_b.as<u32>(), d_unclean);
    s.launches += 4 + (s.radix.launches - before);
    MCU_CUDA(cudaGetLastError());
    rows_out->resize(n);
    seg_out->resize(n);
    std::vector<u32> flags(n_seg);
    MCU_CUDA(cudaMemcpyAsync(rows_out->data(), s.matches.p, n * sizeof(mcu_match), cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaMemcpyAsync(seg_out->data(), s.bt_seg_b.p, n * 4, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaMemcpyAsync(flags.data(), d_unclean, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    MCU_CUDA(cudaStreamSynchronize(st));
    for (u32 i = 0; i < n_seg; ++i)
        if (flags[i]) unclean_segs->push_back(i);
    s.match_count = 0;
    return MCU_OK;
}

}  // namespace mcu
