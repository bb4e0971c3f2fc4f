// The step after the match list, for two genomes, on the device (sm_100a).
//
// Replaces   mems::EliminateOverlaps_v2      LM/ProgressiveAligner.h:300-394 (checkConsistent :272-291, processNewMatch :260-271)
//            GenericMatchList::LengthFilter  LM/MatchList.h:680-692
//            mems::IdentifyBreakpoints       LM/GreedyBreakpointElimination.h:161-226
//            mems::ComputeLCBs_v2            LM/GreedyBreakpointElimination.h:229-250
// as pairwiseAnchorSearch runs them after every gap search (LM/ProgressiveAligner.cpp:656-660: overlaps, then LengthFilter) and
// the pairwise LCB set-up after the initial anchoring (:3408-3418: overlaps with eliminate_both, breakpoints, LCBs).
//
// A match is (len, start0 > 0, start1 signed).  EliminateOverlaps_v2 orders the list on one genome's start and sweeps it: where two
// matches overlap the shorter one is cropped (or deleted), with eliminate_both an inconsistent overlap costs both.  The sweep is
// sequential only inside a run of matches chained by overlaps: a running maximum of the ends finds the runs (one block-wide scan),
// one thread sweeps each run with the reference's own loop, and a stable compaction drops what was deleted.  Breakpoints are a
// neighbour test on the genome-0 labels in genome-1 order.
//
// The ORDER matters where starts tie (after the first pass they do: 25 times on the MDS42 list): the reference's std::sort leaves
// tied rows where libstdc++'s introsort happens to move them, and the sweep's outcome depends on it.  The list is therefore ordered
// by a stable radix sort when no two keys tie (any correct sort gives the same list then), and otherwise by ss_std_sort below: that
// very algorithm (bits/stl_algo.h), run by ONE thread on (key, index) words -- slow (~20 ms for 30 k rows) and exact.
#include "lcb.cuh"

#include "radix.cuh"

namespace mcu {

// ---- std::sort as libstdc++ implements it, on words compared by their high bits only (the low SS_SHIFT bits carry the row index) ----
constexpr int SS_SHIFT = 31;
#define SS_LT(a, b) (((a) >> SS_SHIFT) < ((b) >> SS_SHIFT))

__device__ void ss_adjust_heap(u64* first, i64 hole, i64 len, u64 value)
{
    const i64 top = hole;
    i64 child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (SS_LT(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    i64 parent = (hole - 1) / 2;   // __push_heap
    while (hole > top && SS_LT(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

__device__ void ss_heap_sort(u64* first, i64 len)   // __partial_sort(first, last, last)
{
    if (len >= 2)
        for (i64 parent = (len - 2) / 2;; --parent) {   // __make_heap
            ss_adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
        }
    for (i64 last = len; last > 1; --last) {             // __sort_heap
        const u64 value = first[last - 1];
        first[last - 1] = first[0];
        ss_adjust_heap(first, 0, last - 1, value);
    }
}

__device__ __forceinline__ void ss_unguarded_linear_insert(u64* last)
{
    const u64 val = *last;
    u64* next = last - 1;
    while (SS_LT(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}

__device__ void ss_insertion_sort(u64* first, u64* last)
{
    if (first == last) return;
    for (u64* i = first + 1; i != last; ++i) {
        if (SS_LT(*i, *first)) {
            const u64 val = *i;
            for (u64* p = i; p != first; --p) *p = *(p - 1);   // move_backward(first, i, i + 1)
            *first = val;
        } else
            ss_unguarded_linear_insert(i);
    }
}

// __introsort_loop without recursion: the right part of every partition waits on an explicit stack (at most 2 lg n + 1 deep)
__device__ void ss_std_sort(u64* base, i64 n)
{
    if (n <= 0) return;
    int lg = 0;
    for (i64 m = n; m > 1; m >>= 1) ++lg;
    struct Frame { u64 *first, *last; int depth; };
    Frame stack[130];
    int sp = 0;
    stack[sp++] = Frame{base, base + n, 2 * lg};
    while (sp) {
        Frame f = stack[--sp];
        u64 *first = f.first, *last = f.last;
        int depth = f.depth;
        // the reference recurses into [cut, last) FIRST and then continues with [first, cut): the two ranges are disjoint, so the
        // order in which they are processed does not change what either ends up holding
        while (last - first > 16) {
            if (depth == 0) { ss_heap_sort(first, last - first); break; }
            --depth;
            u64* mid = first + (last - first) / 2;
            u64 *a = first + 1, *b = mid, *c = last - 1, t;   // __move_median_to_first(first, a, b, c)
            if (SS_LT(*a, *b)) {
                if (SS_LT(*b, *c)) { t = *first; *first = *b; *b = t; }
                else if (SS_LT(*a, *c)) { t = *first; *first = *c; *c = t; }
                else { t = *first; *first = *a; *a = t; }
            } else if (SS_LT(*a, *c)) { t = *first; *first = *a; *a = t; }
            else if (SS_LT(*b, *c)) { t = *first; *first = *c; *c = t; }
            else { t = *first; *first = *b; *b = t; }
            u64 *lo = first + 1, *hi = last;                  // __unguarded_partition(first + 1, last, first)
            for (;;) {
                while (SS_LT(*lo, *first)) ++lo;
                --hi;
                while (SS_LT(*first, *hi)) --hi;
                if (!(lo < hi)) break;
                t = *lo; *lo = *hi; *hi = t;
                ++lo;
            }
            stack[sp++] = Frame{lo, last, depth};
            last = lo;
        }
    }
    if (n > 16) {                                             // __final_insertion_sort
        ss_insertion_sort(base, base + 16);
        for (u64* i = base + 16; i != base + n; ++i) ss_unguarded_linear_insert(i);
    } else
        ss_insertion_sort(base, base + n);
}

__global__ void lcb_std_sort_kernel(u64* w, u64 n)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) ss_std_sort(w, (i64)n);
}

// ---- ordering -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ i64 lcb_abs(i64 x) { return x < 0 ? -x : x; }
__device__ __forceinline__ i64 lcb_start(const mcu_match& r, int seq) { return seq ? r.start1 : r.start0; }

__global__ void lcb_keys_kernel(const mcu_match* __restrict__ rows, u64 n, int seq, u64* __restrict__ keys, u32* __restrict__ idx, u64* __restrict__ words)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = (u64)lcb_abs(lcb_start(rows[i], seq));
        keys[i] = k;
        idx[i] = (u32)i;
        words[i] = (k << SS_SHIFT) | i;
    }
}

__global__ void lcb_ties_kernel(const u64* __restrict__ sorted_keys, u64 n, unsigned long long* __restrict__ ties)
{
    unsigned long long c = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (u64)gridDim.x * blockDim.x) c += sorted_keys[i] == sorted_keys[i - 1];
    if (c) atomicAdd(ties, c);
}

__global__ void lcb_words_to_idx_kernel(const u64* __restrict__ words, u64 n, u32* __restrict__ idx)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) idx[i] = (u32)(words[i] & ((1ull << SS_SHIFT) - 1));
}

__global__ void lcb_gather_kernel(const mcu_match* __restrict__ src, const u32* __restrict__ idx, u64 n, mcu_match* __restrict__ dst)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

// ---- block-wide scans (one CTA of 1024 threads walks the array in chunks: the lists are a few 10^4 .. 10^6 rows) -----------------
constexpr int LCB_BLOCK = 1024;

// head[i] = 1 when no earlier row of the ordered list reaches row i's start: row i opens a run of rows chained by overlaps
__global__ void __launch_bounds__(LCB_BLOCK) lcb_heads_kernel(const mcu_match* __restrict__ v, u64 n, int seq, u8* __restrict__ head)
{
    __shared__ i64 warp_max[LCB_BLOCK / 32];
    __shared__ i64 carry_s;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += LCB_BLOCK) {
        const u64 i = base + tid;
        i64 start = 0, end = 0;
        if (i < n) {
            start = lcb_abs(lcb_start(v[i], seq));
            end = start + v[i].len;
        }
        i64 incl = end;   // inclusive running maximum of the ends inside the chunk
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const i64 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (u32)o && t > incl) incl = t;
        }
        if (lane == 31) warp_max[warp] = incl;
        __syncthreads();
        i64 before = carry_s;   // maximum over everything in front of this thread's row
        for (u32 w = 0; w < warp; ++w) before = warp_max[w] > before ? warp_max[w] : before;
        const i64 up = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane > 0 && up > before) before = up;
        if (i < n) head[i] = (i == 0 || start >= before) ? 1 : 0;
        __syncthreads();
        if (tid == LCB_BLOCK - 1) {
            i64 m = carry_s;
            for (u32 w = 0; w < LCB_BLOCK / 32; ++w) m = warp_max[w] > m ? warp_max[w] : m;
            carry_s = m;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void lcb_crop_start(mcu_match& r, i64 amt) { r.len -= amt; if (r.start0 > 0) r.start0 += amt; if (r.start1 > 0) r.start1 += amt; }
__device__ __forceinline__ void lcb_crop_end(mcu_match& r, i64 amt) { r.len -= amt; if (r.start0 < 0) r.start0 -= amt; if (r.start1 < 0) r.start1 -= amt; }

// one thread per run: the reference's double loop (LM/ProgressiveAligner.h:318-382) confined to the run, rows edited in place
__global__ void lcb_sweep_kernel(mcu_match* __restrict__ v, u64 n, int seq, int eliminate_both, const u8* __restrict__ head, u8* __restrict__ alive)
{
    for (u64 h = (u64)blockIdx.x * blockDim.x + threadIdx.x; h < n; h += (u64)gridDim.x * blockDim.x) {
        if (!head[h]) continue;
        u64 e = h + 1;
        while (e < n && !head[e]) ++e;
        for (u64 mi = h; mi < e; ++mi) {
            if (!alive[mi]) continue;
            for (u64 ni = mi + 1; ni < e; ++ni) {
                if (!alive[ni]) continue;
                mcu_match m = v[mi], x = v[ni];
                const i64 len_i = m.len;
                i64 diff = lcb_abs(lcb_start(x, seq)) - lcb_abs(lcb_start(m, seq)) - len_i;
                if (diff >= 0) break;
                diff = -diff;
                const bool smaller = x.len > m.len;   // equal multiplicity (2): the longer one in this genome wins
                const bool consistent = (x.start0 - m.start0) == (x.start1 - m.start1);
                bool deleted = false;
                if ((!consistent && eliminate_both) || smaller) {
                    if (diff >= len_i) { alive[mi] = 0; deleted = true; }
                    else {   // CropRight(diff, seq)
                        if (lcb_start(m, seq) > 0) lcb_crop_end(m, diff); else lcb_crop_start(m, diff);
                        v[mi] = m;
                    }
                }
                if ((!consistent && eliminate_both) || !smaller) {
                    if (diff >= x.len) alive[ni] = 0;
                    else {   // CropLeft(diff, seq)
                        if (lcb_start(x, seq) > 0) lcb_crop_start(x, diff); else lcb_crop_end(x, diff);
                        v[ni] = x;
                    }
                }
                if (deleted) break;
            }
        }
    }
}

// stable compaction: out gets the rows with keep[i] != 0 (and, when min_length > 0, len >= min_length); *count their number
__global__ void __launch_bounds__(LCB_BLOCK) lcb_compact_rows_kernel(const mcu_match* __restrict__ v, const u8* __restrict__ keep, u64 n, u64 min_length,
                                                                   mcu_match* __restrict__ out, unsigned long long* __restrict__ count)
{
    __shared__ u32 warp_sum[LCB_BLOCK / 32];
    __shared__ unsigned long long carry_s;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += LCB_BLOCK) {
        const u64 i = base + tid;
        mcu_match r;
        u32 k = 0;
        if (i < n) {
            r = v[i];
            k = keep[i] && (min_length == 0 || (u64)r.len >= min_length);
        }
        const u32 ballot = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_sum[warp] = __popc(ballot);
        __syncthreads();
        unsigned long long off = carry_s;
        for (u32 w = 0; w < warp; ++w) off += warp_sum[w];
        off += __popc(ballot & ((1u << lane) - 1u));
        if (k) out[off] = r;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = carry_s;
            for (u32 w = 0; w < LCB_BLOCK / 32; ++w) t += warp_sum[w];
            carry_s = t;
        }
        __syncthreads();
    }
    if (tid == 0) *count = carry_s;
}

// indices of the set flags, ascending
__global__ void __launch_bounds__(LCB_BLOCK) lcb_compact_flags_kernel(const u8* __restrict__ flag, u64 n, u64* __restrict__ out, unsigned long long* __restrict__ count)
{
    __shared__ u32 warp_sum[LCB_BLOCK / 32];
    __shared__ unsigned long long carry_s;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += LCB_BLOCK) {
        const u64 i = base + tid;
        const u32 k = i < n && flag[i];
        const u32 ballot = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_sum[warp] = __popc(ballot);
        __syncthreads();
        unsigned long long off = carry_s;
        for (u32 w = 0; w < warp; ++w) off += warp_sum[w];
        off += __popc(ballot & ((1u << lane) - 1u));
        if (k) out[off] = i;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = carry_s;
            for (u32 w = 0; w < LCB_BLOCK / 32; ++w) t += warp_sum[w];
            carry_s = t;
        }
        __syncthreads();
    }
    if (tid == 0) *count = carry_s;
}

// IdentifyBreakpoints' scan (LM/GreedyBreakpointElimination.h:184-221): lab = genome-0 labels in genome-1 order.  The running
// `prev_orient` of the reference is the orientation of the element before, so every position decides on its own.
__global__ void lcb_breakpoints_kernel(const mcu_match* __restrict__ sorted0, const u32* __restrict__ lab, u64 n, u8* __restrict__ is_bp)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 cl = lab[i];
        const mcu_match c = sorted0[cl];
        const bool cur = (c.start1 > 0) == (c.start0 > 0);
        if (i == 0) {
            if (!cur) is_bp[cl] = 1;
        } else {
            const u32 pl = lab[i - 1];
            const mcu_match p = sorted0[pl];
            const bool prev = (p.start1 > 0) == (p.start0 > 0);
            if (!(prev == cur && ((prev && pl + 1 == cl) || (!prev && pl == cl + 1)))) {
                if (prev) is_bp[pl] = 1;
                if (!cur) is_bp[cl] = 1;
            }
        }
        if (i == n - 1 && cur) is_bp[cl] = 1;
        if (i == 0) is_bp[n - 1] = 1;   // breakpoints starts as {size - 1}
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------------------
struct LcbState {
    DevBuf rows_a, rows_b, keys_a, keys_b, idx_a, idx_b, words, head, alive, counters, bp;
    RadixScratch radix;
    cudaStream_t stream = nullptr;
};
static LcbState g_lcb;

void lcb_release()
{
    LcbState& st = g_lcb;
    DevBuf* bufs[] = {&st.rows_a, &st.rows_b, &st.keys_a, &st.keys_b, &st.idx_a, &st.idx_b, &st.words, &st.head, &st.alive, &st.counters, &st.bp,
                      &st.radix.hist, &st.radix.status, &st.radix.counters};
    for (DevBuf* b : bufs) b->release();
    st.radix.epoch = 0;
}

static int lcb_grid(u64 n)
{
    u64 g = div_up(n, 256);
    const u64 cap = (u64)sm_count() * 8;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

// orders `in` (n rows) on |start(seq)| into `out` the way the reference's std::sort would; adds the adjacent ties to *ties
static int lcb_order(LcbState& st, const mcu_match* in, u64 n, int seq, mcu_match* out, u32** idx_out, u64* ties)
{
    cudaStream_t s = st.stream;
    MCU_TRY(st.keys_a.reserve(n * 8)); MCU_TRY(st.keys_b.reserve(n * 8));
    MCU_TRY(st.idx_a.reserve(n * 4)); MCU_TRY(st.idx_b.reserve(n * 4));
    MCU_TRY(st.words.reserve(n * 8));
    MCU_TRY(st.counters.reserve(64));
    unsigned long long* ctr = st.counters.as<unsigned long long>();
    lcb_keys_kernel<<<lcb_grid(n), 256, 0, s>>>(in, n, seq, st.keys_a.as<u64>(), st.idx_a.as<u32>(), st.words.as<u64>());
    bool in_a = true;
    MCU_TRY(radix_sort_pairs<u64>(st.radix, st.keys_a.as<u64>(), st.idx_a.as<u32>(), st.keys_b.as<u64>(), st.idx_b.as<u32>(), n, 34, false, s, &in_a, nullptr));
    const u64* sk = in_a ? st.keys_a.as<u64>() : st.keys_b.as<u64>();
    u32* idx = in_a ? st.idx_a.as<u32>() : st.idx_b.as<u32>();
    MCU_CUDA(cudaMemsetAsync(ctr, 0, 8, s));
    lcb_ties_kernel<<<lcb_grid(n), 256, 0, s>>>(sk, n, ctr);
    unsigned long long t = 0;
    MCU_CUDA(cudaMemcpyAsync(&t, ctr, 8, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    if (t) {   // the reference's order of tied rows is whatever introsort leaves: run it
        lcb_std_sort_kernel<<<1, 1, 0, s>>>(st.words.as<u64>(), n);
        lcb_words_to_idx_kernel<<<lcb_grid(n), 256, 0, s>>>(st.words.as<u64>(), n, idx);
    }
    lcb_gather_kernel<<<lcb_grid(n), 256, 0, s>>>(in, idx, n, out);
    MCU_CUDA(cudaGetLastError());
    if (ties) *ties += t;
    if (idx_out) *idx_out = idx;
    return MCU_OK;
}

static int lcb_init(LcbState& st)
{
    if (!st.stream) MCU_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
    return MCU_OK;
}

int lcb_eliminate_overlaps(const mcu_match* rows, u64 n, int eliminate_both, u64 min_length, mcu_match* rows_out, u64* n_out, u64* ties_out)
{
    LcbState& st = g_lcb;
    if (ties_out) *ties_out = 0;
    *n_out = 0;
    if (n == 0) return MCU_OK;
    if (n >= (1ull << SS_SHIFT)) { set_error("mcu_eliminate_overlaps: too many rows"); return MCU_EINVAL; }
    for (u64 i = 0; i < n; ++i)
        if (rows[i].len <= 0 || rows[i].start0 <= 0 || rows[i].start1 == 0 || rows[i].start0 >= (1ll << 32) || rows[i].start1 >= (1ll << 32) ||
            rows[i].start1 <= -(1ll << 32)) {
            set_error("mcu_eliminate_overlaps: row %llu is not a two-genome match (len > 0, start0 > 0, start1 != 0, |start| < 2^32)", (unsigned long long)i);
            return MCU_EINVAL;
        }
    MCU_TRY(lcb_init(st));
    cudaStream_t s = st.stream;
    MCU_TRY(st.rows_a.reserve(n * sizeof(mcu_match))); MCU_TRY(st.rows_b.reserve(n * sizeof(mcu_match)));
    MCU_TRY(st.head.reserve(n)); MCU_TRY(st.alive.reserve(n)); MCU_TRY(st.counters.reserve(64));
    unsigned long long* ctr = st.counters.as<unsigned long long>();
    mcu_match* cur = st.rows_a.as<mcu_match>();
    mcu_match* oth = st.rows_b.as<mcu_match>();
    MCU_CUDA(cudaMemcpyAsync(cur, rows, n * sizeof(mcu_match), cudaMemcpyHostToDevice, s));
    u64 cnt = n, ties = 0;
    for (int seq = 0; seq < 2; ++seq) {
        const bool last = seq == 1;
        if (cnt >= 2) {
            MCU_TRY(lcb_order(st, cur, cnt, seq, oth, nullptr, &ties));
            lcb_heads_kernel<<<1, LCB_BLOCK, 0, s>>>(oth, cnt, seq, st.head.as<u8>());
            MCU_CUDA(cudaMemsetAsync(st.alive.p, 1, cnt, s));
            lcb_sweep_kernel<<<lcb_grid(cnt), 256, 0, s>>>(oth, cnt, seq, eliminate_both, st.head.as<u8>(), st.alive.as<u8>());
        } else {
            MCU_CUDA(cudaMemcpyAsync(oth, cur, cnt * sizeof(mcu_match), cudaMemcpyDeviceToDevice, s));
            MCU_CUDA(cudaMemsetAsync(st.alive.p, 1, cnt ? cnt : 1, s));
        }
        lcb_compact_rows_kernel<<<1, LCB_BLOCK, 0, s>>>(oth, st.alive.as<u8>(), cnt, last ? min_length : 0, cur, ctr + 1);
        unsigned long long c = 0;
        MCU_CUDA(cudaMemcpyAsync(&c, ctr + 1, 8, cudaMemcpyDeviceToHost, s));
        MCU_CUDA(cudaStreamSynchronize(s));
        MCU_CUDA(cudaGetLastError());
        cnt = c;
        // (a list that shrank below 2 rows in the first pass skips the second sweep like the reference, but not its LengthFilter)
    }
    if (cnt) MCU_CUDA(cudaMemcpyAsync(rows_out, cur, cnt * sizeof(mcu_match), cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    *n_out = cnt;
    if (ties_out) *ties_out = ties;
    return MCU_OK;
}

int lcb_breakpoints(const mcu_match* rows, u64 n, mcu_match* sorted_out, u64* bp_out, u64* n_bp_out, u64* ties_out)
{
    LcbState& st = g_lcb;
    if (ties_out) *ties_out = 0;
    *n_bp_out = 0;
    if (n == 0) return MCU_OK;
    if (n >= (1ull << SS_SHIFT)) { set_error("mcu_lcbs: too many rows"); return MCU_EINVAL; }
    for (u64 i = 0; i < n; ++i)
        if (rows[i].start0 <= 0 || rows[i].start1 == 0 || rows[i].start0 >= (1ll << 32) || rows[i].start1 >= (1ll << 32) || rows[i].start1 <= -(1ll << 32)) {
            set_error("mcu_lcbs: row %llu is not a two-genome match", (unsigned long long)i);
            return MCU_EINVAL;
        }
    MCU_TRY(lcb_init(st));
    cudaStream_t s = st.stream;
    MCU_TRY(st.rows_a.reserve(n * sizeof(mcu_match))); MCU_TRY(st.rows_b.reserve(n * sizeof(mcu_match)));
    MCU_TRY(st.head.reserve(n)); MCU_TRY(st.alive.reserve(n * sizeof(mcu_match))); MCU_TRY(st.counters.reserve(64)); MCU_TRY(st.bp.reserve((n + 1) * 8));
    unsigned long long* ctr = st.counters.as<unsigned long long>();
    mcu_match* in = st.rows_a.as<mcu_match>();
    mcu_match* sorted0 = st.rows_b.as<mcu_match>();
    mcu_match* scratch = st.alive.as<mcu_match>();
    MCU_CUDA(cudaMemcpyAsync(in, rows, n * sizeof(mcu_match), cudaMemcpyHostToDevice, s));
    u64 ties = 0;
    u32* lab = nullptr;
    MCU_TRY(lcb_order(st, in, n, 0, sorted0, nullptr, &ties));
    MCU_TRY(lcb_order(st, sorted0, n, 1, scratch, &lab, &ties));   // lab[i] = genome-0 label of the i-th row in genome-1 order
    MCU_CUDA(cudaMemsetAsync(st.head.p, 0, n, s));
    lcb_breakpoints_kernel<<<lcb_grid(n), 256, 0, s>>>(sorted0, lab, n, st.head.as<u8>());
    lcb_compact_flags_kernel<<<1, LCB_BLOCK, 0, s>>>(st.head.as<u8>(), n, st.bp.as<u64>(), ctr + 2);
    unsigned long long c = 0;
    MCU_CUDA(cudaMemcpyAsync(&c, ctr + 2, 8, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(sorted_out, sorted0, n * sizeof(mcu_match), cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    MCU_CUDA(cudaGetLastError());
    if (c) MCU_CUDA(cudaMemcpyAsync(bp_out, st.bp.p, c * 8, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    *n_bp_out = c;
    if (ties_out) *ties_out = ties;
    return MCU_OK;
}

}  // namespace mcu
