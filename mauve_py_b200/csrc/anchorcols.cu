// Anchor columns of alignment windows on the device: replaces muscle::FindAnchorColsPP (MU/anchoredpp.cpp:354-409)
//   LetterObjScoreXP    MU/anchoredpp.cpp:256-329  (ScoreSeqPairLetters :19-93, per-site ScoreSeqPairGaps :96-250, rows' weights)
//   WindowSmooth        MU/anchors.cpp:9-47
//   FindBestColsComboPP MU/anchoredpp.cpp:335-351
//   MergeBestCols       MU/anchors.cpp:137-186
// One CTA per window; the batch is the grid.  What the reference does column after column is split by what really depends on what:
//   * letter scores: per column;
//   * gap penalties: the reference's running state (bGapping1/2, gap_left_col, cur_gap_score) is reset by every column that holds two
//     letters, so the columns between two such columns form a run that is scored on its own, and without an extension penalty (the DNA
//     setting) the state machine has a closed form per run: two block-wide scans over the columns (ac_score_pair_scans), in the same
//     time whatever the runs look like.  With an extension penalty a run is a float sum in column order: one thread walks each run;
//   * smoothing: ONE running float sum per window whose roundings depend on the order.  Where no addition of the chain rounds, the
//     chain is plain arithmetic and any order gives its values: per segment of 32 columns the prefix sums of (add - sub) are formed by
//     a warp scan in which every addition is tested for exactness (the error term of TwoSum is zero), then total + prefix, - sub,
//     + add are tested the same way, column by column, 32 lanes at once -- FP32 additions only.  All segments of a tile are tested
//     at once against the totals they would start from if nothing before them rounded; the first one that fails is run as the
//     serial chain it always was (two dependent FADDs per column on one lane), which gives the true total behind it, and the
//     segments after it are tested again.  A tile full of rounding columns (gap penalties like -400/3 in the window) is the serial
//     chain outright, and so are the two tiles after it.  Two tile buffers: while one lane sums a tile the other warps store the tile
//     before and fill and prepare the tile after;
//   * best columns: ballots over 4 x 32 consecutive columns per warp and step + a block-wide scan of the warps' counts, which also leaves
//     "best columns in front of column c" for every column;
//   * merging: `next group head` is a function of the head alone (one look-up in that table), the groups are then the walk
//     head -> next head (one lane, a few hundred steps through shared memory), the pick per group in parallel.
// A call with few windows keeps the per-column arrays in dynamic shared memory (CTAs of 1,024 threads, one per SM); a batch runs CTAs of
// 256 threads over global scratch, several per SM.
// Every float operation is the reference's, in the reference's order, in round-to-nearest without contraction (__fadd_rn & co).
// Everything that decides a value sits in functions that also compile for the host (-DMCU_HOST_EMU: the CTA becomes one thread;
// tests/_emu.py, test-only -- the product library has no host path).
#include "anchorcols.cuh"

namespace mcu {

#ifdef MCU_HOST_EMU
#define AC_HD
#define AC_TID 0u
#define AC_NT 1u
#define AC_SYNC() ((void)0)
#define AC_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define AC_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define AC_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define AC_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#define AC_SYNC_OR(x) (x)
#define AC_ATOMIC_MIN(ptr, v) (*(ptr) = *(ptr) < (v) ? *(ptr) : (v))
#define AC_ATOMIC_INC(ptr) ((*(ptr))++)
#define AC_CLOCK() 0ull
#else
#define AC_HD __device__ __forceinline__
#define AC_TID threadIdx.x
#define AC_NT blockDim.x
#define AC_SYNC() __syncthreads()
#define AC_FADD(a, b) __fadd_rn(a, b)
#define AC_FSUB(a, b) __fsub_rn(a, b)
#define AC_FMUL(a, b) __fmul_rn(a, b)
#define AC_FDIV(a, b) __fdiv_rn(a, b)
#define AC_SYNC_OR(x) __syncthreads_or(x)
#define AC_ATOMIC_MIN(ptr, v) atomicMin(ptr, v)
#define AC_ATOMIC_INC(ptr) atomicAdd(ptr, 1u)
#define AC_CLOCK() ((unsigned long long)clock64())
#endif

#define AC_BLOCK 1024
#define AC_TILE 1024   // columns per smoothing tile: 32 segments of 32
#define AC_PHASES 6

struct AcWindow {
    u64 row_off;   // first character of the window's first row
    u64 col_off;   // where the window's per-column outputs / scratch start
    u64 w_off;     // first weight
    u32 ncol, n1, n2, pad;
};

// one tile of the smoothing chain: its operands, results and what the exactness tests need (two of them: one is summed while the
// next is being filled)
struct AcTile {
    float sub[AC_TILE], add[AC_TILE], out[AC_TILE];
    float pre[AC_TILE];            // per segment: inclusive prefix sums of add - sub
    u32 seg_ok[AC_TILE / 32];      // ... every one of them formed without rounding
    float seg_tot[AC_TILE / 32];   // what the segment adds to the running total
};

struct AcShared {
    u8 letter[256];
    AcTile tile[2];                        // (tile[0].sub .. .add also serve as a chunk of `nxt` for the walk over the groups)
    u32 seg_pass[AC_TILE / 32];            // 1: with the total it was tested against, no addition of the segment's chain rounds; 0: one
                                           // does; 2: not tested, the sums in front of it were not exact
    float seg_end[AC_TILE / 32];           // the chain's value behind a segment that passed
    u32 walk_n;
    u32 scan_w[3][AC_BLOCK / 32 + 1];      // ac_scan_min3: the warps' minima
    unsigned long long clk[AC_PHASES + 1];
    u32 seg_fast, seg_chain;               // segments finished without / with the serial chain (reported by mcu_test_anchor_counters)
    float total;
    int first, last;
    u32 warp_sum[AC_BLOCK / 32 + 1];
    u32 nanchor;
};

// the gap penalties of one run of columns that starts at c0 (the first column of the pair's range, or the column after one with two
// letters) and ends before the next column with two letters: ScoreSeqPairGaps :155-250 with its state as it is at a run's start.
AC_HD void ac_gap_run(const u8* r1, const u8* r2, const u8* letter, u32 c0, u32 first, u32 last, u32 ncol, const mcu_anchor_params& p, float* gg)
{
    bool in1 = false, in2 = false;
    u32 left = 0;
    float cur = 0.0f;
    u32 c = c0;
    for (; c <= last; ++c) {
        const bool g1 = letter[r1[c]] == MCU_AC_GAP, g2 = letter[r2[c]] == MCU_AC_GAP;
        if (!g1 && !g2) break;
        if (g1 && g2) continue;
        if (!(g1 ? in1 : in2)) {
            left = c;
            cur = AC_FADD(cur, c == first ? p.term_gap : p.gap_open);
            if (g1) in1 = true;
            else in2 = true;
        } else
            cur = AC_FADD(cur, p.gap_extend);
    }
    if (!in1 && !in2) return;
    u32 end = c;
    if (c > last) {   // open at the pair's last column: a terminal gap, spread to the end of the window (:228-248)
        cur = AC_FSUB(cur, p.gap_open);
        cur = AC_FADD(cur, p.term_gap);
        end = ncol;
    }
    const float per_site = AC_FDIV(cur, (float)(end - left));
    for (u32 k = left; k < end; ++k) gg[k] = per_site;
}

AC_HD float ac_ceil(float x, float ceil_at) { return x > ceil_at ? ceil_at : x; }   // Ceil() of WindowSmooth: both sides are floats widened to double

// s = fl(a + b); true when the addition did not round (the error term of Knuth's TwoSum is zero; overflow gives NaN != 0)
AC_HD bool ac_add_exact(float a, float b, float& s)
{
    s = AC_FADD(a, b);
    const float bv = AC_FSUB(s, a);
    const float err = AC_FADD(AC_FSUB(a, AC_FSUB(s, bv)), AC_FSUB(b, bv));
    return err == 0.0f;
}

#ifndef MCU_HOST_EMU
__device__ __forceinline__ void ac_block_minmax(AcShared& sm, int lo, int hi)
{
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&sm.first, lo);
        atomicMax(&sm.last, hi);
    }
}
// how many of the threads in front of this one raise, in total, `count` (exclusive prefix sum over the CTA's threads), and the CTA's total
// (every thread calls it)
__device__ __forceinline__ u32 ac_block_rank(AcShared& sm, u32 count, u32& total)
{
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    u32 winc = count;
    for (u32 o = 1; o < 32; o <<= 1) {
        const u32 up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
    }
    if (lane == 31) sm.warp_sum[warp] = winc;
    __syncthreads();
    if (warp == 0) {   // the warps' counts -> exclusive prefixes, entry 32 = the total
        const u32 v = lane < nwarps ? sm.warp_sum[lane] : 0;
        u32 inc = v;
        for (u32 o = 1; o < 32; o <<= 1) {
            const u32 up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        sm.warp_sum[lane] = inc - v;
        if (lane == 31) sm.warp_sum[32] = inc;
    }
    __syncthreads();
    total = sm.warp_sum[32];
    return sm.warp_sum[warp] + winc - count;
}
#endif

// the serial chain over columns [k0, k1) of the tile, starting from t: WindowSmooth's loop body (MU/anchors.cpp:38-46), two dependent
// additions per column.  The operands of the next eight columns are in registers before the current eight are added.
AC_HD float ac_chain(AcTile& tl, float t, u32 k0, u32 k1)
{
    u32 k = k0;
#ifndef MCU_HOST_EMU
    // two register sets take turns (no copies between them: a register move costs the chain's pipe as much as an addition)
#define AC_LOAD8(a_, b_, at_)                \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) \
    {                                        \
        a_[j] = tl.sub[(at_) + j];           \
        b_[j] = tl.add[(at_) + j];           \
    }
#define AC_SUM8(a_, b_, at_)                 \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) \
    {                                        \
        tl.out[(at_) + j] = t;               \
        t = AC_FSUB(t, a_[j]);               \
        t = AC_FADD(t, b_[j]);               \
    }
    if (k + 8 <= k1) {
        float a0[8], b0[8], a1[8], b1[8];
        AC_LOAD8(a0, b0, k)
        for (; k + 24 <= k1; k += 16) {
            AC_LOAD8(a1, b1, k + 8)
            AC_SUM8(a0, b0, k)
            AC_LOAD8(a0, b0, k + 16)
            AC_SUM8(a1, b1, k + 8)
        }
        AC_SUM8(a0, b0, k)
        k += 8;
    }
#undef AC_LOAD8
#undef AC_SUM8
#endif
    for (; k < k1; ++k) {
        tl.out[k] = t;
        t = AC_FSUB(t, tl.sub[k]);
        t = AC_FADD(t, tl.add[k]);   // (after the window's last column: a value nobody reads)
    }
    return t;
}

// One segment against the total t0 it would start from: true, its 32 outputs written and the chain's value behind it in seg_end, when
// no addition of the chain -- t - sub, then + add, column after column -- rounds.  With exact prefix sums s_k (seg_ok) the chain's value in
// front of column k is t0 + s_(k-1) provided that sum, the subtraction and the addition that follow are all exact: by induction over k.
// Device: the calling warp, one column per lane.
AC_HD bool ac_verify_segment(AcShared& sm, AcTile& tl, u32 seg, u32 n, float t0)
{
    const u32 base = seg * 32, cnt = n - base < 32 ? n - base : 32;
#ifdef MCU_HOST_EMU
    bool ok = tl.seg_ok[seg] != 0;
    float end = t0;
    for (u32 lane = 0; lane < cnt && ok; ++lane) {
        float P, Q, R;
        ok = ac_add_exact(t0, lane ? tl.pre[base + lane - 1] : 0.0f, P) && ac_add_exact(P, -tl.sub[base + lane], Q) && ac_add_exact(Q, tl.add[base + lane], R);
        end = R;
    }
    if (ok) {
        for (u32 lane = 0; lane < cnt; ++lane) tl.out[base + lane] = AC_FADD(t0, lane ? tl.pre[base + lane - 1] : 0.0f);
        sm.seg_end[seg] = end;
    }
    return ok;
#else
    const u32 lane = threadIdx.x & 31;
    const float mine_pre = tl.pre[base + lane];
    float before = __shfl_up_sync(0xffffffffu, mine_pre, 1);
    if (lane == 0) before = 0.0f;
    float P, Q, R;
    const bool mine = ac_add_exact(t0, before, P) & ac_add_exact(P, -tl.sub[base + lane], Q) & ac_add_exact(Q, tl.add[base + lane], R);
    const bool ok = __all_sync(0xffffffffu, mine || lane >= cnt) && tl.seg_ok[seg] != 0;
    if (ok) {
        if (lane < cnt) tl.out[base + lane] = P;
        if (lane == cnt - 1) sm.seg_end[seg] = R;
    }
    return ok;
#endif
}

// the segments' prefix sums of add - sub and their totals, every addition tested.  Device: warps wid, wid + nw, ... of the callers take
// the segments (all warps of the CTA, or the helper warps while warp 0 is summing the tile before)
AC_HD void ac_tile_prepare(AcTile& tl, u32 n, u32 nseg, u32 wid, u32 nw)
{
#ifdef MCU_HOST_EMU
    for (u32 seg = 0; seg < nseg; ++seg) {
        bool ok = true;
        float run = 0.0f;
        for (u32 lane = 0; lane < 32; ++lane) {
            const u32 k = seg * 32 + lane;
            float d = 0.0f;
            if (k < n) ok = ac_add_exact(tl.add[k], -tl.sub[k], d) && ok;
            ok = ac_add_exact(run, d, run) && ok;
            if (k < AC_TILE) tl.pre[k] = run;
        }
        tl.seg_ok[seg] = ok;
        tl.seg_tot[seg] = run;
    }
#else
    for (u32 seg = wid; seg < nseg; seg += nw) {
        const u32 lane = threadIdx.x & 31, k = seg * 32 + lane;
        float run = 0.0f;
        bool ok = true;
        if (k < n) ok = ac_add_exact(tl.add[k], -tl.sub[k], run);
        for (u32 o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, run, o);
            float sum;
            const bool e = ac_add_exact(run, up, sum);
            if (lane >= o) {
                run = sum;
                ok = ok && e;
            }
        }
        tl.pre[k] = run;
        const bool all = __all_sync(0xffffffffu, ok);
        if (lane == 31) tl.seg_tot[seg] = run;
        if (lane == 0) tl.seg_ok[seg] = all;
    }
#endif
}

// segments whose own sums round (a gap penalty like -400/3 among their operands).  Every warp for itself.
AC_HD u32 ac_tile_bad_segments(const AcTile& tl, u32 nseg)
{
#ifdef MCU_HOST_EMU
    u32 nbad = 0;
    for (u32 seg = 0; seg < nseg; ++seg) nbad += tl.seg_ok[seg] ? 0u : 1u;
    return nbad;
#else
    const u32 lane = threadIdx.x & 31;
    return __popc(__ballot_sync(0xffffffffu, lane < nseg && !tl.seg_ok[lane]));
#endif
}

// The running sum over one prepared tile by the whole CTA (sm.total carried): all segments are tested at once against the totals they would
// start from if nothing before them rounded; the first one that fails is run as the serial chain, which gives the true total behind it, and
// the segments after it are tested again.  True when the tile turned out to be full of rounding columns after all (the caller then
// takes the next tiles as the plain chain).  Every thread calls it; ends with a barrier.
AC_HD bool ac_tile_resolve(AcShared& sm, AcTile& tl, u32 n, u32 nseg)
{
    const u32 tid = AC_TID, nt = AC_NT;
    u32 seg0 = 0;
    for (u32 round = 0; seg0 < nseg; ++round) {
        const float t = sm.total;   // the chain's value in front of segment seg0
        if (round >= 6) {
            AC_SYNC();   // (everybody has read sm.total)
            if (tid == 0) {
                sm.total = ac_chain(tl, t, seg0 * 32, n);
                sm.seg_chain += nseg - seg0;
            }
            AC_SYNC();
            return false;
        }
        // every segment from seg0 on against the total it would start from if none before it rounded
#ifdef MCU_HOST_EMU
        {
            float ts = t;
            bool ts_ok = true;
            for (u32 seg = seg0; seg < nseg; ++seg) {
                sm.seg_pass[seg] = !ts_ok ? 2u : ac_verify_segment(sm, tl, seg, n, ts) ? 1u : 0u;
                ts_ok = ac_add_exact(ts, tl.seg_tot[seg], ts) && ts_ok;
            }
        }
#else
        for (u32 seg = seg0 + (tid >> 5); seg < nseg; seg += nt >> 5) {
            const u32 lane = tid & 31;
            float part = seg0 + lane < seg ? tl.seg_tot[seg0 + lane] : 0.0f;   // nseg <= 32: one lane per earlier segment
            bool exact = true;
            for (u32 o = 16; o; o >>= 1) {
                float sum;
                exact = ac_add_exact(part, __shfl_xor_sync(0xffffffffu, part, o), sum) && exact;
                part = sum;
            }
            float ts;
            exact = ac_add_exact(t, part, ts) && exact;
            const u32 verdict = !__all_sync(0xffffffffu, exact) ? 2u : ac_verify_segment(sm, tl, seg, n, ts) ? 1u : 0u;
            if (lane == 0) sm.seg_pass[seg] = verdict;
        }
#endif
        AC_SYNC();
        u32 f = nseg, nfail = 0;
#ifdef MCU_HOST_EMU
        for (u32 seg = seg0; seg < nseg; ++seg) {
            if (sm.seg_pass[seg] != 1u && f == nseg) f = seg;
            if (sm.seg_pass[seg] == 0u) ++nfail;
        }
#else
        {   // every warp for itself: lane l looks at segment seg0 + l
            const u32 seg = seg0 + (tid & 31);
            const u32 verdict = seg < nseg ? sm.seg_pass[seg] : 1u;
            const u32 open = __ballot_sync(0xffffffffu, verdict != 1u);
            nfail = __popc(__ballot_sync(0xffffffffu, verdict == 0u));
            if (open) f = seg0 + (__ffs(open) - 1);
        }
#endif
        const bool give_up = f < nseg && round == 0 && nfail > 3;
        if (tid == 0) {   // (everybody read sm.total before the barrier above; the flags are written again only after the one below)
            const float tf = f == seg0 ? t : sm.seg_end[f - 1];   // the chain's true value in front of segment f
            sm.seg_fast += f - seg0;
            if (f == nseg)
                sm.total = tf;
            else if (give_up) {   // many: the chain from the first one to the end of the tile
                sm.total = ac_chain(tl, tf, f * 32, n);
                sm.seg_chain += nseg - f;
            } else {
                const u32 k1 = (f + 1) * 32 < n ? (f + 1) * 32 : n;
                sm.total = ac_chain(tl, tf, f * 32, k1);
                sm.seg_chain += 1;
            }
        }
        AC_SYNC();
        if (give_up) return nfail > nseg / 2;
        if (f == nseg) return false;
        seg0 = f + 1;
    }
    return false;
}

// exclusive prefix minimum over the CTA's threads (thread order) of three values at once, continued from the minima of the chunks
// before (carry, kept by every thread): in go the threads' own minima, out come the minima of everything in front of the thread.
// Every thread calls it.
AC_HD void ac_scan_min3(AcShared& sm, u32& a, u32& b, u32& c, u32 carry[3])
{
#ifdef MCU_HOST_EMU
    const u32 ia = a, ib = b, ic = c;
    a = carry[0];
    b = carry[1];
    c = carry[2];
    carry[0] = ia < carry[0] ? ia : carry[0];
    carry[1] = ib < carry[1] ? ib : carry[1];
    carry[2] = ic < carry[2] ? ic : carry[2];
#else
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (u32 o = 1; o < 32; o <<= 1) {
        const u32 ua = __shfl_up_sync(0xffffffffu, a, o), ub = __shfl_up_sync(0xffffffffu, b, o), uc = __shfl_up_sync(0xffffffffu, c, o);
        if (lane >= o) {
            a = min(a, ua);
            b = min(b, ub);
            c = min(c, uc);
        }
    }
    if (lane == 31) {
        sm.scan_w[0][warp] = a;
        sm.scan_w[1][warp] = b;
        sm.scan_w[2][warp] = c;
    }
    a = __shfl_up_sync(0xffffffffu, a, 1);   // inclusive -> exclusive inside the warp
    b = __shfl_up_sync(0xffffffffu, b, 1);
    c = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) a = b = c = 0xffffffffu;
    __syncthreads();
    if (warp < 3) {   // warp v: the minima of the warps in front of every warp for value v; entry 32 = the chunk's minimum
        const u32 mine = lane < nwarps ? sm.scan_w[warp][lane] : 0xffffffffu;
        u32 inc = mine;
        for (u32 o = 1; o < 32; o <<= 1) {
            const u32 up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = min(inc, up);
        }
        u32 exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 0xffffffffu;
        sm.scan_w[warp][lane] = exc;
        if (lane == 31) sm.scan_w[warp][32] = inc;
    }
    __syncthreads();
    a = min(min(a, sm.scan_w[0][warp]), carry[0]);
    b = min(min(b, sm.scan_w[1][warp]), carry[1]);
    c = min(min(c, sm.scan_w[2][warp]), carry[2]);
    carry[0] = min(carry[0], sm.scan_w[0][32]);
    carry[1] = min(carry[1], sm.scan_w[1][32]);
    carry[2] = min(carry[2], sm.scan_w[2][32]);
    __syncthreads();
#endif
}

// first / last column where not both rows have a gap (MU/anchoredpp.cpp:46-74); the whole window when there is none
AC_HD void ac_pair_range(AcShared& sm, const u8* code, u32 L, u32& first, u32& last)
{
    const u32 tid = AC_TID, nt = AC_NT;
    if (tid == 0) {
        sm.first = (int)L;
        sm.last = -1;
    }
    AC_SYNC();
    int lo = (int)L, hi = -1;
    for (u32 c = tid; c < L; c += nt)
        if (code[c] != 0x2d) {   // not gap | gap
            lo = lo < (int)c ? lo : (int)c;
            hi = (int)c;
        }
#ifdef MCU_HOST_EMU
    sm.first = lo;
    sm.last = hi;
#else
    ac_block_minmax(sm, lo, hi);
#endif
    AC_SYNC();
    first = sm.last < 0 ? 0u : (u32)sm.first;
    last = sm.last < 0 ? L - 1 : (u32)sm.last;
}

#define AC_NONE 0xffffffffu
#define AC_V 4   // columns per thread and chunk in the two scans over a pair of rows
// a column's pair of letters in six bits: 0..3 residues, 4 a letter outside the alphabet, 5 a gap; row 2 in bits 3..5
AC_HD u32 ac_code(const u8* letter, u8 c1, u8 c2)
{
    const u32 a = letter[c1], b = letter[c2];
    return (a < 4 ? a : a == MCU_AC_GAP ? 5u : 4u) | ((b < 4 ? b : b == MCU_AC_GAP ? 5u : 4u) << 3);
}
#define AC_G1(k) (((k) & 7u) == 5u)
#define AC_G2(k) (((k) >> 3) == 5u)

// One pair of rows, no extension penalty (the setting in force for DNA): score[c] += ww * (letters + gap penalties), with the reference's
// gap state machine (:155-250) in closed form.  Its state is reset by every column with two letters, so a run of columns between two such
// columns is scored on its own: each row's FIRST gap column of the run opens a gap -- the first column of the pair's range at the
// terminal price, :172 / :198 --, every other gap column adds gap_extend = 0, and the penalty is spread from the LATER of the two opening
// columns (gap_left_col is overwritten, :170 / :196) to the run's end, or to the end of the window when the run is still open at the
// pair's last column (:228-248).  Two scans over the columns do it whatever the runs look like: backwards, the next column with two
// letters / with a gap in row 1 only / in row 2 only (at a run's first column that is the whole run: its three numbers go to
// d_left / d_end / d_per there); forwards, the run a column belongs to.
// code: one byte per column (scratch); d_left, d_end, d_per: one word per column (scratch, only the runs' first columns are used)
AC_HD void ac_score_pair_scans(AcShared& sm, const u8* __restrict__ r1, const u8* __restrict__ r2, u32 L, float ww, const mcu_anchor_params& p,
                               float* score, u8* code, u32* __restrict__ d_left, u32* __restrict__ d_end, u32* __restrict__ d_per)
{
    const u32 tid = AC_TID, nt = AC_NT;
    const u8* letter = sm.letter;
#pragma unroll 4
    for (u32 c = tid; c < L; c += nt) code[c] = (u8)ac_code(letter, r1[c], r2[c]);
    AC_SYNC();
    u32 first, last;
    ac_pair_range(sm, code, L, first, last);
    // Four columns per thread and chunk (AC_V): the thread's own columns are scanned in registers, the threads' minima by the CTA.
    // backwards: thread t of a chunk looks at columns hi - 1 - (4 t + j), so that a prefix minimum in (t, j) order is a suffix minimum
    // over columns
    {
        u32 carry[3] = {AC_NONE, AC_NONE, AC_NONE};
        for (u32 hi = last + 1; hi > first; hi = hi > AC_V * nt ? hi - AC_V * nt : 0) {
            u32 col[AC_V], kk[AC_V], m_two[AC_V], m_g1[AC_V], m_g2[AC_V];
            u32 r_two = AC_NONE, r_g1 = AC_NONE, r_g2 = AC_NONE;
#pragma unroll
            for (int j = 0; j < AC_V; ++j) {
                const u32 idx = AC_V * tid + j;
                const bool valid = idx < hi && hi - 1 - idx >= first;
                const u32 c = valid ? hi - 1 - idx : AC_NONE;
                const u32 k = valid ? code[c] : 0;
                const bool g1 = AC_G1(k), g2 = AC_G2(k);
                col[j] = c;
                kk[j] = k;
                if (valid && !g1 && !g2) r_two = c;        // (columns descend: a later one is the smaller index)
                if (valid && g1 && !g2) r_g1 = c;
                if (valid && g2 && !g1) r_g2 = c;
                m_two[j] = r_two;
                m_g1[j] = r_g1;
                m_g2[j] = r_g2;
            }
            ac_scan_min3(sm, r_two, r_g1, r_g2, carry);   // now: the minima over everything behind this thread's columns
#pragma unroll
            for (int j = 0; j < AC_V; ++j) {
                const u32 c = col[j];
                if (c == AC_NONE) continue;
                const u32 k = kk[j];
                if (!AC_G1(k) && !AC_G2(k)) continue;
                if (c != first) {   // a run starts behind a column with two letters
                    const u32 kb = code[c - 1];
                    if (AC_G1(kb) || AC_G2(kb)) continue;
                }
                const u32 n_two = m_two[j] < r_two ? m_two[j] : r_two, n_g1 = m_g1[j] < r_g1 ? m_g1[j] : r_g1,
                          n_g2 = m_g2[j] < r_g2 ? m_g2[j] : r_g2;
                const u32 e = n_two < last + 1 ? n_two : last + 1;
                const u32 f1 = n_g1 < e ? n_g1 : AC_NONE, f2 = n_g2 < e ? n_g2 : AC_NONE;
                u32 left = AC_NONE, end = e;
                float per_site = 0.0f;
                if (f1 != AC_NONE || f2 != AC_NONE) {
                    const u32 fa = f1 < f2 ? f1 : f2, fb = f1 < f2 ? f2 : f1;   // fb none: only one row has gaps in this run
                    float cur = AC_FADD(0.0f, fa == first ? p.term_gap : p.gap_open);
                    left = fa;
                    if (fb != AC_NONE) {
                        cur = AC_FADD(cur, p.gap_open);
                        left = fb;
                    }
                    if (e > last) {
                        cur = AC_FSUB(cur, p.gap_open);
                        cur = AC_FADD(cur, p.term_gap);
                        end = L;
                    }
                    per_site = AC_FDIV(cur, (float)(end - left));
                }
                d_left[c] = left;
                d_end[c] = end;
#ifdef MCU_HOST_EMU
                memcpy(&d_per[c], &per_site, 4);
#else
                d_per[c] = __float_as_uint(per_site);
#endif
            }
            if (hi <= AC_V * nt) break;
        }
    }
    AC_SYNC();
    // forwards: the run's first column in front of (or at) every column, then the column's share of the run's penalty
    {
        u32 carry[3] = {AC_NONE, AC_NONE, AC_NONE};
        for (u32 c0 = 0; c0 < L; c0 += AC_V * nt) {
            u32 kk[AC_V], hh[AC_V];
            u32 r_h = AC_NONE, x1 = AC_NONE, x2 = AC_NONE;
#pragma unroll
            for (int j = 0; j < AC_V; ++j) {
                const u32 c = c0 + AC_V * tid + j;
                const bool valid = c < L;
                const u32 k = valid ? code[c] : 0;
                kk[j] = k;
                bool head = valid && (AC_G1(k) || AC_G2(k)) && c >= first && c <= last;
                if (head && c != first) {
                    const u32 kb = j ? kk[j - 1] : code[c - 1];
                    head = !AC_G1(kb) && !AC_G2(kb);
                }
                // (the latest first-column-of-a-run at or in front of c: the largest one, as a minimum of complements.  A column that
                //  lies behind that run's end -- one with two letters, or all-gap columns behind the pair's range -- gets nothing from it)
                if (head) r_h = ~(c + 1);
                hh[j] = r_h;
            }
            ac_scan_min3(sm, r_h, x1, x2, carry);
#pragma unroll
            for (int j = 0; j < AC_V; ++j) {
                const u32 c = c0 + AC_V * tid + j;
                if (c >= L) continue;
                const u32 h = hh[j] < r_h ? hh[j] : r_h, k = kk[j];
                float g = 0.0f;
                if (h != AC_NONE) {
                    const u32 hc = ~h - 1;
                    const u32 left = d_left[hc];
                    if (left != AC_NONE && c >= left && c < d_end[hc]) {
#ifdef MCU_HOST_EMU
                        memcpy(&g, &d_per[hc], 4);
#else
                        g = __uint_as_float(d_per[hc]);
#endif
                    }
                }
                const u32 a = k & 7u, b = k >> 3;
                const float mm = (a < 4 && b < 4) ? p.subst[a][b] : 0.0f;
                score[c] = AC_FADD(score[c], AC_FMUL(ww, AC_FADD(mm, g)));
            }
        }
    }
    AC_SYNC();
}

// One pair of rows with an extension penalty: every run's penalty is a float sum in column order, walked by the thread that owns the
// run's first column (ac_gap_run).  gg: one float per column, zero on entry and on exit.
AC_HD void ac_score_pair_walks(AcShared& sm, const u8* __restrict__ r1, const u8* __restrict__ r2, u32 L, float ww, const mcu_anchor_params& p,
                               float* score, float* gg)
{
    const u32 tid = AC_TID, nt = AC_NT;
    const u8* letter = sm.letter;
    if (tid == 0) {
        sm.first = (int)L;
        sm.last = -1;
    }
    AC_SYNC();
    {
        int lo = (int)L, hi = -1;
        for (u32 c = tid; c < L; c += nt)
            if (letter[r1[c]] != MCU_AC_GAP || letter[r2[c]] != MCU_AC_GAP) {
                lo = lo < (int)c ? lo : (int)c;
                hi = (int)c;
            }
#ifdef MCU_HOST_EMU
        sm.first = lo;
        sm.last = hi;
#else
        ac_block_minmax(sm, lo, hi);
#endif
    }
    AC_SYNC();
    const u32 first = sm.last < 0 ? 0u : (u32)sm.first, last = sm.last < 0 ? L - 1 : (u32)sm.last;
    for (u32 c = first + tid; c <= last; c += nt) {
        const bool two = letter[r1[c]] != MCU_AC_GAP && letter[r2[c]] != MCU_AC_GAP;
        if (two) continue;
        if (c != first) {
            const bool two_before = letter[r1[c - 1]] != MCU_AC_GAP && letter[r2[c - 1]] != MCU_AC_GAP;
            if (!two_before) continue;
        }
        ac_gap_run(r1, r2, letter, c, first, last, L, p, gg);
    }
    AC_SYNC();
    for (u32 c = tid; c < L; c += nt) {
        const u32 a = letter[r1[c]], b = letter[r2[c]];
        const float mm = (a < 4 && b < 4) ? p.subst[a][b] : 0.0f;   // outside [first, last] both rows have gaps: 0 as well
        score[c] = AC_FADD(score[c], AC_FMUL(ww, AC_FADD(mm, gg[c])));
        gg[c] = 0.0f;
    }
    AC_SYNC();
}

// (smooth may be the same memory as gg: the gap penalties are dead when the smoothing starts; `before` must be neither)
AC_HD void ac_window(AcShared& sm, const AcWindow w, const u8* __restrict__ rows_all, const float* __restrict__ weights, const mcu_anchor_params& p,
                     float* score, float* smooth, float* gg, u32* before, u32* __restrict__ best, u32* __restrict__ nxt,
                     u32* __restrict__ heads, u32* __restrict__ cols_out, u32* __restrict__ count_out)
{
    const u32 tid = AC_TID, nt = AC_NT;
    const u32 L = w.ncol;
    const u8* rows = rows_all + w.row_off;
    for (u32 i = tid; i < 256; i += nt) sm.letter[i] = p.letter_of_char[i];
#pragma unroll 4
    for (u32 c = tid; c < L; c += nt) {
        score[c] = 0.0f;
        gg[c] = 0.0f;
    }
    if (tid == 0) {
        sm.nanchor = 0;
        sm.seg_fast = 0;
        sm.seg_chain = 0;
        sm.clk[0] = AC_CLOCK();
    }
    AC_SYNC();
    const u8* letter = sm.letter;

    // ---- LetterObjScoreXP: the pairs in the reference's order, every column accumulating in that order
    for (u32 i = 0; i < w.n1; ++i)
        for (u32 j = 0; j < w.n2; ++j) {
            const u8* r1 = rows + (u64)i * L;
            const u8* r2 = rows + (u64)(w.n1 + j) * L;
            const float w1 = weights ? weights[w.w_off + i] : 1.0f, w2 = weights ? weights[w.w_off + w.n1 + j] : 1.0f;
            const float ww = AC_FMUL(w1, w2);
            if (p.gap_extend == 0.0f) ac_score_pair_scans(sm, r1, r2, L, ww, p, score, (u8*)gg, best, nxt, heads);
            else ac_score_pair_walks(sm, r1, r2, L, ww, p, score, gg);
        }

    if (tid == 0) sm.clk[1] = AC_CLOCK();

    // ---- WindowSmooth
    const u32 W = p.smooth_window, w2 = W / 2;
    if (L <= W) {
        for (u32 c = tid; c < L; c += nt) smooth[c] = 0.0f;
    } else {
        for (u32 c = tid; c < w2; c += nt) {   // the window's edges (MU/anchors.cpp:25-29)
            smooth[c] = 0.0f;
            smooth[L - c - 1] = 0.0f;
        }
        if (tid == 0) {
            float t = 0.0f;
            for (u32 i = 0; i < W; ++i) t = AC_FADD(t, ac_ceil(score[i], p.smooth_ceil));
            sm.total = t;
        }
        const u32 i_last = L - w2 - 1;
        const float fw = (float)W;
        const u32 ntiles = (i_last - w2 + 1 + AC_TILE - 1) / AC_TILE;
        // Tiles of AC_TILE columns, two buffers.  While one lane sums tile k (the serial chain), the other warps divide and store tile
        // k - 1 and fill and prepare tile k + 1: one barrier per tile, the chain is all that is left on the critical path.  A tile that
        // may hold exact segments is resolved by the whole CTA instead (ac_tile_resolve).
#ifdef MCU_HOST_EMU
        const u32 hid = 0, nh = 1, hwid = 0, hnw = 1;
        const bool chain_role = true, helper_role = true;
#else
        const u32 hid = tid - 32, nh = nt - 32, hwid = (tid >> 5) - 1, hnw = (nt >> 5) - 1;   // helpers: every warp but the first
        const bool chain_role = tid < 32, helper_role = tid >= 32;
#endif
#define AC_TILE_N(k_) ((i_last - (w2 + (k_) * AC_TILE) + 1) < AC_TILE ? (i_last - (w2 + (k_) * AC_TILE) + 1) : AC_TILE)
#define AC_TILE_FILL(nx, k_, first_, step_)                                                        \
    {                                                                                              \
        const u32 i0_ = w2 + (k_) * AC_TILE, n_ = AC_TILE_N(k_);                                   \
        for (u32 q = (first_); q < n_; q += (step_)) {                                             \
            const u32 i = i0_ + q;                                                                 \
            (nx).sub[q] = ac_ceil(score[i - w2], p.smooth_ceil);                                   \
            (nx).add[q] = i + w2 + 1 < L ? ac_ceil(score[i + w2 + 1], p.smooth_ceil) : 0.0f; /* unused at i_last */ \
        }                                                                                          \
    }
#define AC_TILE_STORE(pv, k_, first_, step_)                                                       \
    {                                                                                              \
        const u32 i0_ = w2 + (k_) * AC_TILE, n_ = AC_TILE_N(k_);                                   \
        for (u32 q = (first_); q < n_; q += (step_)) smooth[i0_ + q] = AC_FDIV((pv).out[q], fw);   \
    }
        AC_TILE_FILL(sm.tile[0], 0u, tid, nt)
        AC_SYNC();
        ac_tile_prepare(sm.tile[0], AC_TILE_N(0u), (AC_TILE_N(0u) + 31) / 32, tid >> 5, nt >> 5 ? nt >> 5 : 1);
        AC_SYNC();
        u32 hold = 0;   // tiles left to take as the plain chain after a tile full of rounding columns (same value in every thread)
        // tile k in `tl`; tiles k - 1 and k + 1 in `ot` (called with the two buffers by name, so that the chain's addresses are constants)
        auto one_tile = [&](const u32 k, AcTile& tl, AcTile& ot) {
            const u32 n = AC_TILE_N(k), nseg = (n + 31) / 32;
            const u32 nbad = hold ? nseg : ac_tile_bad_segments(tl, nseg);
            if (hold || nbad > 3) {
                if (hold) --hold;
                else if (nbad > nseg / 2) hold = 2;
                if (chain_role && (tid & 31) == 0) {
                    sm.total = ac_chain(tl, sm.total, 0, n);
                    sm.seg_chain += nseg;
                }
                if (helper_role) {
                    if (k > 0) AC_TILE_STORE(ot, k - 1, hid, nh)
                    if (k + 1 < ntiles) {
                        AC_TILE_FILL(ot, k + 1, hid, nh)
                        if (!hold) {   // (a tile that will be taken as the chain anyway needs no preparation)
#ifndef MCU_HOST_EMU
                            asm volatile("bar.sync 1, %0;" ::"r"(nh) : "memory");   // the helpers among themselves: the tile is filled
#endif
                            ac_tile_prepare(ot, AC_TILE_N(k + 1), (AC_TILE_N(k + 1) + 31) / 32, hwid, hnw);
                        }
                    }
                }
                AC_SYNC();
            } else {
                if (k > 0) AC_TILE_STORE(ot, k - 1, tid, nt)
                if (k + 1 < ntiles) {
                    AC_TILE_FILL(ot, k + 1, tid, nt)
                    AC_SYNC();
                    ac_tile_prepare(ot, AC_TILE_N(k + 1), (AC_TILE_N(k + 1) + 31) / 32, tid >> 5, nt >> 5 ? nt >> 5 : 1);
                }
                if (ac_tile_resolve(sm, tl, n, nseg)) hold = 2;
            }
        };
        for (u32 k = 0; k < ntiles; k += 2) {
            one_tile(k, sm.tile[0], sm.tile[1]);
            if (k + 1 < ntiles) one_tile(k + 1, sm.tile[1], sm.tile[0]);
        }
        if ((ntiles - 1) & 1) AC_TILE_STORE(sm.tile[1], ntiles - 1, tid, nt)
        else AC_TILE_STORE(sm.tile[0], ntiles - 1, tid, nt)
#undef AC_TILE_N
#undef AC_TILE_FILL
#undef AC_TILE_STORE
    }
    AC_SYNC();

    if (tid == 0) sm.clk[2] = AC_CLOCK();

    // ---- FindBestColsComboPP: the columns that pass both thresholds, in order; before[c] = how many of them lie in front of column c
    u32 nbest = 0;
#ifdef MCU_HOST_EMU
    for (u32 c = 0; c < L; ++c) {
        before[c] = nbest;
        if (!(score[c] < p.min_best_col) && !(smooth[c] < p.min_smooth)) best[nbest++] = c;
    }
#else
    for (u32 c0 = 0; c0 < L; c0 += AC_V * nt) {   // per step every warp takes 4 x 32 consecutive columns (one ballot per 32)
        const u32 lane = tid & 31, wbase = c0 + (tid >> 5) * (AC_V * 32);
        u32 b[AC_V], wtot = 0;
#pragma unroll
        for (int j = 0; j < AC_V; ++j) {
            const u32 c = wbase + j * 32 + lane;
            b[j] = __ballot_sync(0xffffffffu, c < L && !(score[c] < p.min_best_col) && !(smooth[c] < p.min_smooth));
            wtot += __popc(b[j]);
        }
        u32 total;
        u32 rank = ac_block_rank(sm, lane == 0 ? wtot : 0u, total);   // lane 0: the best columns of the warps in front
        rank = nbest + __shfl_sync(0xffffffffu, rank, 0);
#pragma unroll
        for (int j = 0; j < AC_V; ++j) {
            const u32 c = wbase + j * 32 + lane;
            const u32 mine = rank + __popc(b[j] & ((1u << lane) - 1u));
            if (c < L) before[c] = mine;
            if (b[j] & (1u << lane)) best[mine] = c;
            rank += __popc(b[j]);
        }
        nbest += total;
        __syncthreads();
    }
#endif
    AC_SYNC();

    if (tid == 0) sm.clk[3] = AC_CLOCK();

    // ---- MergeBestCols: groups of best columns closer to the group's first one than the spacing.  The group that starts at best[n]
    //      ends in front of the first best column at or beyond best[n] + spacing: one look-up in `before`
#define AC_GROUP_END(n_) ((u64)best[n_] + p.anchor_spacing < L ? before[(u64)best[n_] + p.anchor_spacing] : nbest)
    if (tid == 0) sm.clk[4] = AC_CLOCK();
    {   // the walk head -> next head, one lane, through chunks of group ends in shared memory (the two smoothing tiles' memory)
        u32* s_nxt = (u32*)&sm.tile[0];
        const u32 CH = (u32)(2 * sizeof(AcTile) / sizeof(u32));
        if (tid == 0) sm.walk_n = 0;
        AC_SYNC();
        for (;;) {
            const u32 c0 = sm.walk_n;
            if (c0 >= nbest) break;
            const u32 cn = nbest - c0 < CH ? nbest - c0 : CH;
            for (u32 k0 = 0; k0 < cn; k0 += 8 * nt) {   // eight look-ups per thread in flight
                u32 stop[8], end[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const u32 k = k0 + j * nt + tid;
                    stop[j] = k < cn ? best[c0 + k] + p.anchor_spacing : 0xffffffffu;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) end[j] = stop[j] < L ? before[stop[j]] : nbest;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const u32 k = k0 + j * nt + tid;
                    if (k < cn) s_nxt[k] = end[j];
                }
            }
            AC_SYNC();
            if (tid == 0) {
                u32 n = c0, k = sm.nanchor;
                while (n < c0 + cn) {
                    heads[k++] = n;
                    n = s_nxt[n - c0];
                }
                sm.nanchor = k;
                sm.walk_n = n;
            }
            AC_SYNC();
        }
    }
    AC_SYNC();
    if (tid == 0) sm.clk[5] = AC_CLOCK();
    const u32 nanchor = sm.nanchor;
    for (u32 k = tid; k < nanchor; k += nt) {
        const u32 n = heads[k], within = AC_GROUP_END(n) - n - 1, head = best[n];
        u32 pick = head;
        if (within == 1) {
            const u32 other = best[n + 1];
            pick = score[head] > score[other] ? head : other;
        } else if (within > 1) {
            // "closest to the centre" as the reference computes it (MU/anchors.cpp:164-179): the distance is taken from the group's FIRST
            // column, over the members but the first and the last, and the first smaller one wins -- in an ascending list that is the
            // second member, whose distance is below the spacing by the definition of the group
            pick = best[n + 1];
        }
        cols_out[k] = pick;
    }
    if (tid == 0) {
        *count_out = nanchor;
        sm.clk[6] = AC_CLOCK();
    }
#undef AC_GROUP_END
}

#ifndef MCU_HOST_EMU

// smem_cols > 0: a window of up to that many columns keeps its per-column scores and gap penalties / smoothed scores in (dynamic) shared
// memory -- the latency form, one CTA per SM, for calls with few windows; they reach global memory only when the caller wants them.
__global__ __launch_bounds__(AC_BLOCK) void anchor_cols_kernel(const AcWindow* __restrict__ windows, const u8* __restrict__ rows,
                                                               const float* __restrict__ weights, const mcu_anchor_params p, float* score, float* smooth,
                                                               float* gg, u32* best, u32* nxt, u32* heads, u32* cols_out, u32* counts,
                                                               unsigned long long* seg_counters, u32 smem_cols, int want_scores)
{
    __shared__ AcShared sm;
    extern __shared__ float ac_dyn[];
    const AcWindow w = windows[blockIdx.x];
    if (w.ncol == 0) {
        if (threadIdx.x == 0) counts[blockIdx.x] = 0;
        return;
    }
    const bool in_smem = w.ncol <= smem_cols;
    float* my_score = in_smem ? ac_dyn : score + w.col_off;
    float* my_gg = in_smem ? ac_dyn + smem_cols : gg + w.col_off;
    float* my_smooth = in_smem ? my_gg : smooth + w.col_off;
    u32* before = in_smem ? (u32*)(gg + w.col_off) : (u32*)my_gg;   // global mode: the gap penalties' memory once they are dead
    ac_window(sm, w, rows, weights, p, my_score, my_smooth, my_gg, before, best + w.col_off, nxt + w.col_off, heads + w.col_off, cols_out + w.col_off,
              counts + blockIdx.x);
    if (in_smem && want_scores) {
        __syncthreads();
        for (u32 c = threadIdx.x; c < w.ncol; c += blockDim.x) {
            score[w.col_off + c] = my_score[c];
            smooth[w.col_off + c] = my_smooth[c];
        }
    }
    if (threadIdx.x == 0) {
        atomicAdd(seg_counters, (unsigned long long)sm.seg_fast);
        atomicAdd(seg_counters + 1, (unsigned long long)sm.seg_chain);
        if (blockIdx.x == 0)   // SM cycles of the first window's phases: scoring, smoothing, best columns, group search, group walk, picks
            for (int i = 0; i < AC_PHASES; ++i) seg_counters[2 + i] = sm.clk[i + 1] - sm.clk[i];
    }
}

// the anchor columns of all windows packed one after the other (a batch's result is a few columns per thousand: only those travel back)
__global__ void anchor_cols_pack_kernel(const AcWindow* __restrict__ windows, const u32* __restrict__ cols, const u32* __restrict__ counts,
                                        const u64* __restrict__ dense_off, u32* __restrict__ dense)
{
    const AcWindow w = windows[blockIdx.x];
    const u32 n = counts[blockIdx.x];
    const u64 to = dense_off[blockIdx.x];
    for (u32 k = threadIdx.x; k < n; k += blockDim.x) dense[to + k] = cols[w.col_off + k];
}

#define AC_STAGE_BYTES (1u << 20)
#define AC_COUNTER_BYTES 64u
struct AcState {
    DevBuf windows, rows, weights, score, smooth, gg, best, nxt, heads, cols, counts, dense, dense_off, seg_counters, in_blob, out_blob;
    u8 *h_in = nullptr, *h_out = nullptr;   // pinned staging blocks of the small-call path
    u64 last_counters[2 + AC_PHASES] = {0};
    size_t max_dyn_smem = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
static AcState g_ac;

void ac_release()
{
    AcState& st = g_ac;
    DevBuf* bufs[] = {&st.windows, &st.rows, &st.weights, &st.score, &st.smooth, &st.gg, &st.best, &st.nxt, &st.heads, &st.cols, &st.counts,
                      &st.dense, &st.dense_off, &st.seg_counters, &st.in_blob, &st.out_blob};
    for (DevBuf* b : bufs) b->release();
    if (st.h_in) cudaFreeHost(st.h_in);
    if (st.h_out) cudaFreeHost(st.h_out);
    st.h_in = st.h_out = nullptr;
}

void ac_last_counters(u64* out8)
{
    for (int i = 0; i < 2 + AC_PHASES; ++i) out8[i] = g_ac.last_counters[i];
}

void ac_default_params(mcu_anchor_params* p)
{
    // MuscleInterface::ProfileAlignFast's set-up (LM/MuscleInterface.cpp:1086-1106): SetAlpha(ALPHA_DNA) MU/alpha.cpp:123-141,
    // SetPPScore(PPSCORE_SPN) -> SetDefaultsSPN_DNA MU/params.cpp:296-313 with NUC_SP MU/nucmx.cpp:8-25, TERMGAPS_Half MU/params.cpp:138
    static const int nuc[4][4] = {{91, -114, -31, -123}, {-114, 100, -125, -31}, {-31, -125, 100, -114}, {-123, -31, -114, 91}};
    memset(p, 0, sizeof *p);
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) p->subst[a][b] = (float)nuc[a][b] + 60.0f;
    p->gap_open = -400.0f;
    p->gap_extend = 0.0f;
    p->term_gap = -200.0f;
    p->smooth_ceil = 999.0f;
    p->min_best_col = 90.0f;
    p->min_smooth = 90.0f;
    p->smooth_window = 21;
    p->anchor_spacing = 96;
    memset(p->letter_of_char, 0xfe, 256);
    const char* res = "ACGT";
    for (int a = 0; a < 4; ++a) p->letter_of_char[(u8)res[a]] = p->letter_of_char[(u8)(res[a] | 0x20)] = (u8)a;
    p->letter_of_char[(u8)'U'] = p->letter_of_char[(u8)'u'] = 3;
    const char* wild = "MRWSYKVHDBXN";
    for (int a = 0; wild[a]; ++a) p->letter_of_char[(u8)wild[a]] = p->letter_of_char[(u8)(wild[a] | 0x20)] = (u8)(4 + a);
    p->letter_of_char[(u8)'-'] = p->letter_of_char[(u8)'.'] = MCU_AC_GAP;
}

int ac_batch(u64 n, const char* rows, const u64* row_off, const u32* ncol, const u32* n1, const u32* n2, const float* weights,
             const mcu_anchor_params* params, const u64* col_off, u32* cols_out, u32* n_cols_out, float* score_out, float* smooth_out, float* device_ms)
{
    AcState& st = g_ac;
    if (device_ms) *device_ms = 0.0f;
    if (n == 0) return MCU_OK;
    mcu_anchor_params p;
    if (params) p = *params;
    else ac_default_params(&p);
    if (p.smooth_window % 2 != 1) { set_error("mcu_anchor_cols_batch: the smoothing window must be odd (WindowSmooth, MU/anchors.cpp:14-15)"); return MCU_EINVAL; }
    if (p.anchor_spacing == 0 || p.anchor_spacing >= (1u << 30)) { set_error("mcu_anchor_cols_batch: anchor spacing out of range"); return MCU_EINVAL; }
    if (n >= (1ull << 31)) { set_error("mcu_anchor_cols_batch: too many windows"); return MCU_EINVAL; }
    std::vector<AcWindow> win(n);
    u64 n_weights = 0, rows_bytes = 0, cols_total = 0;
    for (u64 i = 0; i < n; ++i) {
        if (n1[i] == 0 || n2[i] == 0) { set_error("mcu_anchor_cols_batch: window %llu has an alignment without rows", (unsigned long long)i); return MCU_EINVAL; }
        if (ncol[i] >= (1u << 31)) { set_error("mcu_anchor_cols_batch: window %llu is too long", (unsigned long long)i); return MCU_EINVAL; }
        if (col_off[i + 1] < col_off[i] + ncol[i]) { set_error("mcu_anchor_cols_batch: col_off leaves window %llu fewer than ncol entries", (unsigned long long)i); return MCU_EINVAL; }
        AcWindow& w = win[i];
        w.row_off = row_off[i];
        w.col_off = col_off[i];
        w.w_off = n_weights;
        w.ncol = ncol[i];
        w.n1 = n1[i];
        w.n2 = n2[i];
        w.pad = 0;
        n_weights += (u64)n1[i] + n2[i];
        const u64 end = row_off[i] + ((u64)n1[i] + n2[i]) * ncol[i];
        rows_bytes = end > rows_bytes ? end : rows_bytes;
    }
    cols_total = col_off[n];
    if (!st.stream) {
        int dev = 0, optin = 0;
        MCU_CUDA(cudaGetDevice(&dev));
        MCU_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cudaFuncAttributes fa;
        MCU_CUDA(cudaFuncGetAttributes(&fa, anchor_cols_kernel));
        const long room = (long)optin - (long)fa.sharedSizeBytes - 1024;
        st.max_dyn_smem = room > 0 ? (size_t)room : 0;
        MCU_CUDA(cudaFuncSetAttribute(anchor_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st.max_dyn_smem));
        MCU_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        MCU_CUDA(cudaEventCreate(&st.ev0));
        MCU_CUDA(cudaEventCreate(&st.ev1));
    }
    cudaStream_t s = st.stream;
    const u64 ct = cols_total ? cols_total : 1;
    MCU_TRY(st.score.reserve(ct * 4)); MCU_TRY(st.smooth.reserve(ct * 4)); MCU_TRY(st.gg.reserve(ct * 4));
    MCU_TRY(st.best.reserve(ct * 4)); MCU_TRY(st.nxt.reserve(ct * 4)); MCU_TRY(st.heads.reserve(ct * 4));
    // One window per call is the aligner's case, and there every transfer counts: a call whose inputs and outputs are small goes through
    // two pinned staging blocks -- windows, weights and rows down in ONE copy, counters, counts and columns back in ONE copy
    const u64 in_w = (n * sizeof(AcWindow) + 15) & ~15ull, in_f = (n_weights * sizeof(float) + 15) & ~15ull;
    const u64 in_bytes = in_w + in_f + rows_bytes;
    const u64 out_c = (n * 4 + 15) & ~15ull, out_bytes = AC_COUNTER_BYTES + out_c + ct * 4;
    const bool staged = in_bytes <= AC_STAGE_BYTES && out_bytes <= AC_STAGE_BYTES;
    const AcWindow* d_windows;
    const u8* d_rows;
    const float* d_weights;
    u32 *d_counts, *d_cols;
    unsigned long long* d_counters;
    if (staged) {
        if (!st.h_in) {
            MCU_CUDA(cudaMallocHost(&st.h_in, AC_STAGE_BYTES));
            MCU_CUDA(cudaMallocHost(&st.h_out, AC_STAGE_BYTES));
        }
        MCU_TRY(st.in_blob.reserve(AC_STAGE_BYTES));
        MCU_TRY(st.out_blob.reserve(AC_STAGE_BYTES));
        memcpy(st.h_in, win.data(), n * sizeof(AcWindow));
        if (weights) memcpy(st.h_in + in_w, weights, n_weights * sizeof(float));
        if (rows_bytes) memcpy(st.h_in + in_w + in_f, rows, rows_bytes);
        MCU_CUDA(cudaMemcpyAsync(st.in_blob.p, st.h_in, in_bytes, cudaMemcpyHostToDevice, s));
        d_windows = st.in_blob.as<AcWindow>();
        d_weights = (const float*)(st.in_blob.as<u8>() + in_w);
        d_rows = st.in_blob.as<u8>() + in_w + in_f;
        d_counters = st.out_blob.as<unsigned long long>();
        d_counts = (u32*)(st.out_blob.as<u8>() + AC_COUNTER_BYTES);
        d_cols = (u32*)(st.out_blob.as<u8>() + AC_COUNTER_BYTES + out_c);
    } else {
        MCU_TRY(st.windows.reserve(n * sizeof(AcWindow)));
        MCU_TRY(st.rows.reserve(rows_bytes ? rows_bytes : 1));
        MCU_TRY(st.weights.reserve((n_weights ? n_weights : 1) * sizeof(float)));
        MCU_TRY(st.cols.reserve(ct * 4));
        MCU_TRY(st.counts.reserve(n * 4));
        MCU_TRY(st.seg_counters.reserve(AC_COUNTER_BYTES));
        MCU_CUDA(cudaMemcpyAsync(st.windows.p, win.data(), n * sizeof(AcWindow), cudaMemcpyHostToDevice, s));
        if (rows_bytes) MCU_CUDA(cudaMemcpyAsync(st.rows.p, rows, rows_bytes, cudaMemcpyHostToDevice, s));
        if (weights) MCU_CUDA(cudaMemcpyAsync(st.weights.p, weights, n_weights * sizeof(float), cudaMemcpyHostToDevice, s));
        d_windows = st.windows.as<AcWindow>();
        d_weights = st.weights.as<float>();
        d_rows = st.rows.as<u8>();
        d_counters = st.seg_counters.as<unsigned long long>();
        d_counts = st.counts.as<u32>();
        d_cols = st.cols.as<u32>();
    }
    MCU_CUDA(cudaMemsetAsync(d_counters, 0, AC_COUNTER_BYTES, s));
    MCU_CUDA(cudaEventRecord(st.ev0, s));
    // few windows: the latency form (1024 threads, per-column arrays in shared memory); a full batch: CTAs of 256 threads, several per
    // SM, so that one window's serial stretches run under the others' parallel ones
    const bool latency_form = n <= 2ull * (u64)sm_count();
    u32 smem_cols = 0;
    if (latency_form) {
        u32 longest = 0;
        for (u64 i = 0; i < n; ++i) longest = ncol[i] > longest ? ncol[i] : longest;
        const u32 cap = (u32)((st.max_dyn_smem / 8) & ~31u);
        smem_cols = ((longest < cap ? longest : cap) + 31u) & ~31u;
        if (smem_cols > cap) smem_cols = cap;
    }
    anchor_cols_kernel<<<(unsigned)n, latency_form ? AC_BLOCK : 256, (size_t)smem_cols * 8, s>>>(
        d_windows, d_rows, weights ? d_weights : nullptr, p, st.score.as<float>(), st.smooth.as<float>(), st.gg.as<float>(), st.best.as<u32>(),
        st.nxt.as<u32>(), st.heads.as<u32>(), d_cols, d_counts, d_counters, smem_cols, (score_out || smooth_out) ? 1 : 0);
    MCU_CUDA(cudaGetLastError());
    MCU_CUDA(cudaEventRecord(st.ev1, s));
    if (cols_total) {
        if (score_out) MCU_CUDA(cudaMemcpyAsync(score_out, st.score.p, cols_total * 4, cudaMemcpyDeviceToHost, s));
        if (smooth_out) MCU_CUDA(cudaMemcpyAsync(smooth_out, st.smooth.p, cols_total * 4, cudaMemcpyDeviceToHost, s));
    }
    const bool packed = !staged && n > 1 && cols_total * 4 > (1u << 20);   // a batch's columns are packed on the device before they travel
    if (staged) {
        MCU_CUDA(cudaMemcpyAsync(st.h_out, st.out_blob.p, AC_COUNTER_BYTES + out_c + cols_total * 4, cudaMemcpyDeviceToHost, s));
        MCU_CUDA(cudaStreamSynchronize(s));
        memcpy(st.last_counters, st.h_out, sizeof st.last_counters);
        memcpy(n_cols_out, st.h_out + AC_COUNTER_BYTES, n * 4);
        if (cols_total) memcpy(cols_out, st.h_out + AC_COUNTER_BYTES + out_c, cols_total * 4);
    } else {
        MCU_CUDA(cudaMemcpyAsync(n_cols_out, d_counts, n * 4, cudaMemcpyDeviceToHost, s));
        MCU_CUDA(cudaMemcpyAsync(st.last_counters, d_counters, sizeof st.last_counters, cudaMemcpyDeviceToHost, s));
        if (cols_total && !packed) MCU_CUDA(cudaMemcpyAsync(cols_out, d_cols, cols_total * 4, cudaMemcpyDeviceToHost, s));
        MCU_CUDA(cudaStreamSynchronize(s));
    }
    if (device_ms) MCU_CUDA(cudaEventElapsedTime(device_ms, st.ev0, st.ev1));
    if (cols_total && packed) {
        std::vector<u64> off(n + 1);
        off[0] = 0;
        for (u64 i = 0; i < n; ++i) off[i + 1] = off[i] + n_cols_out[i];
        const u64 total = off[n];
        if (total) {
            MCU_TRY(st.dense_off.reserve((n + 1) * 8));
            MCU_TRY(st.dense.reserve(total * 4));
            std::vector<u32> dense(total);
            MCU_CUDA(cudaMemcpyAsync(st.dense_off.p, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
            anchor_cols_pack_kernel<<<(unsigned)n, 128, 0, s>>>(st.windows.as<AcWindow>(), st.cols.as<u32>(), st.counts.as<u32>(), st.dense_off.as<u64>(),
                                                                 st.dense.as<u32>());
            MCU_CUDA(cudaGetLastError());
            MCU_CUDA(cudaMemcpyAsync(dense.data(), st.dense.p, total * 4, cudaMemcpyDeviceToHost, s));
            MCU_CUDA(cudaStreamSynchronize(s));
            for (u64 i = 0; i < n; ++i)
                if (n_cols_out[i]) memcpy(cols_out + col_off[i], dense.data() + off[i], (size_t)n_cols_out[i] * 4);
        }
    }
    return MCU_OK;
}

#else  // MCU_HOST_EMU ----------------------------------------------------------------------------------------------------

// TEST-ONLY host driver of ac_window (tests/_emu.py builds this file with -DMCU_HOST_EMU into tests/_emu/libmcu_emu.so): the CTA is
// one thread, the barriers are nothing.  Returns the number of anchor columns.
static u64 g_emu_seg[2] = {0, 0};
extern "C" void emu_anchor_counters(u64* out2) { out2[0] = g_emu_seg[0]; out2[1] = g_emu_seg[1]; }
extern "C" long long emu_anchor_cols(const u8* rows, u32 n1, u32 n2, u32 ncol, const float* weights, const mcu_anchor_params* p, u32* cols_out,
                                     float* score_out, float* smooth_out)
{
    if (ncol == 0) return 0;
    AcShared* sm = new AcShared;
    AcWindow w;
    w.row_off = 0; w.col_off = 0; w.w_off = 0; w.ncol = ncol; w.n1 = n1; w.n2 = n2; w.pad = 0;
    float* gg = new float[ncol];
    u32 *best = new u32[ncol], *nxt = new u32[ncol], *heads = new u32[ncol];
    u32 count = 0;
    u32* before = new u32[ncol];
    ac_window(*sm, w, rows, weights, *p, score_out, smooth_out, gg, before, best, nxt, heads, cols_out, &count);
    delete[] before;
    g_emu_seg[0] += sm->seg_fast;
    g_emu_seg[1] += sm->seg_chain;
    delete sm; delete[] gg; delete[] best; delete[] nxt; delete[] heads;
    return (long long)count;
}

#endif  // MCU_HOST_EMU

}  // namespace mcu
