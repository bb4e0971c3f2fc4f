// Anchor columns of alignment windows on the device: replaces muscle::FindAnchorColsPP (MU/anchoredpp.cpp:354-409)
//   LetterObjScoreXP    MU/anchoredpp.cpp:256-329  (ScoreSeqPairLetters :19-93, per-site ScoreSeqPairGaps :96-250, rows' weights)
//   WindowSmooth        MU/anchors.cpp:9-47
//   FindBestColsComboPP MU/anchoredpp.cpp:335-351
//   MergeBestCols       MU/anchors.cpp:137-186
// One CTA per window; the batch is the grid.  What the reference does column after column is split by what really depends on what:
//   * letter scores: per column;
//   * gap penalties: the reference's running state (bGapping1/2, gap_left_col, cur_gap_score) is reset by every column that holds two
//     letters, so the columns between two such columns form a run that is scored on its own: one thread walks each run with the
//     reference's state machine and spreads the penalty, all runs at once;
//   * smoothing: ONE running float sum per window whose roundings depend on the order -- kept as the serial chain it is (two FADDs per
//     column on one lane), fed from shared memory tile by tile, everything around it in parallel;
//   * best columns: flag + ordered block-wide compaction;
//   * merging: `next group head` is a function of the head alone (a binary search per best column, in parallel), the groups are then the
//     walk head -> next head (one lane, a few hundred steps), the pick per group in parallel.
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
#else
#define AC_HD __device__ __forceinline__
#define AC_TID threadIdx.x
#define AC_NT blockDim.x
#define AC_SYNC() __syncthreads()
#define AC_FADD(a, b) __fadd_rn(a, b)
#define AC_FSUB(a, b) __fsub_rn(a, b)
#define AC_FMUL(a, b) __fmul_rn(a, b)
#define AC_FDIV(a, b) __fdiv_rn(a, b)
#endif

#define AC_BLOCK 512
#define AC_TILE 2048

struct AcWindow {
    u64 row_off;   // first character of the window's first row
    u64 col_off;   // where the window's per-column outputs / scratch start
    u64 w_off;     // first weight
    u32 ncol, n1, n2, pad;
};

struct AcShared {
    u8 letter[256];
    float sub[AC_TILE], add[AC_TILE], out[AC_TILE];
    float total;
    int first, last;
    u32 warp_sum[AC_BLOCK / 32];
    u32 base, nbest, nanchor;
};

// the gap penalties of one run of columns that starts at c0 (the first column of the pair's range, or the column after one with two
// letters) and ends before the next column with two letters: ScoreSeqPairGaps :155-250 with its state as it is at a run's start
AC_HD void ac_gap_run(const u8* r1, const u8* r2, const u8* letter, u32 c0, u32 first, u32 last, u32 ncol, const mcu_anchor_params& p, float* gg)
{
    bool in1 = false, in2 = false;
    u32 left = 0;
    float cur = 0.0f;
    u32 c = c0;
    for (; c <= last; ++c) {
        const bool g1 = letter[r1[c]] == MCU_AC_GAP, g2 = letter[r2[c]] == MCU_AC_GAP;
        if (g1 && g2) continue;
        if (!g1 && !g2) break;
        bool& in = g1 ? in1 : in2;
        if (!in) {
            left = c;
            cur = AC_FADD(cur, c == first ? p.term_gap : p.gap_open);
            in = true;
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
// exclusive rank of this thread's flag inside the CTA and the CTA's total (every thread calls it)
__device__ __forceinline__ u32 ac_block_rank(AcShared& sm, bool flag, u32& total)
{
    const u32 b = __ballot_sync(0xffffffffu, flag);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm.warp_sum[warp] = __popc(b);
    __syncthreads();
    u32 before = 0, all = 0;
    for (u32 w = 0; w < AC_BLOCK / 32; ++w) {
        const u32 v = sm.warp_sum[w];
        before += w < warp ? v : 0;
        all += v;
    }
    __syncthreads();
    total = all;
    return before + __popc(b & ((1u << lane) - 1));
}
#endif

// One window, start to end.  score / smooth / gg / best / nxt / heads: ncol entries each at the window's col_off.
AC_HD void ac_window(AcShared& sm, const AcWindow w, const u8* rows_all, const float* weights, const mcu_anchor_params& p, float* score, float* smooth,
                     float* gg, u32* best, u32* nxt, u32* heads, u32* cols_out, u32* count_out)
{
    const u32 tid = AC_TID, nt = AC_NT;
    const u32 L = w.ncol;
    const u8* rows = rows_all + w.row_off;
    for (u32 i = tid; i < 256; i += nt) sm.letter[i] = p.letter_of_char[i];
    for (u32 c = tid; c < L; c += nt) {
        score[c] = 0.0f;
        smooth[c] = 0.0f;
        gg[c] = 0.0f;
    }
    if (tid == 0) {
        sm.base = 0;
        sm.nanchor = 0;
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
            // first / last column where not both rows have a gap (:46-74)
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
            // the runs of columns with a gap in one row, each by the thread that owns its first column
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

    // ---- WindowSmooth: the running sum is one serial chain (thread 0); its operands come through shared memory
    const u32 W = p.smooth_window, w2 = W / 2;
    if (L > W) {
        if (tid == 0) {
            float t = 0.0f;
            for (u32 i = 0; i < W; ++i) t = AC_FADD(t, ac_ceil(score[i], p.smooth_ceil));
            sm.total = t;
        }
        const u32 i_last = L - w2 - 1;
        for (u32 i0 = w2; i0 <= i_last; i0 += AC_TILE) {
            const u32 n = (i_last - i0 + 1) < AC_TILE ? (i_last - i0 + 1) : AC_TILE;
            for (u32 k = tid; k < n; k += nt) {
                const u32 i = i0 + k;
                sm.sub[k] = ac_ceil(score[i - w2], p.smooth_ceil);
                sm.add[k] = i + w2 + 1 < L ? ac_ceil(score[i + w2 + 1], p.smooth_ceil) : 0.0f;   // not used at i_last
            }
            AC_SYNC();
            if (tid == 0) {
                float t = sm.total;
                const float fw = (float)W;
                for (u32 k = 0; k < n; ++k) {
                    sm.out[k] = AC_FDIV(t, fw);
                    if (i0 + k == i_last) break;
                    t = AC_FSUB(t, sm.sub[k]);
                    t = AC_FADD(t, sm.add[k]);
                }
                sm.total = t;
            }
            AC_SYNC();
            for (u32 k = tid; k < n; k += nt) smooth[i0 + k] = sm.out[k];
            AC_SYNC();
        }
    }
    AC_SYNC();

    // ---- FindBestColsComboPP: the columns that pass both thresholds, in order
    for (u32 c0 = 0; c0 < L; c0 += nt) {
        const u32 c = c0 + tid;
        const bool flag = c < L && !(score[c] < p.min_best_col) && !(smooth[c] < p.min_smooth);
#ifdef MCU_HOST_EMU
        if (flag) best[sm.base++] = c;
#else
        u32 total;
        const u32 rank = ac_block_rank(sm, flag, total);
        if (flag) best[sm.base + rank] = c;
        __syncthreads();
        if (tid == 0) sm.base += total;
        __syncthreads();
#endif
    }
    AC_SYNC();
    const u32 nbest = sm.base;

    // ---- MergeBestCols: groups of best columns closer to the group's first one than the spacing
    for (u32 n = tid; n < nbest; n += nt) {   // where the group that starts at n ends: the first i > n with best[i] - best[n] >= spacing
        const u32 head = best[n];
        u32 lo = n + 1, hi = nbest;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (best[mid] - head >= p.anchor_spacing) hi = mid;
            else lo = mid + 1;
        }
        nxt[n] = lo;
    }
    AC_SYNC();
    if (tid == 0) {
        u32 k = 0;
        for (u32 n = 0; n < nbest; n = nxt[n]) heads[k++] = n;
        sm.nanchor = k;
    }
    AC_SYNC();
    const u32 nanchor = sm.nanchor;
    for (u32 k = tid; k < nanchor; k += nt) {
        const u32 n = heads[k], within = nxt[n] - n - 1, head = best[n];
        u32 pick = head;
        if (within == 1) {
            const u32 other = best[n + 1];
            pick = score[head] > score[other] ? head : other;
        } else if (within > 1) {
            // "closest to the centre" as the reference computes it: the distance is taken from the group's first column, and the last
            // member of the group is not looked at (MU/anchors.cpp:164-179)
            int closest = (int)p.anchor_spacing;
            for (u32 i = n + 1; i < n + within; ++i) {
                int d = (int)(best[i] - head);
                if (d < 0) d = -d;
                if (d < closest) {
                    pick = best[i];
                    closest = d;
                }
            }
        }
        cols_out[k] = pick;
    }
    if (tid == 0) *count_out = nanchor;
}

#ifndef MCU_HOST_EMU

__global__ __launch_bounds__(AC_BLOCK) void anchor_cols_kernel(const AcWindow* __restrict__ windows, const u8* __restrict__ rows,
                                                               const float* __restrict__ weights, const mcu_anchor_params p, float* score, float* smooth,
                                                               float* gg, u32* best, u32* nxt, u32* heads, u32* cols_out, u32* counts)
{
    __shared__ AcShared sm;
    const AcWindow w = windows[blockIdx.x];
    if (w.ncol == 0) {
        if (threadIdx.x == 0) counts[blockIdx.x] = 0;
        return;
    }
    ac_window(sm, w, rows, weights, p, score + w.col_off, smooth + w.col_off, gg + w.col_off, best + w.col_off, nxt + w.col_off, heads + w.col_off,
              cols_out + w.col_off, counts + blockIdx.x);
}

struct AcState {
    DevBuf windows, rows, weights, score, smooth, gg, best, nxt, heads, cols, counts;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
static AcState g_ac;

void ac_release()
{
    AcState& st = g_ac;
    DevBuf* bufs[] = {&st.windows, &st.rows, &st.weights, &st.score, &st.smooth, &st.gg, &st.best, &st.nxt, &st.heads, &st.cols, &st.counts};
    for (DevBuf* b : bufs) b->release();
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
        MCU_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        MCU_CUDA(cudaEventCreate(&st.ev0));
        MCU_CUDA(cudaEventCreate(&st.ev1));
    }
    cudaStream_t s = st.stream;
    const u64 ct = cols_total ? cols_total : 1;
    MCU_TRY(st.windows.reserve(n * sizeof(AcWindow)));
    MCU_TRY(st.rows.reserve(rows_bytes ? rows_bytes : 1));
    MCU_TRY(st.weights.reserve((n_weights ? n_weights : 1) * sizeof(float)));
    MCU_TRY(st.score.reserve(ct * 4)); MCU_TRY(st.smooth.reserve(ct * 4)); MCU_TRY(st.gg.reserve(ct * 4));
    MCU_TRY(st.best.reserve(ct * 4)); MCU_TRY(st.nxt.reserve(ct * 4)); MCU_TRY(st.heads.reserve(ct * 4)); MCU_TRY(st.cols.reserve(ct * 4));
    MCU_TRY(st.counts.reserve(n * 4));
    MCU_CUDA(cudaMemcpyAsync(st.windows.p, win.data(), n * sizeof(AcWindow), cudaMemcpyHostToDevice, s));
    if (rows_bytes) MCU_CUDA(cudaMemcpyAsync(st.rows.p, rows, rows_bytes, cudaMemcpyHostToDevice, s));
    if (weights) MCU_CUDA(cudaMemcpyAsync(st.weights.p, weights, n_weights * sizeof(float), cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaEventRecord(st.ev0, s));
    anchor_cols_kernel<<<(unsigned)n, AC_BLOCK, 0, s>>>(st.windows.as<AcWindow>(), st.rows.as<u8>(), weights ? st.weights.as<float>() : nullptr, p,
                                                         st.score.as<float>(), st.smooth.as<float>(), st.gg.as<float>(), st.best.as<u32>(),
                                                         st.nxt.as<u32>(), st.heads.as<u32>(), st.cols.as<u32>(), st.counts.as<u32>());
    MCU_CUDA(cudaGetLastError());
    MCU_CUDA(cudaEventRecord(st.ev1, s));
    MCU_CUDA(cudaMemcpyAsync(n_cols_out, st.counts.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (cols_total) {
        MCU_CUDA(cudaMemcpyAsync(cols_out, st.cols.p, cols_total * 4, cudaMemcpyDeviceToHost, s));
        if (score_out) MCU_CUDA(cudaMemcpyAsync(score_out, st.score.p, cols_total * 4, cudaMemcpyDeviceToHost, s));
        if (smooth_out) MCU_CUDA(cudaMemcpyAsync(smooth_out, st.smooth.p, cols_total * 4, cudaMemcpyDeviceToHost, s));
    }
    MCU_CUDA(cudaStreamSynchronize(s));
    if (device_ms) MCU_CUDA(cudaEventElapsedTime(device_ms, st.ev0, st.ev1));
    return MCU_OK;
}

#else  // MCU_HOST_EMU ----------------------------------------------------------------------------------------------------

// TEST-ONLY host driver of ac_window (tests/_emu.py builds this file with -DMCU_HOST_EMU into tests/_emu/libmcu_emu.so): the CTA is
// one thread, the barriers are nothing.  Returns the number of anchor columns.
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
    ac_window(*sm, w, rows, weights, *p, score_out, smooth_out, gg, best, nxt, heads, cols_out, &count);
    delete sm; delete[] gg; delete[] best; delete[] nxt; delete[] heads;
    return (long long)count;
}

#endif  // MCU_HOST_EMU

}  // namespace mcu
