// Gapped DP for single-sequence profiles that contain DNA wildcards: a float32 wavefront kernel with the reference's arithmetic.
//
// The integer kernels of dp.cu cover the case in which every profile column is one of A, C, G, T: there all of NWSmall's float
// scores are integers and an int32 wavefront is exact.  A wildcard column (N, X, R, Y, ...) makes the substitution scores
// non-dyadic floats, and then parity needs the reference's own operations: every product, sum and comparison
// in float32 with round-to-nearest and no FMA contraction.  The operations of ONE CELL are fixed by that (nwf_cell below); the
// order in which cells are visited is not (a cell depends on its three upper-left neighbours only), so the matrix is swept as the
// same skewed wavefront dp.cu uses: a warp per region, 32 lanes x 8 rows per stripe, values in registers, cross-lane traffic by
// shuffle, 4 traceback bits per cell.  All values are decided by __host__ __device__ functions that tests/ also run on the CPU
// (-DMCU_HOST_EMU: the grid replaced by two loops over the same cell function).
//
//   counts of a column      MSA::GetFractionalWeightedCounts  MU/msa2.cpp:20-90 (its DNA branch compares nucleotide codes with the
//                           amino-acid constants AX_R = 14 / AX_Y = 19, so only 'X' (NX_X = 14) is split G/A = 1/2, every other
//                           wildcard, N included, gives 1/20 to each of the four letters)
//   profile position        ProfileFromMSA  MU/profilefrommsa.cpp:246-322, SortCounts :159-183
//   substitution score      ScoreProfPos2SPN  MU/scorepp.cpp:80-92
//   recurrence, traceback   NWSmall  MU/nwsmall.cpp:500-670 (macros :68-142), BitTraceBack  MU/bittraceback.cpp:138-
//   terminal gaps           SetTermGaps  MU/termgaps.cpp:6-38 (Half falls through Ext: 0 * -1)
#include "dp.cuh"

#include <algorithm>
#include <vector>

namespace mcu {

#define NWF_HD __host__ __device__ __forceinline__

#ifdef __CUDA_ARCH__
#define NWF_ADD(x, y) __fadd_rn((x), (y))
#define NWF_MUL(x, y) __fmul_rn((x), (y))
#else
static inline float nwf_add_host(float x, float y) { volatile float r = x + y; return r; }
static inline float nwf_mul_host(float x, float y) { volatile float r = x * y; return r; }
#define NWF_ADD(x, y) nwf_add_host((x), (y))
#define NWF_MUL(x, y) nwf_mul_host((x), (y))
#endif

// column class of a sequence byte: 0..3 = A C G T, 4 = 'X' (G/A halves), 5 = any other DNA wildcard, -1 = not a DNA letter
NWF_HD int nwf_class(u8 c)
{
    const unsigned x = ((unsigned)c | 0x20u) - (unsigned)'a';   // letter index, case folded
    if (x >= 26u) return -1;
    // a c g t -> 0 1 2 3; x -> 4; m r w s y k v h d b n -> 5 (MU/alpha.cpp:123-141); everything else -1
    switch (x) {
    case 0: return 0; case 2: return 1; case 6: return 2; case 19: return 3; case 23: return 4;
    case 12: case 17: case 22: case 18: case 24: case 10: case 21: case 7: case 3: case 1: case 13: return 5;
    default: return -1;
    }
}

struct NwfPos {
    float fc[4];
    unsigned order[4];
    float aa[4];
};

NWF_HD float nwf_nuc(int i, int j)   // NUC_SP, MU/nucmx.cpp:8-25 (the +60 centre already added); A C G T = 0 1 2 3
{
    const int d = i ^ j;
    if (d == 0) return (i == 0 || i == 3) ? 151.0f : 160.0f;   // A:A T:T / C:C G:G
    if (d == 1) return -54.0f;                                  // A:C, G:T
    if (d == 2) return 29.0f;                                   // transitions A:G, C:T
    return (i == 0 || i == 3) ? -63.0f : -65.0f;                // A:T / C:G
}

NWF_HD void nwf_build(int cls, NwfPos& p)
{
    for (int i = 0; i < 4; ++i) p.fc[i] = 0.0f;
    const float w = 1.0f;
    if (cls < 4) p.fc[cls] = NWF_ADD(p.fc[cls], w);
    else if (cls == 4) { p.fc[2] = NWF_ADD(p.fc[2], w / 2); p.fc[0] = NWF_ADD(p.fc[0], w / 2); }
    else { const float f = w / 20; for (int i = 0; i < 4; ++i) p.fc[i] = NWF_ADD(p.fc[i], f); }
    for (int i = 0; i < 4; ++i) p.order[i] = (unsigned)i;
    bool any = true;
    while (any) {   // SortCounts: bubble sort, descending, stable
        any = false;
        for (int i = 0; i < 3; ++i) {
            const unsigned i1 = p.order[i], i2 = p.order[i + 1];
            if (p.fc[i1] < p.fc[i2]) { p.order[i + 1] = i1; p.order[i] = i2; any = true; }
        }
    }
    for (int i = 0; i < 4; ++i) {
        float sum = 0.0f;
        for (int j = 0; j < 4; ++j) sum = NWF_ADD(sum, NWF_MUL(p.fc[j], nwf_nuc(i, j)));
        p.aa[i] = sum;
    }
}

// ScoreProfPos2SPN(PA, PB) for the column classes ca, cb
NWF_HD float nwf_pair_score(int ca, int cb)
{
    NwfPos A, B;
    nwf_build(ca, A);
    nwf_build(cb, B);
    float score = 0.0f;
    for (int n = 0; n < 4; ++n) {
        const unsigned l = A.order[n];
        const float f = A.fc[l];
        if (f == 0.0f) break;
        score = NWF_ADD(score, NWF_MUL(f, B.aa[l]));
    }
    return NWF_ADD(score, -0.0f);   // Score - g_scoreCenter with g_scoreCenter = 0
}

// ---- the cell -----------------------------------------------------------------------------------------------------------------
// NWSmall in cell form (MU/nwsmall.cpp:68-142 are its three update macros; :500-670 the row loop around them).  With two
// single-sequence profiles every column has occupancy 1, so the gap scores are the integer case's: open = close = -400 / 2 except the
// terminal zeros (MU/termgaps.cpp:6-38), which, exactly as in dp.cu, only ever meet MINUS_INFINITY or the first row / column; the
// "+ 0" of the gap-extend score is kept as an operation.  What changes with a wildcard column is S (non-dyadic) and with it the
// rounding of every sum below -- so no term may be regrouped: each line is one of the reference's float operations.
//   S      ScoreProfPos2SPN(PA[i-1], PB[j-1])           diag   best of cell (i-1, j-1)   (or the first row / column constants)
//   upM/upD   M, D of cell (i-1, j)                      leftM/leftI   M, I of cell (i, j-1)
struct NwfCell {
    float M, D, I, best;
    u32 nib;   // bits 0-1: state the best predecessor of M[i+1][j+1] is in (0 M, 1 D, 2 I); bit 2: D came from M; bit 3: I came from M
};

NWF_HD NwfCell nwf_cell(float S, float diag, float upM, float upD, float leftM, float leftI)
{
    const float gap = -200.0f;   // 1.0f * g_scoreGapOpen / 2 with g_scoreGapOpen = -400 (MU/params.cpp:296-303, MU/profilefrommsa.cpp:290-291)
    NwfCell c;
    c.M = NWF_ADD(S, diag);                                       // MNext[j+1] = S, then += best (RECURSE_M)
    const float DD = NWF_ADD(upD, 0.0f), MD = NWF_ADD(upM, gap);  // RECURSE_D: stay in D only if strictly better
    const bool keepD = DD > MD;
    c.D = keepD ? DD : MD;
    const float II = NWF_ADD(leftI, 0.0f), MI = NWF_ADD(leftM, gap);   // RECURSE_I: M wins ties
    const bool fromM = MI >= II;
    c.I = fromM ? MI : II;
    const float DM = NWF_ADD(c.D, gap), IM = NWF_ADD(c.I, gap), MM = c.M;   // closing the gap; M, then D, then I on ties
    u32 x;
    if (MM >= DM && MM >= IM) { c.best = MM; x = 0u; }
    else if (DM >= MM && DM >= IM) { c.best = DM; x = 1u; }
    else { c.best = IM; x = 2u; }
    c.nib = x | (keepD ? 0u : 4u) | (fromM ? 8u : 0u);
    return c;
}

NWF_HD float nwf_ninf() { return -1e37f; }   // MINUS_INFINITY, MU/types.h

// BitTraceBack (MU/bittraceback.cpp:138-) one edge back: `nib_here` = nibble of cell (pa, pb), `nib_diag` = nibble of cell (pa-1, pb-1)
// (the predecessor bits of M[pa][pb] are produced by the cell diagonally above it; first row / column: nwsmall.cpp:586-592)
NWF_HD char nwf_prev_edge(char edge, u32 pa, u32 pb, u32 nib_here, u32 nib_diag)
{
    if (edge == 'M') {
        if (pa >= 2 && pb >= 2) { const u32 x = nib_diag & 3u; return x == 0 ? 'M' : (x == 1 ? 'D' : 'I'); }
        return pa >= 2 ? 'D' : 'I';
    }
    if (edge == 'D') return (nib_here & 4u) ? 'M' : 'D';
    return (nib_here & 8u) ? 'M' : 'I';
}

#ifndef MCU_HOST_EMU
// ---- wavefront kernel: the structure of dp.cu's integer kernel with float registers ---------------------------------------------
// One warp per region; rows in stripes of 32 lanes x 8 rows; lane l is l columns behind lane l-1, so a step needs the previous
// step's bottom-row (M, D, best) of the lane above: three __shfl_up_sync plus one for the column class.  4 traceback bits per cell,
// one 32-bit word per lane and step, stored step-major (coalesced), in the layout nw_traceback of dp.cu reads.
constexpr int WF_R = 8;
constexpr int WF_STRIPE = 32 * WF_R;
constexpr unsigned WF_FULL = 0xffffffffu;

struct NwfArgs {
    const u8* a;
    const u8* b;
    const u64* a_off;
    const u64* b_off;
    const u64* tb_off;     // words, per region
    u32* tb;
    float4* boundary;      // per warp: 2 x bstride entries {M, D, best of the previous column, -} of a stripe's last row
    u64 bstride;
    float4* result;        // per region {M, D, I, -} at (la, lb)
    unsigned* counter;
    u32 n;
};

__global__ void __launch_bounds__(256) nwf_forward_kernel(const __grid_constant__ NwfArgs g)
{
    __shared__ float sub[36];
    __shared__ float4 ring[8][2][32];
    if (threadIdx.x < 36) sub[threadIdx.x] = nwf_pair_score((int)threadIdx.x / 6, (int)threadIdx.x % 6);
    __syncthreads();
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gwarp = (u64)blockIdx.x * 8 + warp;
    const float NINF = nwf_ninf();
    for (;;) {
        u32 p = 0;
        if (lane == 0) p = atomicAdd(g.counter, 1u);
        p = __shfl_sync(WF_FULL, p, 0);
        if (p >= g.n) break;
        const u8* A = g.a + g.a_off[p];
        const u8* B = g.b + g.b_off[p];
        const u32 la = (u32)(g.a_off[p + 1] - g.a_off[p]), lb = (u32)(g.b_off[p + 1] - g.b_off[p]);
        const u32 T = lb + 31;
        const u32 nstripes = (la + WF_STRIPE - 1) / WF_STRIPE;
        const u32 fin_lane = ((la - 1) % WF_STRIPE) / WF_R;
        const int fin_r = (int)((la - 1) % WF_R);
        float capM = 0.f, capD = 0.f, capI = 0.f;
        for (u32 s = 0; s < nstripes; ++s) {
            const u32 i0 = s * WF_STRIPE + lane * WF_R;
            int ca[WF_R];
            float Ml[WF_R], Il[WF_R], Bp[WF_R];
#pragma unroll
            for (int r = 0; r < WF_R; ++r) {
                ca[r] = i0 + r < la ? 6 * nwf_class(A[i0 + r]) : 0;
                Ml[r] = NINF;      // M[i][0]
                Il[r] = NINF;      // I[i][0]
                Bp[r] = -200.0f;   // "best" of column 0: gives M[i][1] = S + open_A[0] + 0 + close_A[i-2] = S - 200 (nwsmall.cpp:586-592)
            }
            float diag = -200.0f, out_M = NINF, out_D = NINF, out_B = -200.0f;
            int cb_cur = 0;
            const float4* bprev = g.boundary + (gwarp * 2 + ((s + 1) & 1)) * g.bstride;
            float4* bcur = g.boundary + (gwarp * 2 + (s & 1)) * g.bstride;
            const bool last = s + 1 == nstripes;
            const u32 cap_t = lb - 1 + fin_lane;
            u32* tbs = g.tb + g.tb_off[p] + (u64)s * T * 32 + lane;
            for (u32 t = 0; t < T; ++t) {
                if ((t & 31u) == 0) {   // inputs of lane 0 for columns t+1 .. t+32: the row above this stripe
                    const u32 jk = t + 1 + lane;
                    float4 v = make_float4(NINF, NINF, -200.0f, 0.f);   // row 0: M = D = -inf; best[0][j] = -200 gives M[1][j] = S - 200
                    if (jk <= lb) {
                        v.w = __int_as_float(nwf_class(B[jk - 1]));
                        if (s == 0) {
                            if (jk == 1 && la > 1) v.z = 0.0f;   // M[1][1] = S; with la == 1 the last-row form S + 0 - 200 applies (:645-650)
                        } else {
                            const float4 q = __ldcg(bprev + jk);
                            v.x = q.x;
                            v.y = q.y;
                            if (jk > 1) v.z = q.z;
                        }
                    }
                    ring[warp][(t >> 5) & 1][lane] = v;
                    __syncwarp();
                }
                float upM = __shfl_up_sync(WF_FULL, out_M, 1);
                float upD = __shfl_up_sync(WF_FULL, out_D, 1);
                const float upB = __shfl_up_sync(WF_FULL, out_B, 1);
                int cb = __shfl_up_sync(WF_FULL, cb_cur, 1);
                float dg = diag;
                diag = upB;
                if (lane == 0) {
                    const float4 v = ring[warp][(t >> 5) & 1][t & 31u];
                    upM = v.x;
                    upD = v.y;
                    dg = v.z;
                    cb = __float_as_int(v.w);
                }
                cb_cur = cb;
                const int j = (int)t - (int)lane + 1;
                if (j >= 1 && j <= (int)lb) {
                    u32 tbw = 0;
                    float best = 0.f;
#pragma unroll
                    for (int r = 0; r < WF_R; ++r) {
                        const NwfCell c = nwf_cell(sub[ca[r] + cb], dg, upM, upD, Ml[r], Il[r]);
                        tbw |= c.nib << (4 * r);
                        if (last && t == cap_t && r == fin_r) { capM = c.M; capD = c.D; capI = c.I; }
                        dg = Bp[r];
                        Bp[r] = c.best;
                        best = c.best;
                        upM = c.M;
                        upD = c.D;
                        Ml[r] = c.M;
                        Il[r] = c.I;
                    }
                    out_M = upM;
                    out_D = upD;
                    out_B = best;
                    __stcs(tbs + (u64)t * 32, tbw);
                    if (lane == 31 && !last) {
                        bcur[j].x = out_M;
                        bcur[j].y = out_D;
                        bcur[j + 1].z = out_B;
                    }
                }
            }
            __threadfence_block();
            __syncwarp();
        }
        if (lane == fin_lane) g.result[p] = make_float4(capM, capD, capI, 0.f);
        __syncwarp();
    }
}

struct NwfTbArgs {
    const u64* a_off;
    const u64* b_off;
    const u32* tb;
    const u64* tb_off;
    const float4* result;
    const u64* path_off;
    char* path;
    u32* path_len;
    u64* path_start;
    float* score;
    u32 n;
};

__device__ __forceinline__ u32 nwf_nibble(const u32* __restrict__ tb, u32 T, u32 i, u32 j)
{
    const u32 row = i - 1;
    const u32 s = row / WF_STRIPE, lane = (row % WF_STRIPE) / WF_R, r = row % WF_R;
    const u32 t = (j - 1) + lane;
    return (__ldg(tb + ((u64)s * T + t) * 32 + lane) >> (4 * r)) & 15u;
}

__global__ void __launch_bounds__(128) nwf_traceback_kernel(NwfTbArgs g)
{
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const u32 la = (u32)(g.a_off[p + 1] - g.a_off[p]), lb = (u32)(g.b_off[p + 1] - g.b_off[p]);
    const u32* tb = g.tb + g.tb_off[p];
    const u32 T = lb + 31;
    const float4 res = g.result[p];
    float sc = res.x;
    char edge = 'M';
    if (res.y > sc) { sc = res.y; edge = 'D'; }   // nwsmall.cpp:645-656
    if (res.z > sc) { sc = res.z; edge = 'I'; }
    g.score[p] = sc;
    char* out = g.path + g.path_off[p] + la + lb;   // written backwards, right-aligned in the slot
    u32 pa = la, pb = lb, n = 0;
    for (;;) {
        *--out = edge;
        ++n;
        const u32 here = (edge == 'M' || pa == 0 || pb == 0) ? 0u : nwf_nibble(tb, T, pa, pb);   // first row / column: the gap runs to the corner
        const u32 dg = (edge == 'M' && pa >= 2 && pb >= 2) ? nwf_nibble(tb, T, pa - 1, pb - 1) : 0u;
        const char next = nwf_prev_edge(edge, pa, pb, here, dg);
        if (edge != 'I') --pa;
        if (edge != 'D') --pb;
        if (pa == 0 && pb == 0) break;
        edge = next;
        if ((edge == 'M' && (pa == 0 || pb == 0)) || (edge == 'D' && pa == 0) || (edge == 'I' && pb == 0)) { n = 0; break; }   // the reference Quit()s
    }
    g.path_len[p] = n;
    g.path_start[p] = (u64)(out - g.path);
}

// moves every right-aligned path to the front of its slot: one warp per region
__global__ void __launch_bounds__(256) nwf_shift_kernel(u32 count, const u64* path_off, const u64* path_start, const u32* path_len, char* path)
{
    const u32 lane = threadIdx.x & 31;
    const u64 wid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nw = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 p = wid; p < count; p += nw) {
        const u64 dst = path_off[p], src = path_start[p];
        const u32 n = path_len[p];
        if (src == dst) continue;
        for (u32 c = 0; c < n; c += 32) {
            char v = 0;
            if (c + lane < n) v = path[src + c + lane];
            __syncwarp();
            if (c + lane < n) path[dst + c + lane] = v;
            __syncwarp();
        }
    }
}

static DevBuf w_a, w_b, w_aoff, w_boff, w_tboff, w_pathoff, w_tb, w_bnd, w_res, w_path, w_plen, w_pstart, w_score, w_ctr;

void nwf_release()
{
    DevBuf* bufs[] = {&w_a, &w_b, &w_aoff, &w_boff, &w_tboff, &w_pathoff, &w_tb, &w_bnd, &w_res, &w_path, &w_plen, &w_pstart, &w_score, &w_ctr};
    for (DevBuf* b : bufs) b->release();
}
static cudaStream_t w_stream = nullptr;

// regions [i0, i1) of the caller's batch (offsets rebased by the caller's arrays themselves: absolute offsets are used)
static int nwf_run_slice(u64 i0, u64 i1, const char* a, const u64* a_off, const char* b, const u64* b_off, const u64* path_off, char* path_out,
                         u32* path_len, float* score, const std::vector<u64>& tbw, float* ms_accum)
{
    const u64 n = i1 - i0;
    cudaStream_t s = w_stream;
    const u64 a0 = a_off[i0], b0 = b_off[i0], p0 = path_off[i0];
    const u64 abytes = a_off[i1] - a0, bbytes = b_off[i1] - b0, pbytes = path_off[i1] - p0;
    std::vector<u64> ao(n + 1), bo(n + 1), po(n + 1), to(n + 1, 0);
    u64 max_lb = 0;
    for (u64 k = 0; k <= n; ++k) { ao[k] = a_off[i0 + k] - a0; bo[k] = b_off[i0 + k] - b0; po[k] = path_off[i0 + k] - p0; }
    for (u64 k = 0; k < n; ++k) { to[k + 1] = to[k] + tbw[i0 + k]; max_lb = std::max(max_lb, bo[k + 1] - bo[k]); }
    const int ctas = sm_count() * 2;
    const u64 nwarps = (u64)ctas * 8, bstride = max_lb + 4;
    MCU_TRY(w_a.reserve(abytes + 16));
    MCU_TRY(w_b.reserve(bbytes + 16));
    MCU_TRY(w_aoff.reserve((n + 1) * 8));
    MCU_TRY(w_boff.reserve((n + 1) * 8));
    MCU_TRY(w_tboff.reserve((n + 1) * 8));
    MCU_TRY(w_pathoff.reserve((n + 1) * 8));
    MCU_TRY(w_tb.reserve(to[n] * 4 + 16));
    MCU_TRY(w_bnd.reserve(nwarps * 2 * bstride * sizeof(float4)));
    MCU_TRY(w_res.reserve(n * sizeof(float4)));
    MCU_TRY(w_path.reserve(pbytes + 16));
    MCU_TRY(w_plen.reserve(n * 4));
    MCU_TRY(w_pstart.reserve(n * 8));
    MCU_TRY(w_score.reserve(n * 4));
    MCU_TRY(w_ctr.reserve(64));
    MCU_CUDA(cudaMemcpyAsync(w_a.p, a + a0, abytes, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_b.p, b + b0, bbytes, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_aoff.p, ao.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_boff.p, bo.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_tboff.p, to.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_pathoff.p, po.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemsetAsync(w_ctr.p, 0, 64, s));
    MCU_CUDA(cudaMemsetAsync(w_path.p, 0, pbytes, s));
    cudaEvent_t e0, e1;
    MCU_CUDA(cudaEventCreate(&e0));
    MCU_CUDA(cudaEventCreate(&e1));
    NwfArgs g;
    g.a = w_a.as<u8>(); g.b = w_b.as<u8>(); g.a_off = w_aoff.as<u64>(); g.b_off = w_boff.as<u64>(); g.tb_off = w_tboff.as<u64>();
    g.tb = w_tb.as<u32>(); g.boundary = w_bnd.as<float4>(); g.bstride = bstride; g.result = w_res.as<float4>();
    g.counter = w_ctr.as<unsigned>(); g.n = (u32)n;
    MCU_CUDA(cudaEventRecord(e0, s));
    nwf_forward_kernel<<<(unsigned)std::min<u64>(div_up(n, 8), (u64)ctas), 256, 0, s>>>(g);
    NwfTbArgs t;
    t.a_off = g.a_off; t.b_off = g.b_off; t.tb = g.tb; t.tb_off = g.tb_off; t.result = g.result; t.path_off = w_pathoff.as<u64>();
    t.path = w_path.as<char>(); t.path_len = w_plen.as<u32>(); t.path_start = w_pstart.as<u64>(); t.score = w_score.as<float>(); t.n = (u32)n;
    nwf_traceback_kernel<<<(unsigned)div_up(n, 128), 128, 0, s>>>(t);
    nwf_shift_kernel<<<(unsigned)std::min<u64>(div_up(n * 32, 256), (u64)sm_count() * 8), 256, 0, s>>>((u32)n, t.path_off, t.path_start, t.path_len, t.path);
    MCU_CUDA(cudaEventRecord(e1, s));
    MCU_CUDA(cudaGetLastError());
    MCU_CUDA(cudaMemcpyAsync(path_out + p0, w_path.p, pbytes, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(path_len + i0, w_plen.p, n * 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(score + i0, w_score.p, n * 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_accum += ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MCU_OK;
}

int nw_batch_wild(u64 n, const char* a, const u64* a_off, const char* b, const u64* b_off, const u64* path_off, char* path_out, u32* path_len,
                  float* score, float* device_ms)
{
    if (device_ms) *device_ms = 0.f;
    if (n == 0) return MCU_OK;
    if (!a || !b || !a_off || !b_off || !path_off || !path_out || !path_len || !score) { set_error("mcu_nw_batch_wild: NULL pointer"); return MCU_EINVAL; }
    if (n >= 0xFFFFFFFFull) { set_error("mcu_nw_batch_wild: too many regions"); return MCU_EINVAL; }
    std::vector<u64> tbw(n);
    for (u64 i = 0; i < n; ++i) {
        if (a_off[i + 1] <= a_off[i] || b_off[i + 1] <= b_off[i]) { set_error("mcu_nw_batch_wild: region %llu is empty", (unsigned long long)i); return MCU_EINVAL; }
        const u64 la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        if (la > 0x3FFFFFFull || lb > 0x3FFFFFFull) { set_error("mcu_nw_batch_wild: region %llu too long", (unsigned long long)i); return MCU_EINVAL; }
        if (path_off[i + 1] - path_off[i] < la + lb) { set_error("mcu_nw_batch_wild: path slot %llu smaller than la+lb", (unsigned long long)i); return MCU_EINVAL; }
        for (u64 k = a_off[i]; k < a_off[i + 1]; ++k)
            if (nwf_class((u8)a[k]) < 0) { set_error("mcu_nw_batch_wild: byte %d is not a DNA letter or wildcard", (int)(u8)a[k]); return MCU_EALPHA; }
        for (u64 k = b_off[i]; k < b_off[i + 1]; ++k)
            if (nwf_class((u8)b[k]) < 0) { set_error("mcu_nw_batch_wild: byte %d is not a DNA letter or wildcard", (int)(u8)b[k]); return MCU_EALPHA; }
        tbw[i] = div_up(la, WF_STRIPE) * (lb + 31) * 32;
    }
    if (!w_stream) MCU_CUDA(cudaStreamCreateWithFlags(&w_stream, cudaStreamNonBlocking));
    size_t free_b = 0, total_b = 0;
    MCU_CUDA(cudaMemGetInfo(&free_b, &total_b));
    u64 budget_words = (u64)(free_b * 0.6) / 4;
    if (budget_words > (32ull << 30) / 4) budget_words = (32ull << 30) / 4;
    float ms = 0.f;
    u64 i = 0;
    while (i < n) {   // slices whose traceback words fit the device
        u64 words = 0, j = i;
        while (j < n && (j == i || words + tbw[j] <= budget_words)) words += tbw[j++];
        if (words > budget_words) { set_error("mcu_nw_batch_wild: region %llu needs %llu MiB of traceback, more than the device has free", (unsigned long long)i, (unsigned long long)(words >> 18)); return MCU_ENOMEM; }
        MCU_TRY(nwf_run_slice(i, j, a, a_off, b, b_off, path_off, path_out, path_len, score, tbw, &ms));
        i = j;
    }
    if (device_ms) *device_ms = ms;
    for (u64 k = 0; k < n; ++k)
        if (path_len[k] == 0) { set_error("mcu_nw_batch_wild: inconsistent traceback for region %llu", (unsigned long long)k); return MCU_ECUDA; }
    return MCU_OK;
}

#else  // MCU_HOST_EMU ----------------------------------------------------------------------------------------------------
}  // namespace mcu
extern "C" {
// TEST-ONLY host driver (tests/_emu.py): one region through the SAME cell and traceback functions the kernels execute, with the
// grid replaced by two loops (cell (i, j) after (i-1, j), (i, j-1), (i-1, j-1): any such order gives the wavefront's values).
// Returns the path length.
__attribute__((visibility("default"))) long long emu_nw_wild(const char* a, unsigned la, const char* b, unsigned lb, char* path_out, float* score_out)
{
    using namespace mcu;
    for (unsigned i = 0; i < la; ++i) if (nwf_class((u8)a[i]) < 0) return -1;
    for (unsigned j = 0; j < lb; ++j) if (nwf_class((u8)b[j]) < 0) return -1;
    if (la == 0 || lb == 0) return -1;
    float sub[36];
    for (int k = 0; k < 36; ++k) sub[k] = nwf_pair_score(k / 6, k % 6);
    const float NINF = nwf_ninf();
    const size_t W = (size_t)lb + 1;
    std::vector<float> M(2 * W), D(2 * W), Bst(2 * W), Irow(W);
    std::vector<u8> nib(((size_t)la + 1) * W, 0);
    // row 0
    for (size_t j = 0; j <= lb; ++j) { M[j] = NINF; D[j] = NINF; Bst[j] = -200.0f; }
    Bst[0] = la > 1 ? 0.0f : -200.0f;
    float fM = 0, fD = 0, fI = 0;
    for (unsigned i = 1; i <= la; ++i) {
        float* Mp = &M[((i - 1) & 1) * W]; float* Mc = &M[(i & 1) * W];
        float* Dp = &D[((i - 1) & 1) * W]; float* Dc = &D[(i & 1) * W];
        float* Bp = &Bst[((i - 1) & 1) * W]; float* Bc = &Bst[(i & 1) * W];
        Mc[0] = NINF; Bc[0] = -200.0f;
        float leftI = NINF;
        for (unsigned j = 1; j <= lb; ++j) {
            const NwfCell c = nwf_cell(sub[6 * nwf_class((u8)a[i - 1]) + nwf_class((u8)b[j - 1])], Bp[j - 1], Mp[j], Dp[j], Mc[j - 1], leftI);
            Mc[j] = c.M; Dc[j] = c.D; Bc[j] = c.best; leftI = c.I;
            nib[(size_t)i * W + j] = (u8)c.nib;
            if (i == la && j == lb) { fM = c.M; fD = c.D; fI = c.I; }
        }
    }
    float sc = fM;
    char edge = 'M';
    if (fD > sc) { sc = fD; edge = 'D'; }
    if (fI > sc) { sc = fI; edge = 'I'; }
    if (score_out) *score_out = sc;
    std::vector<char> rev;
    unsigned pa = la, pb = lb;
    for (;;) {
        rev.push_back(edge);
        const u32 here = nib[(size_t)pa * W + pb];
        const u32 dg = (pa >= 1 && pb >= 1) ? nib[(size_t)(pa - 1) * W + (pb - 1)] : 0u;
        const char next = nwf_prev_edge(edge, pa, pb, here, dg);
        if (edge != 'I') --pa;
        if (edge != 'D') --pb;
        if (pa == 0 && pb == 0) break;
        edge = next;
        if ((edge == 'M' && (pa == 0 || pb == 0)) || (edge == 'D' && pa == 0) || (edge == 'I' && pb == 0)) return -1;
        if (rev.size() > (size_t)la + lb) return -1;
    }
    for (size_t k = 0; k < rev.size(); ++k) path_out[k] = rev[rev.size() - 1 - k];
    return (long long)rev.size();
}
}
namespace mcu {
#endif

}  // namespace mcu
