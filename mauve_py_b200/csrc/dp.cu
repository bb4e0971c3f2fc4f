// Gapped DP for batches of inter-anchor regions (sm_100a).
//
// Replaces muscle::GlobalAlign -> NWSmall (MU/nwsmall.cpp:500-670, macros :68-142) + BitTraceBack
// (MU/bittraceback.cpp:138-) for two single-sequence ACGT profiles, i.e. exactly what
// AlignTwoProfs (MU/aligntwoprofs.cpp:23) runs per inter-anchor range in the two-genome case.
//
// Scores are the reference's floats, which are exact integers here:
//   substitution NUC_SP (MU/nucmx.cpp:8-25, HOXD70 + 60 centre), gap open = close = -400/2
//   (MU/params.cpp:296-303, MU/profilefrommsa.cpp:290-291), gap extend 0, terminal open/close 0
//   (MU/termgaps.cpp:19-33).  Working the reference loops through shows that the terminal
//   zeros only ever reach the result through the first row/column initialisation
//   (M[i][1] = S - 200, M[1][j] = S - 200, M[1][1] = S; la == 1 gives M[1][1] = S - 200), that the
//   closing terms closeA[la-1] / closeB[lb-1] are never read, and that openA[0]/openB[0] are only
//   added to MINUS_INFINITY.  The recurrence below is therefore (D' = D - 200, I' = I - 200):
//       M[i][j]  = S(a_i, b_j) + best[i-1][j-1]           best[0][0] = 0 (or -200 if la == 1), best[i][0] = best[0][j] = -200
//       D'[i][j] = max(D'[i-1][j], M[i-1][j] - 400)       tie -> "from M"  (nwsmall.cpp:70-84:  DD > MD keeps D)
//       I'[i][j] = max(I'[i][j-1], M[i][j-1] - 400)       tie -> "from M"  (:93-102: MI >= II takes M)
//       best[i][j] = max(M, D', I')                       ties M, then D, then I (:113-141)
//   final edge: M, then D if D > M, then I if I > max (:645-656), compared without the close term.
//
// Parallelisation: one warp per region, rows in stripes of 32 lanes x R rows; the lanes sweep the
// columns as a skewed wavefront (lane l is l columns behind lane l-1), so every step needs only
// the previous step's bottom-row values of the lane above: three __shfl_up_sync (M-400, D', best)
// plus one for the column base.  Everything else lives in registers (3R ints per lane).  The
// stripe's last row goes through a small per-warp buffer (L2 resident) to feed the next stripe.
// Traceback bits: 4 per cell (2: predecessor state of M[i+1][j+1]; 1: D came from M; 1: I came
// from M), R cells = one 32-bit word per lane per step, stored step-major so that every step is
// one coalesced 128-byte store per warp: 0.5 B/cell, the only HBM traffic that scales with cells.
// Integer-ALU bound by design; no tensor cores (nothing here is a contraction).
#include "dp.cuh"

#include <algorithm>
#include <vector>

namespace mcu {

constexpr int NW_R = 8;
constexpr int NW_STRIPE = 32 * NW_R;
constexpr int NW_NINF = -(1 << 29);  // never beats a real score (|score| < 2^23), never overflows
constexpr int NW_WARPS = 8;
constexpr int NW_COOP_STRIPES = 3;  // regions with at least this many stripes are pipelined across the warps of a CTA
constexpr unsigned FULL = 0xffffffffu;

struct NwArgs {
    const u8* a;
    const u8* b;
    const u64* a_off;
    const u64* b_off;
    const u32* order;   // problems sorted by descending cell count
    u32 first, count;   // slice of `order` handled by this launch
    const uint2* units; // work units of this launch: one long region for the whole CTA, or up to NW_WARPS short ones
    u32 nunits;
    u32* tb;            // traceback words
    const u64* tb_off;  // per slot of `order` (absolute index): word offset into tb
    int4* boundary;     // per warp 2 x bstride entries {M-400, D', best(shifted by one column), -}
    u64 bstride;
    int4* result;       // per problem {M, D, I, -} at (la, lb)
    unsigned* counter;
    u32* err;
};

__device__ __forceinline__ u32 dna_code(u32 c, u32& bad)
{
    u32 u = c & 0xDFu;
    bad |= !(u == 'A' || u == 'C' || u == 'G' || u == 'T');
    return ((u >> 1) ^ (u >> 2)) & 3u;
}

// bytes a=0..3 of word c: NUC_SP[a][c] + 65 (MU/nucmx.cpp:8-25 with the +60 centre applied)
__device__ __forceinline__ u32 sub_column(u32 c)
{
    const u32 cA = 216u | (11u << 8) | (94u << 16) | (2u << 24);
    const u32 cC = 11u | (225u << 8) | (0u << 16) | (94u << 24);
    const u32 cG = 94u | (0u << 8) | (225u << 16) | (11u << 24);
    const u32 cT = 2u | (94u << 8) | (11u << 16) | (216u << 24);
    return c == 0 ? cA : (c == 1 ? cC : (c == 2 ? cG : cT));
}

// One column of R rows for this lane.  CAP additionally captures (M, D', I') of row cap_r.
template <bool CAP>
__device__ __forceinline__ u32 nw_rows(const u32 (&asel)[NW_R], int (&Ml)[NW_R], int (&Il)[NW_R], int (&Bp)[NW_R], u32 sb, int& upM, int& upD,
                                       int dg, int& out_best, int cap_r, int& capM, int& capD, int& capI)
{
    u32 tbw = 0;
    int best = 0;
    // Gap states are carried with a +400 bias (D'' = D' + 400, I'' = I' + 400) and M is carried unbiased, so neither gap recurrence
    // needs an add and best = max(max(D'', I'') - 400, M) is one fused add+max.  Every comparison below has the same constant on both
    // sides as the unbiased form (experiments/check_dp_biased_recurrence.py).
    // traceback nibble of row r: bit0 = best > M (= max(D', I') > M), bit1 = I'' > D'' (state of best: x = bit0 + (bit0 & bit1)),
    // bit2 = D'' came from M (not DD > MD), bit3 = I'' came from M (MI >= II).  One setp + one predicated or per bit.
#define NW_TB_BIT(cmp, lhs, rhs, bit) \
    asm("{\n\t.reg .pred p;\n\tsetp." cmp ".s32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(tbw) : "r"(lhs), "r"(rhs), "r"(bit))
#pragma unroll
    for (int r = 0; r < NW_R; ++r) {
        const int M = (int)__byte_perm(sb, 0u, asel[r]) + dg - 65;
        const int D = max(upD, upM);            // DD > MD: stay in D, else (ties too) come from M
        const int I = max(Ml[r], Il[r]);        // MI >= II: come from M
        const int DI = max(D, I);
        best = max(DI - 400, M);
        NW_TB_BIT("gt", best, M, 1u << (4 * r));
        NW_TB_BIT("gt", I, D, 2u << (4 * r));
        NW_TB_BIT("ge", upM, upD, 4u << (4 * r));
        NW_TB_BIT("ge", Ml[r], Il[r], 8u << (4 * r));
        if (CAP && r == cap_r) { capM = M; capD = D; capI = I; }
        dg = Bp[r];
        Bp[r] = best;
        upM = M;
        Ml[r] = M;
        Il[r] = I;
        upD = D;
    }
#undef NW_TB_BIT
    out_best = best;
    return tbw;
}

// One region.  `rank`/`team` = this warp's index and the number of warps sharing the region: warp k takes the stripes
// k, k+team, ... and, when team > 1, follows the warp working on the stripe above through `progress` (columns of that
// stripe whose bottom row is already in the boundary buffer), so a long region keeps a whole CTA busy as a pipeline
// of stripes instead of one warp.
struct NwProblem {
    const u8* A;
    const u8* B;
    u32 la, lb;
    u32* tb;
    int4* bnd[2];   // boundary rows, alternating by stripe parity
    int4* result;
};

__device__ __forceinline__ void nw_region(const NwProblem& P, u32 rank, u32 team, int4 (*ring)[32], volatile unsigned long long* progress, u32& bad)
{
    const u32 lane = threadIdx.x & 31;
    const u32 la = P.la, lb = P.lb;
    const u32 T = lb + 31;
    const u32 nstripes = (la + NW_STRIPE - 1) / NW_STRIPE;
    const u32 fin_lane = ((la - 1) % NW_STRIPE) / NW_R;
    const int fin_r = (int)((la - 1) % NW_R);
    int capM = 0, capD = 0, capI = 0;
    for (u32 s = rank; s < nstripes; s += team) {
        const u32 i0 = s * NW_STRIPE + lane * NW_R;
        u32 asel[NW_R];
        int Ml[NW_R], Il[NW_R], Bp[NW_R];
#pragma unroll
        for (int r = 0; r < NW_R; ++r) {
            u32 c = 0;
            if (i0 + r < la) c = dna_code(P.A[i0 + r], bad);
            asel[r] = 0x4440u | c;
            Ml[r] = NW_NINF;   // M[i][0]
            Il[r] = NW_NINF;   // I''[i][0]
            Bp[r] = -200;      // best[i][0]
        }
        int diag = -200, out_M = NW_NINF, out_D = NW_NINF, out_B = -200;
        u32 sb_cur = 0;
        const int4* bprev = P.bnd[(s + 1) & 1];
        int4* bcur = P.bnd[s & 1];
        const bool last = s + 1 == nstripes;
        const u32 cap_t = lb - 1 + fin_lane;
        u32* tbs = P.tb + (u64)s * T * 32 + lane;
        for (u32 t = 0; t < T; ++t) {
            if ((t & 31u) == 0) {
                // stage the inputs of lane 0 for columns t+1 .. t+32: the row above this stripe
                const u32 jk = t + 1 + lane;
                if (team > 1 && s > 0) {  // wait until the stripe above has published these columns
                    const u32 need = min(t + 32, lb);
                    const unsigned long long want = ((unsigned long long)s << 32) | need;  // tag = stripe index + 1 of the producer
                    while (true) {
                        const unsigned long long v = progress[(s - 1) & 31];
                        if ((v >> 32) == s && (u32)v >= need) break;
                        (void)want;
                    }
                    __syncwarp();
                }
                int4 v = make_int4(NW_NINF, NW_NINF, -200, 0);
                if (jk <= lb) {
                    u32 bb = 0;
                    v.w = (int)sub_column(dna_code(P.B[jk - 1], bb));
                    if (s == 0) bad |= bb;
                    if (s == 0) {
                        if (jk == 1 && la > 1) v.z = 0;
                    } else {
                        const int4 q = __ldcg(bprev + jk);
                        v.x = q.x;
                        v.y = q.y;
                        if (jk > 1) v.z = q.z;
                    }
                }
                ring[(t >> 5) & 1][lane] = v;
                __syncwarp();
            }
            int upM = __shfl_up_sync(FULL, out_M, 1);
            int upD = __shfl_up_sync(FULL, out_D, 1);
            const int upB = __shfl_up_sync(FULL, out_B, 1);
            u32 sb = __shfl_up_sync(FULL, sb_cur, 1);  // the substitution column travels with the wavefront
            int dg = diag;
            diag = upB;  // best[i0][j] of the lane above: the diagonal input of the next column
            if (lane == 0) {
                const int4 v = ring[(t >> 5) & 1][t & 31u];
                upM = v.x;
                upD = v.y;
                dg = v.z;
                sb = (u32)v.w;
            }
            sb_cur = sb;
            const int j = (int)t - (int)lane + 1;
            if (j >= 1 && j <= (int)lb) {
                u32 tbw;
                if (last && t == cap_t)
                    tbw = nw_rows<true>(asel, Ml, Il, Bp, sb, upM, upD, dg, out_B, fin_r, capM, capD, capI);
                else
                    tbw = nw_rows<false>(asel, Ml, Il, Bp, sb, upM, upD, dg, out_B, fin_r, capM, capD, capI);
                out_M = upM;
                out_D = upD;
                __stcs(tbs + (u64)t * 32, tbw);
                if (lane == 31 && !last) {
                    *(int2*)(bcur + j) = make_int2(out_M, out_D);
                    bcur[j + 1].z = out_B;
                }
            }
            if (team > 1 && !last && ((t & 31u) == 31u || t + 1 == T)) {
                // publish: lane 31 has finished columns 1 .. t-30
                __syncwarp();
                if (lane == 31 && t >= 31) {
                    __threadfence_block();
                    progress[s & 31] = ((unsigned long long)(s + 1) << 32) | (t - 30);
                }
            }
        }
        __threadfence_block();
        __syncwarp();
    }
    if ((nstripes - 1) % team == rank && lane == fin_lane) *P.result = make_int4(capM, capD - 200, capI - 200, 0);   // D = D'' - 200, I = I'' - 200
    __syncwarp();
}

__global__ void __launch_bounds__(NW_WARPS * 32, 2) nw_forward_kernel(NwArgs g)
{
    __shared__ int4 ring_s[NW_WARPS][2][32];
    __shared__ unsigned long long progress[32];
    __shared__ u32 s_unit;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gwarp = (u64)blockIdx.x * NW_WARPS + warp;
    const u64 cta_w0 = (u64)blockIdx.x * NW_WARPS;
    u32 bad = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_unit = atomicAdd(g.counter, 1u);
        if (threadIdx.x < 32) progress[threadIdx.x] = 0;
        __syncthreads();
        const u32 u = s_unit;
        if (u >= g.nunits) break;
        const uint2 unit = g.units[u];  // x = first slot in `order`, y = count (bit 31: cooperative)
        const bool coop = unit.y >> 31;
        const u32 count = unit.y & 0x7fffffffu;
        const u32 slot = coop ? unit.x : unit.x + warp;
        if (!coop && warp >= count) continue;
        const u32 p = g.order[slot];
        const u64 ao = g.a_off[p], bo = g.b_off[p];
        NwProblem P;
        P.A = g.a + ao;
        P.B = g.b + bo;
        P.la = (u32)(g.a_off[p + 1] - ao);
        P.lb = (u32)(g.b_off[p + 1] - bo);
        P.tb = g.tb + g.tb_off[slot];
        P.result = g.result + p;
        if (coop) {  // the CTA shares the boundary buffers of its first two warps
            P.bnd[0] = g.boundary + (cta_w0 * 2) * g.bstride;
            P.bnd[1] = g.boundary + (cta_w0 * 2 + 1) * g.bstride;
            nw_region(P, warp, NW_WARPS, ring_s[warp], progress, bad);
        } else {
            P.bnd[0] = g.boundary + (gwarp * 2) * g.bstride;
            P.bnd[1] = g.boundary + (gwarp * 2 + 1) * g.bstride;
            nw_region(P, 0, 1, ring_s[warp], progress, bad);
        }
        (void)lane;
    }
    if (bad) atomicOr(g.err, 1u);
}

// ---- BitTraceBack (MU/bittraceback.cpp:138-): one thread per region -----------------------
struct TbArgs {
    const u64* a_off;
    const u64* b_off;
    const u32* order;
    u32 first, count;
    const u32* tb;
    const u64* tb_off;
    const int4* result;
    const u64* path_off;
    char* path;       // device copy of the caller's path buffer
    u32* path_len;
    u64* path_start;  // where the (right-aligned) path starts inside its slot
    i64* score;
};

__device__ __forceinline__ u32 nw_nibble(const u32* __restrict__ tb, u32 T, u32 i, u32 j)
{
    const u32 row = i - 1;
    const u32 s = row / NW_STRIPE, lane = (row % NW_STRIPE) / NW_R, r = row % NW_R;
    const u32 t = (j - 1) + lane;
    const u32 w = __ldg(tb + ((u64)s * T + t) * 32 + lane);
    return (w >> (4 * r)) & 15u;
}

__global__ void __launch_bounds__(128) nw_traceback_kernel(TbArgs g)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.count) return;
    const u32 p = g.order[g.first + k];
    const u32 la = (u32)(g.a_off[p + 1] - g.a_off[p]), lb = (u32)(g.b_off[p + 1] - g.b_off[p]);
    const u32* tb = g.tb + g.tb_off[k];
    const u32 T = lb + 31;
    const int4 res = g.result[p];
    int sc = res.x;
    char edge = 'M';
    if (res.y > sc) { sc = res.y; edge = 'D'; }
    if (res.z > sc) { sc = res.z; edge = 'I'; }
    g.score[p] = sc;
    const u64 slot = g.path_off[p];
    char* out = g.path + slot + la + lb;
    u32 pa = la, pb = lb, n = 0;
    for (;;) {
        *--out = edge;
        ++n;
        char next;
        if (edge == 'M') {
            if (pa >= 2 && pb >= 2) {
                const u32 nb = nw_nibble(tb, T, pa - 1, pb - 1);
                const u32 x = (nb & 1u) + (nb & (nb >> 1) & 1u);   // bit0 = not M; bit1 = I beats D
                next = x == 0 ? 'M' : (x == 1 ? 'D' : 'I');
            } else if (pa >= 2) next = 'D';   // first column: reached through a leading gap in B (nwsmall.cpp:586-592)
            else next = 'I';                  // first row
            --pa;
            --pb;
        } else if (edge == 'D') {
            next = (pb >= 1 && (nw_nibble(tb, T, pa, pb) & 4u)) ? 'M' : 'D';
            --pa;
        } else {
            next = (pa >= 1 && (nw_nibble(tb, T, pa, pb) & 8u)) ? 'M' : 'I';
            --pb;
        }
        if (pa == 0 && pb == 0) break;
        edge = next;
        if ((edge == 'M' && (pa == 0 || pb == 0)) || (edge == 'D' && pa == 0) || (edge == 'I' && pb == 0)) {
            // cannot happen with a consistent matrix (the reference Quit()s); stop instead of running away
            n = 0;
            break;
        }
    }
    g.path_len[p] = n;
    g.path_start[p] = (u64)(out - g.path);
}

// moves every right-aligned path to the front of its slot: one warp per region
__global__ void __launch_bounds__(256) nw_shift_kernel(const u32* order, u32 first, u32 count, const u64* path_off, const u64* path_start,
                                                      const u32* path_len, char* path)
{
    const u32 lane = threadIdx.x & 31;
    const u64 wid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nw = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 k = wid; k < count; k += nw) {
        const u32 p = order[first + k];
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

// ---- host driver ---------------------------------------------------------------------------
struct NwState {
    DevBuf a, b, a_off, b_off, order, tb, tb_off, boundary, result, counter, path_off, path, path_len, path_start, score, units;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    u64 stats[5] = {0, 0, 0, 0, 0};
};
static NwState g_nw;

void nw_release()   // mcu_shutdown: give the cached device buffers back
{
    DevBuf* bufs[] = {&g_nw.a, &g_nw.b, &g_nw.a_off, &g_nw.b_off, &g_nw.order, &g_nw.tb, &g_nw.tb_off, &g_nw.boundary, &g_nw.result, &g_nw.counter,
                      &g_nw.path_off, &g_nw.path, &g_nw.path_len, &g_nw.path_start, &g_nw.score, &g_nw.units};
    for (DevBuf* b : bufs) b->release();
}

void nw_last_stats(u64* out5)
{
    for (int i = 0; i < 5; ++i) out5[i] = g_nw.stats[i];
}

// ---- INT32 issue-rate microbenchmark: the denominator of the DP roofline --------------------------------------------
// Register-resident: 8 independent add/max chains per thread, no memory traffic; the DP inner loop is made of exactly these
// instruction classes (VIADD / VIMNMX; ptxas fuses each add+max pair here into one VIADDMNMX).  Counted: 8 * iters INSTRUCTIONS per thread.
__global__ void __launch_bounds__(256) int32_peak_kernel(int* __restrict__ sink, int iters, int b, int c)
{
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
#define MCU_STEP(x) asm volatile("add.s32 %0, %0, %1;\n\tmax.s32 %0, %0, %2;" : "+r"(x) : "r"(b), "r"(c))
        MCU_STEP(a0); MCU_STEP(a1); MCU_STEP(a2); MCU_STEP(a3); MCU_STEP(a4); MCU_STEP(a5); MCU_STEP(a6); MCU_STEP(a7);
#undef MCU_STEP
    }
    const int r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x7fffffff) sink[0] = r;   // never true for the arguments used; keeps the chains alive
}

// G thread-level integer instructions per second sustained by the whole device (each one a fused add+max)
int int32_peak(double* gops_out, float* ms_out)
{
    static cudaStream_t stream = nullptr;   // its own stream: nw_batch creates g_nw's stream together with its events
    if (!stream) MCU_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    MCU_CUDA(cudaEventCreate(&e0));
    MCU_CUDA(cudaEventCreate(&e1));
    int* sink = nullptr;
    MCU_CUDA(cudaMalloc(&sink, 256));
    const int iters = 1 << 16, blocks = sm_count() * 8, threads = 256;
    int32_peak_kernel<<<blocks, threads, 0, stream>>>(sink, 64, 1, -5);   // warm-up
    MCU_CUDA(cudaEventRecord(e0, stream));
    int32_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters, 1, -5);
    MCU_CUDA(cudaEventRecord(e1, stream));
    MCU_CUDA(cudaStreamSynchronize(stream));
    MCU_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (ms_out) *ms_out = ms;
    if (gops_out) *gops_out = ms > 0.f ? 8.0 * (double)iters * (double)blocks * (double)threads / ((double)ms * 1e-3) / 1e9 : 0.0;
    return MCU_OK;
}

int nw_batch(u64 n, const char* a, const u64* a_off, const char* b, const u64* b_off, const u64* path_off, char* path_out, u32* path_len,
             i64* score, float* device_ms)
{
    NwState& st = g_nw;
    for (int i = 0; i < 5; ++i) st.stats[i] = 0;
    if (device_ms) *device_ms = 0.f;
    if (n == 0) return MCU_OK;
    if (!a || !b || !a_off || !b_off || !path_off || !path_out || !path_len || !score) { set_error("mcu_nw_batch: NULL pointer"); return MCU_EINVAL; }
    if (n >= 0xFFFFFFFFull) { set_error("mcu_nw_batch: too many regions"); return MCU_EINVAL; }
    u64 max_lb = 0;
    for (u64 i = 0; i < n; ++i) {
        if (a_off[i + 1] <= a_off[i] || b_off[i + 1] <= b_off[i]) { set_error("mcu_nw_batch: region %llu is empty", (unsigned long long)i); return MCU_EINVAL; }
        u64 la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        if (la > 0x3FFFFFFull || lb > 0x3FFFFFFull) { set_error("mcu_nw_batch: region %llu too long", (unsigned long long)i); return MCU_EINVAL; }
        if (path_off[i + 1] - path_off[i] < la + lb) { set_error("mcu_nw_batch: path slot %llu smaller than la+lb", (unsigned long long)i); return MCU_EINVAL; }
        max_lb = std::max(max_lb, lb);
    }
    if (!st.stream) {
        MCU_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        MCU_CUDA(cudaEventCreate(&st.e0));
        MCU_CUDA(cudaEventCreate(&st.e1));
    }
    cudaStream_t s = st.stream;
    const u64 abytes = a_off[n], bbytes = b_off[n], pbytes = path_off[n];

    // schedule: largest regions first (LPT), so the tail of a launch is made of small ones
    std::vector<u32> order(n);
    std::vector<u64> cells(n), tbw(n);
    u64 total_cells = 0;
    for (u64 i = 0; i < n; ++i) {
        order[i] = (u32)i;
        u64 la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        cells[i] = la * lb;
        total_cells += cells[i];
        tbw[i] = div_up(la, NW_STRIPE) * (lb + 31) * 32;
    }
    std::stable_sort(order.begin(), order.end(), [&](u32 x, u32 y) { return cells[x] > cells[y]; });

    size_t free_b = 0, total_b = 0;
    MCU_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const u64 fixed = abytes + bbytes + pbytes + n * (8 * 5 + 4 * 2 + 16 + 8) + (1ull << 28);
    u64 budget_words = free_b > fixed ? (u64)((free_b - fixed) * 0.8) / 4 : 0;
    const u64 cap_words = (96ull << 30) / 4;
    if (budget_words > cap_words) budget_words = cap_words;

    // sub-batches whose traceback words fit the budget
    std::vector<u64> tb_off(n);
    std::vector<std::pair<u32, u32>> batches;  // (first, count)
    u64 max_batch_words = 0;
    {
        u64 i = 0;
        while (i < n) {
            u64 words = 0, j = i;
            while (j < n && (j == i || words + tbw[order[j]] <= budget_words)) {
                tb_off[j] = words;
                words += tbw[order[j]];
                ++j;
            }
            if (words > budget_words) { set_error("mcu_nw_batch: one region needs %llu MiB of traceback, more than the device has free", (unsigned long long)(words >> 18)); return MCU_ENOMEM; }
            batches.push_back({(u32)i, (u32)(j - i)});
            max_batch_words = std::max(max_batch_words, words);
            i = j;
        }
    }

    const int ctas = sm_count() * 2;
    const u64 nwarps = (u64)ctas * NW_WARPS;
    const u64 bstride = max_lb + 4;
    // work units per sub-batch: a region of >= NW_COOP_STRIPES stripes occupies a whole CTA (its warps pipeline the stripes),
    // shorter regions are handed out NW_WARPS at a time (neighbours in the sorted order have similar sizes)
    std::vector<uint2> units;
    std::vector<std::pair<u32, u32>> unit_ranges;  // per sub-batch: (first unit, unit count)
    for (auto& bt : batches) {
        const u32 u0 = (u32)units.size();
        u32 i = bt.first;
        const u32 end = bt.first + bt.second;
        while (i < end) {
            const u64 la = a_off[order[i] + 1] - a_off[order[i]];
            if (div_up(la, NW_STRIPE) >= NW_COOP_STRIPES) {
                units.push_back(make_uint2(i, 1u | 0x80000000u));
                ++i;
            } else {
                const u32 c = std::min<u32>(NW_WARPS, end - i);
                units.push_back(make_uint2(i, c));
                i += c;
            }
        }
        unit_ranges.push_back({u0, (u32)units.size() - u0});
    }
    MCU_TRY(st.a.reserve(abytes + 16));
    MCU_TRY(st.b.reserve(bbytes + 16));
    MCU_TRY(st.a_off.reserve((n + 1) * 8));
    MCU_TRY(st.b_off.reserve((n + 1) * 8));
    MCU_TRY(st.path_off.reserve((n + 1) * 8));
    MCU_TRY(st.order.reserve(n * 4));
    MCU_TRY(st.tb_off.reserve(n * 8));
    MCU_TRY(st.result.reserve(n * 16));
    MCU_TRY(st.path_len.reserve(n * 4));
    MCU_TRY(st.path_start.reserve(n * 8));
    MCU_TRY(st.score.reserve(n * 8));
    MCU_TRY(st.path.reserve(pbytes + 16));
    MCU_TRY(st.counter.reserve(256));
    MCU_TRY(st.units.reserve(units.size() * sizeof(uint2) + 16));
    MCU_TRY(st.boundary.reserve(nwarps * 2 * bstride * sizeof(int4)));
    MCU_TRY(st.tb.reserve(max_batch_words * 4 + 16));

    MCU_CUDA(cudaMemcpyAsync(st.a.p, a, abytes, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.b.p, b, bbytes, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.a_off.p, a_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.b_off.p, b_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.path_off.p, path_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.order.p, order.data(), n * 4, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.tb_off.p, tb_off.data(), n * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.units.p, units.data(), units.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemsetAsync(st.counter.p, 0, 256, s));

    MCU_CUDA(cudaEventRecord(st.e0, s));
    u32* ctr = st.counter.as<u32>();  // [0..31] one work counter per sub-batch (mod 32, re-zeroed), [32] error flag
    int bi = 0;
    for (auto& bt : batches) {
        unsigned* counter = ctr + (bi % 32);
        if (bi >= 32) MCU_CUDA(cudaMemsetAsync(counter, 0, 4, s));
        NwArgs fa;
        fa.a = st.a.as<u8>(); fa.b = st.b.as<u8>();
        fa.a_off = st.a_off.as<u64>(); fa.b_off = st.b_off.as<u64>();
        fa.order = st.order.as<u32>(); fa.first = bt.first; fa.count = bt.second;
        fa.tb = st.tb.as<u32>(); fa.tb_off = st.tb_off.as<u64>();
        fa.units = st.units.as<uint2>() + unit_ranges[bi].first; fa.nunits = unit_ranges[bi].second;
        fa.boundary = st.boundary.as<int4>(); fa.bstride = bstride;
        fa.result = st.result.as<int4>(); fa.counter = counter; fa.err = ctr + 32;
        int grid = (int)std::min<u64>(fa.nunits, (u64)ctas);
        nw_forward_kernel<<<grid, NW_WARPS * 32, 0, s>>>(fa);
        TbArgs ta;
        ta.a_off = fa.a_off; ta.b_off = fa.b_off; ta.order = fa.order; ta.first = bt.first; ta.count = bt.second;
        ta.tb = fa.tb; ta.tb_off = fa.tb_off + bt.first; ta.result = fa.result; ta.path_off = st.path_off.as<u64>();
        ta.path = st.path.as<char>(); ta.path_len = st.path_len.as<u32>(); ta.path_start = st.path_start.as<u64>();
        ta.score = st.score.as<i64>();
        nw_traceback_kernel<<<(unsigned)div_up(bt.second, 128), 128, 0, s>>>(ta);
        u64 sg = std::min<u64>(div_up((u64)bt.second * 32, 256), (u64)sm_count() * 8);
        nw_shift_kernel<<<(unsigned)sg, 256, 0, s>>>(fa.order, bt.first, bt.second, ta.path_off, ta.path_start, ta.path_len, ta.path);
        st.stats[1] += 1;
        st.stats[2] += 2;
        ++bi;
    }
    MCU_CUDA(cudaEventRecord(st.e1, s));
    MCU_CUDA(cudaGetLastError());
    u32 err_flag = 0;
    MCU_CUDA(cudaMemcpyAsync(&err_flag, ctr + 32, 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(path_out, st.path.p, pbytes, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(path_len, st.path_len.p, n * 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(score, st.score.p, n * 8, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    if (err_flag) { set_error("mcu_nw_batch: letter outside ACGT (the integer-exact kernel covers the ACGT case only)"); return MCU_EALPHA; }
    for (u64 i = 0; i < n; ++i)
        if (path_len[i] == 0) { set_error("mcu_nw_batch: inconsistent traceback for region %llu", (unsigned long long)i); return MCU_ECUDA; }
    if (device_ms) cudaEventElapsedTime(device_ms, st.e0, st.e1);
    st.stats[0] = total_cells;
    st.stats[3] = batches.size();
    st.stats[4] = max_batch_words * 4;
    return MCU_OK;
}

}  // namespace mcu
