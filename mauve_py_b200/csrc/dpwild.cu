// Gapped DP for single-sequence profiles that contain DNA wildcards: NWSmall + BitTraceBack in the reference's float arithmetic.
//
// The integer kernels of dp.cu cover the case in which every profile column is one of A, C, G, T: there all of NWSmall's float
// scores are integers and an int32 wavefront is exact.  A wildcard column (N, X, R, Y, ...) makes the substitution scores
// non-dyadic floats, and then parity needs the reference's own operations: every product, sum and comparison
// in float32, in the reference's expression order, no FMA contraction.  This file does exactly that, one THREAD per region:
// such regions are rare (a window of a finished genome seldom holds an N) and small, so the point is coverage and exactness,
// not throughput; all values are decided by __host__ __device__ functions that tests/ also run on the CPU (-DMCU_HOST_EMU).
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

enum { NWF_BIT_DM = 1, NWF_BIT_IM = 2, NWF_BIT_xM = 3, NWF_BIT_MD = 4, NWF_BIT_MI = 8 };   // MU/types.h:28-44

// One region.  sub: 36 substitution scores [class a][class b]; rows: 4 * (lb + 1) floats of scratch; tb: (la + 1) * (lb + 1) bytes,
// zeroed; path: la + lb bytes.  Returns the path length (edges 'M' 'D' 'I', first edge first), 0 on an inconsistent traceback.
NWF_HD u32 nwf_align_one(const u8* __restrict__ a, u32 la, const u8* __restrict__ b, u32 lb, const float* __restrict__ sub, float* __restrict__ rows,
                         u8* __restrict__ tb, char* __restrict__ path, float* score_out)
{
    const u64 W = (u64)lb + 1;
    float* MPrev = rows;
    float* MCurr = rows + W;
    float* MNext = rows + 2 * W;
    float* DRow = rows + 3 * W;
    const float NINF = -1e37f, gap = -200.0f;   // 1.0f * -400 / 2
#define NWF_S(i, j) sub[6 * nwf_class(a[i]) + nwf_class(b[j])]
#define NWF_OPEN_A(i) ((i) == 0 ? -0.0f : gap)
#define NWF_OPEN_B(j) ((j) == 0 ? -0.0f : gap)
#define NWF_CLOSE_A(i) (((i) == la - 1 && la > 1) ? -0.0f : gap)
#define NWF_CLOSE_B(j) (((j) == lb - 1 && lb > 1) ? -0.0f : gap)
#define NWF_TB(i, j) tb[(u64)(i) * W + (j)]
#define NWF_REC_D(i, j) { const float DD = NWF_ADD(DRow[j], 0.0f), MD = NWF_ADD(MPrev[j], NWF_OPEN_A((i) - 1)); \
        if (DD > MD) DRow[j] = DD; else { DRow[j] = MD; NWF_TB(i, j) |= NWF_BIT_MD; } }
#define NWF_REC_I(i, j) { Iij = NWF_ADD(Iij, 0.0f); const float MI = NWF_ADD(MCurr[(j) - 1], NWF_OPEN_B((j) - 1)); \
        if (MI >= Iij) { Iij = MI; NWF_TB(i, j) |= NWF_BIT_MI; } }
#define NWF_REC_M(i, j) { const float DM = NWF_ADD(DRow[j], NWF_CLOSE_A((i) - 1)), IM = NWF_ADD(Iij, NWF_CLOSE_B((j) - 1)), MM = MCurr[j]; \
        if (MM >= DM && MM >= IM) MNext[(j) + 1] = NWF_ADD(MNext[(j) + 1], MM); \
        else if (DM >= MM && DM >= IM) { MNext[(j) + 1] = NWF_ADD(MNext[(j) + 1], DM); NWF_TB((i) + 1, (j) + 1) |= NWF_BIT_DM; } \
        else { MNext[(j) + 1] = NWF_ADD(MNext[(j) + 1], IM); NWF_TB((i) + 1, (j) + 1) |= NWF_BIT_IM; } }
    float Iij;
    for (u32 j = 0; j <= lb; ++j) DRow[j] = NINF;
    MPrev[0] = 0.0f;
    for (u32 j = 1; j <= lb; ++j) MPrev[j] = NINF;
    MCurr[0] = NINF;
    MCurr[1] = NWF_S(0, 0);
    for (u32 j = 2; j <= lb; ++j) {
        MCurr[j] = NWF_ADD(NWF_ADD(NWF_ADD(NWF_S(0, j - 1), NWF_OPEN_B(0)), 0.0f), NWF_CLOSE_B(j - 2));
        NWF_TB(1, j) |= NWF_BIT_IM;
    }
    for (u32 i = 1; i < la; ++i) {
        Iij = NINF;
        DRow[0] = NWF_ADD(NWF_OPEN_A(0), 0.0f);
        MCurr[0] = NINF;
        if (i == 1) { MCurr[1] = NWF_S(0, 0); NWF_TB(i, 1) &= (u8)~NWF_BIT_xM; }
        else {
            MCurr[1] = NWF_ADD(NWF_ADD(NWF_ADD(NWF_S(i - 1, 0), NWF_OPEN_A(0)), 0.0f), NWF_CLOSE_A(i - 2));
            NWF_TB(i, 1) = (u8)((NWF_TB(i, 1) & ~NWF_BIT_xM) | NWF_BIT_DM);
        }
        for (u32 j = 1; j < lb; ++j) MNext[j + 1] = NWF_S(i, j);
        for (u32 j = 1; j < lb; ++j) { NWF_REC_D(i, j) NWF_REC_I(i, j) NWF_REC_M(i, j) }
        NWF_REC_D(i, lb) NWF_REC_I(i, lb)
        float* t = MPrev; MPrev = MCurr; MCurr = MNext; MNext = t;   // Rotate
    }
    MCurr[0] = NINF;
    if (la > 1) MCurr[1] = NWF_ADD(NWF_ADD(NWF_ADD(NWF_S(la - 1, 0), 0.0f), NWF_OPEN_A(0)), NWF_CLOSE_A(la - 2));
    else MCurr[1] = NWF_ADD(NWF_ADD(NWF_S(la - 1, 0), NWF_OPEN_A(0)), NWF_CLOSE_A(0));
    NWF_TB(la, 1) = (u8)((NWF_TB(la, 1) & ~NWF_BIT_xM) | NWF_BIT_DM);
    DRow[0] = NINF;
    for (u32 j = 1; j <= lb; ++j) NWF_REC_D(la, j)
    Iij = NINF;
    for (u32 j = 1; j <= lb; ++j) NWF_REC_I(la, j)
    const float MAB = MCurr[lb], DAB = DRow[lb], IAB = Iij;
    float Score = MAB;
    char edge = 'M';
    if (DAB > Score) { Score = DAB; edge = 'D'; }
    if (IAB > Score) { Score = IAB; edge = 'I'; }
    if (score_out) *score_out = Score;
    // BitTraceBack: emitted backwards, reversed in place afterwards
    u32 pa = la, pb = lb, plen = 0;
    for (;;) {
        const u8 bits = NWF_TB(pa, pb);
        char next;
        if (plen >= la + lb) return 0;
        path[plen++] = edge;
        if (edge == 'M') {
            if (pa == 0 || pb == 0) return 0;
            next = (bits & NWF_BIT_xM) == 0 ? 'M' : ((bits & NWF_BIT_xM) == NWF_BIT_DM ? 'D' : 'I');
            --pa; --pb;
        } else if (edge == 'D') {
            if (pa == 0) return 0;
            next = (bits & NWF_BIT_MD) ? 'M' : 'D';
            --pa;
        } else {
            if (pb == 0) return 0;
            next = (bits & NWF_BIT_MI) ? 'M' : 'I';
            --pb;
        }
        if (pa == 0 && pb == 0) break;
        edge = next;
    }
    for (u32 x = 0; x < plen / 2; ++x) { const char c = path[x]; path[x] = path[plen - 1 - x]; path[plen - 1 - x] = c; }
    return plen;
#undef NWF_S
#undef NWF_OPEN_A
#undef NWF_OPEN_B
#undef NWF_CLOSE_A
#undef NWF_CLOSE_B
#undef NWF_TB
#undef NWF_REC_D
#undef NWF_REC_I
#undef NWF_REC_M
}

#ifndef MCU_HOST_EMU
struct NwfArgs {
    const u8* a;
    const u8* b;
    const u64* a_off;
    const u64* b_off;
    const u64* rows_off;   // floats
    const u64* tb_off;     // bytes
    const u64* path_off;
    float* rows;
    u8* tb;
    char* path;
    u32* path_len;
    float* score;
    u32 n;
};

__global__ void __launch_bounds__(64) nw_wild_kernel(const __grid_constant__ NwfArgs g)
{
    __shared__ float sub[36];
    if (threadIdx.x < 36) sub[threadIdx.x] = nwf_pair_score((int)threadIdx.x / 6, (int)threadIdx.x % 6);
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
        const u32 la = (u32)(g.a_off[i + 1] - g.a_off[i]), lb = (u32)(g.b_off[i + 1] - g.b_off[i]);
        g.path_len[i] = nwf_align_one(g.a + g.a_off[i], la, g.b + g.b_off[i], lb, sub, g.rows + g.rows_off[i], g.tb + g.tb_off[i],
                                      g.path + g.path_off[i], g.score + i);
    }
}

static DevBuf w_a, w_b, w_aoff, w_boff, w_rowsoff, w_tboff, w_pathoff, w_rows, w_tb, w_path, w_plen, w_score;
static cudaStream_t w_stream = nullptr;

// cells of one region above which the caller should use the reference's code (one thread walks the whole matrix)
static const u64 NWF_MAX_CELLS = 16ull << 20;

int nw_batch_wild(u64 n, const char* a, const u64* a_off, const char* b, const u64* b_off, const u64* path_off, char* path_out, u32* path_len,
                  float* score, float* device_ms)
{
    if (device_ms) *device_ms = 0.f;
    if (n == 0) return MCU_OK;
    if (!a || !b || !a_off || !b_off || !path_off || !path_out || !path_len || !score) { set_error("mcu_nw_batch_wild: NULL pointer"); return MCU_EINVAL; }
    if (n >= 0xFFFFFFFFull) { set_error("mcu_nw_batch_wild: too many regions"); return MCU_EINVAL; }
    std::vector<u64> rows_off(n + 1, 0), tb_off(n + 1, 0);
    for (u64 i = 0; i < n; ++i) {
        if (a_off[i + 1] <= a_off[i] || b_off[i + 1] <= b_off[i]) { set_error("mcu_nw_batch_wild: region %llu is empty", (unsigned long long)i); return MCU_EINVAL; }
        const u64 la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        if (la * lb > NWF_MAX_CELLS) { set_error("mcu_nw_batch_wild: region %llu has more than %llu cells", (unsigned long long)i, (unsigned long long)NWF_MAX_CELLS); return MCU_EINVAL; }
        if (path_off[i + 1] - path_off[i] < la + lb) { set_error("mcu_nw_batch_wild: path slot %llu smaller than la+lb", (unsigned long long)i); return MCU_EINVAL; }
        for (u64 k = a_off[i]; k < a_off[i + 1]; ++k)
            if (nwf_class((u8)a[k]) < 0) { set_error("mcu_nw_batch_wild: byte %d is not a DNA letter or wildcard", (int)(u8)a[k]); return MCU_EALPHA; }
        for (u64 k = b_off[i]; k < b_off[i + 1]; ++k)
            if (nwf_class((u8)b[k]) < 0) { set_error("mcu_nw_batch_wild: byte %d is not a DNA letter or wildcard", (int)(u8)b[k]); return MCU_EALPHA; }
        rows_off[i + 1] = rows_off[i] + 4 * (lb + 1);
        tb_off[i + 1] = tb_off[i] + (la + 1) * (lb + 1);
    }
    if (!w_stream) MCU_CUDA(cudaStreamCreateWithFlags(&w_stream, cudaStreamNonBlocking));
    cudaStream_t s = w_stream;
    const u64 abytes = a_off[n], bbytes = b_off[n], pbytes = path_off[n];
    MCU_TRY(w_a.reserve(abytes + 16));
    MCU_TRY(w_b.reserve(bbytes + 16));
    MCU_TRY(w_aoff.reserve((n + 1) * 8));
    MCU_TRY(w_boff.reserve((n + 1) * 8));
    MCU_TRY(w_rowsoff.reserve((n + 1) * 8));
    MCU_TRY(w_tboff.reserve((n + 1) * 8));
    MCU_TRY(w_pathoff.reserve((n + 1) * 8));
    MCU_TRY(w_rows.reserve(rows_off[n] * sizeof(float) + 16));
    MCU_TRY(w_tb.reserve(tb_off[n] + 16));
    MCU_TRY(w_path.reserve(pbytes + 16));
    MCU_TRY(w_plen.reserve(n * 4));
    MCU_TRY(w_score.reserve(n * 4));
    MCU_CUDA(cudaMemcpyAsync(w_a.p, a, abytes, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_b.p, b, bbytes, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_aoff.p, a_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_boff.p, b_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_rowsoff.p, rows_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_tboff.p, tb_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(w_pathoff.p, path_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemsetAsync(w_tb.p, 0, tb_off[n], s));
    MCU_CUDA(cudaMemsetAsync(w_path.p, 0, pbytes, s));
    cudaEvent_t e0, e1;
    MCU_CUDA(cudaEventCreate(&e0));
    MCU_CUDA(cudaEventCreate(&e1));
    NwfArgs g;
    g.a = w_a.as<u8>(); g.b = w_b.as<u8>();
    g.a_off = w_aoff.as<u64>(); g.b_off = w_boff.as<u64>();
    g.rows_off = w_rowsoff.as<u64>(); g.tb_off = w_tboff.as<u64>(); g.path_off = w_pathoff.as<u64>();
    g.rows = w_rows.as<float>(); g.tb = w_tb.as<u8>(); g.path = w_path.as<char>();
    g.path_len = w_plen.as<u32>(); g.score = w_score.as<float>();
    g.n = (u32)n;
    MCU_CUDA(cudaEventRecord(e0, s));
    nw_wild_kernel<<<(unsigned)std::min<u64>(div_up(n, 64), (u64)sm_count() * 16), 64, 0, s>>>(g);
    MCU_CUDA(cudaEventRecord(e1, s));
    MCU_CUDA(cudaGetLastError());
    MCU_CUDA(cudaMemcpyAsync(path_out, w_path.p, pbytes, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(path_len, w_plen.p, n * 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(score, w_score.p, n * 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    if (device_ms) cudaEventElapsedTime(device_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (u64 i = 0; i < n; ++i)
        if (path_len[i] == 0) { set_error("mcu_nw_batch_wild: inconsistent traceback for region %llu", (unsigned long long)i); return MCU_ECUDA; }
    return MCU_OK;
}

#else  // MCU_HOST_EMU ----------------------------------------------------------------------------------------------------
}  // namespace mcu
extern "C" {
// TEST-ONLY host driver (tests/_emu.py): one region through the value functions above; returns the path length
__attribute__((visibility("default"))) long long emu_nw_wild(const char* a, unsigned la, const char* b, unsigned lb, char* path_out, float* score_out)
{
    for (unsigned i = 0; i < la; ++i) if (mcu::nwf_class((u8)a[i]) < 0) return -1;
    for (unsigned j = 0; j < lb; ++j) if (mcu::nwf_class((u8)b[j]) < 0) return -1;
    if (la == 0 || lb == 0) return -1;
    float sub[36];
    for (int k = 0; k < 36; ++k) sub[k] = mcu::nwf_pair_score(k / 6, k % 6);
    float* rows = (float*)malloc(4 * ((size_t)lb + 1) * sizeof(float));
    memset(rows, 0xFF, 4 * ((size_t)lb + 1) * sizeof(float));   // NaN: the device scratch is not initialised either, nothing may depend on it
    u8* tb = (u8*)calloc(((size_t)la + 1) * ((size_t)lb + 1), 1);
    const unsigned n = mcu::nwf_align_one((const u8*)a, la, (const u8*)b, lb, sub, rows, tb, path_out, score_out);
    free(rows);
    free(tb);
    return n ? (long long)n : -1;
}
}
namespace mcu {
#endif

}  // namespace mcu
