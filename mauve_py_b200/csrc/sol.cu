// Seed occurrence list and anchor scores: the consumers of the sorted mer list after matching.
//
//   SeedOccurrenceList::construct   LM/SeedOccurrenceList.h:22-78   -> sol_count_kernel + sol_smooth_kernel
//   smoothFrequencies               LM/SeedOccurrenceList.h:103-119
//   GetPairwiseAnchorScore          LM/GreedyBreakpointElimination.h:403-476 -> anchor_score_kernel + anchor_lcb_kernel
//   (GetAlignment LM/AbstractGappedAlignment.h:84-104, computeMatchScores LM/Scoring.h:119-139)
//
// The reference walks the sorted list once more (one GetSeedMer per rank: a third full pass over the genome) and then
// smooths sequentially; here the sorted keys are already in HBM after the SML build, multiplicities are a local test on
// them, and the window mean is a gather.  Arithmetic follows the reference bit for bit: counts are floats, the window sum
// is an exact integer in a double, the mean is one IEEE double division rounded once to float.
//
// Everything that decides a VALUE lives in __host__ __device__ functions, so that tests/ can run exactly this code on the
// CPU (built with -DMCU_HOST_EMU into a test-only library; the product library has no host path).
#include "sol.cuh"

#include <algorithm>

namespace mcu {

#define SOL_HD __host__ __device__ __forceinline__

// ---- multiplicity of the seed at sorted rank i ---------------------------------------------------------------------
// keys are sorted ascending, key = canon << 2 | genome << 1 | strand (anchor.cu seedgen_kernel), so `key >> 2` is the
// reference's `mer & seed_mask` (SeedOccurrenceList.h:43) and equal seeds are adjacent.  Short runs (almost all of them)
// are measured by walking; a run that outlasts the walk is bounded by bisection (the list is sorted).
#define SOL_WALK 8
template <typename K>
SOL_HD u32 sol_run_length(const K* __restrict__ keys, u64 n, u64 i)
{
    const K k = keys[i] >> 2;
    u64 lo = i, hi = i + 1;
    int steps = 0;
    while (lo > 0 && steps < SOL_WALK && (K)(keys[lo - 1] >> 2) == k) { --lo; ++steps; }
    if (lo > 0 && (K)(keys[lo - 1] >> 2) == k) {  // first rank whose seed is >= k
        u64 a = 0, b = lo;
        while (a < b) {
            const u64 mid = a + ((b - a) >> 1);
            if ((K)(keys[mid] >> 2) < k) a = mid + 1; else b = mid;
        }
        lo = a;
    }
    steps = 0;
    while (hi < n && steps < SOL_WALK && (K)(keys[hi] >> 2) == k) { ++hi; ++steps; }
    if (hi < n && (K)(keys[hi] >> 2) == k) {  // first rank whose seed is > k
        u64 a = hi, b = n;
        while (a < b) {
            const u64 mid = a + ((b - a) >> 1);
            if ((K)(keys[mid] >> 2) <= k) a = mid + 1; else b = mid;
        }
        hi = a;
    }
    return (u32)(hi - lo);
}

// count[j] as the reference holds it before smoothing: (float)multiplicity for seed start positions, 1 for the last L-1
// positions (:57-58) and for the virtual positions before the sequence start (:108-110)
SOL_HD u64 sol_term(const u32* __restrict__ raw, u64 npos, i64 j) { return (j < 0 || (u64)j >= npos) ? 1ull : (u64)(float)raw[j]; }

SOL_HD float sol_mean(u64 sum, int L)
{
#ifdef __CUDA_ARCH__
    return __double2float_rn(__ddiv_rn((double)sum, (double)L));
#else
    return (float)((double)sum / (double)L);
#endif
}

// smoothed frequency of position k (0 <= k < n): mean of the counts of the L seeds starting at k-L+1 .. k; the last
// position keeps its raw count (the loop at :111 assigns count[i-1] for i < Length() only); zeros become 1 (:63-65)
SOL_HD float sol_frequency(const u32* __restrict__ raw, u64 npos, u64 n, int L, u64 k)
{
    float f;
    if (k + 1 < n) {
        u64 sum = 0;
        for (int j = 0; j < L; ++j) sum += sol_term(raw, npos, (i64)k - j);
        f = sol_mean(sum, L);
    } else {
        f = (float)sol_term(raw, npos, (i64)k);
    }
    return f == 0.0f ? 1.0f : f;
}

// ---- anchor score of one column ----------------------------------------------------------------------------------------
// SortedMerList::BasicDNATable (LM/SortedMerList.cpp:29-47) of a byte, and of its image under gnFilter::DNAComplementFilter
// (libGenome/gnFilter.cpp:509-545; unmapped bytes count as code 0), as 2-bit entries indexed by the letter.
#define SOL_LUT(l, v) ((u64)(v) << (2 * ((l) - 'a')))
#define SOL_FWD_LUT (SOL_LUT('b', 1) | SOL_LUT('c', 1) | SOL_LUT('y', 1) | SOL_LUT('g', 2) | SOL_LUT('s', 2) | SOL_LUT('k', 2) | SOL_LUT('t', 3))
#define SOL_RC_LUT (SOL_LUT('g', 1) | SOL_LUT('v', 1) | SOL_LUT('r', 1) | SOL_LUT('c', 2) | SOL_LUT('s', 2) | SOL_LUT('m', 2) | SOL_LUT('a', 3))
SOL_HD unsigned sol_dna_code(u8 c, bool revcomp)
{
    const unsigned x = ((unsigned)c | 0x20u) - (unsigned)'a';
    const u64 lut = revcomp ? SOL_RC_LUT : SOL_FWD_LUT;
    return x < 26u ? (unsigned)((lut >> (2 * x)) & 3u) : 0u;
}

struct AnchorScoreArgs {
    const u8* seq[2];
    const float* freq[2];
    const mcu_match* rows;
    u64 n_rows;
    int matrix[16];
    int penalize_repeats;
    i64* match_score;
};

// GreedyBreakpointElimination.h:436-456 for column c of an ungapped match: genome g contributes base left_g-1+c (forward)
// or the complement of base left_g-1+len-1-c (reverse strand, GetAlignment's ReverseFilter); the uniqueness product is
// read at left_g-1+c for BOTH orientations (:441-442), as the reference does.
SOL_HD int anchor_col_score(const AnchorScoreArgs& a, const int* __restrict__ matrix, u64 left0, u64 left1, u64 len, bool rev0, bool rev1, u64 c)
{
    const unsigned t0 = sol_dna_code(a.seq[0][rev0 ? left0 - 1 + len - 1 - c : left0 - 1 + c], rev0);
    const unsigned t1 = sol_dna_code(a.seq[1][rev1 ? left1 - 1 + len - 1 - c : left1 - 1 + c], rev1);
    int score = matrix[4 * t0 + t1];
    if (score > 0) {
#ifdef __CUDA_ARCH__
        float uniprod = __fmul_rn(a.freq[0][left0 - 1 + c], a.freq[1][left1 - 1 + c]);
        if (uniprod == 0.0f) uniprod = 1.0f;
        if (a.penalize_repeats) score = (int)__dmul_rn((double)score, __ddiv_rn(2.0, (double)uniprod)) - score;
        else score = (int)__fdiv_rn((float)score, uniprod);
#else
        volatile float uniprod = a.freq[0][left0 - 1 + c] * a.freq[1][left1 - 1 + c];
        if (uniprod == 0.0f) uniprod = 1.0f;
        if (a.penalize_repeats) {
            volatile double q = 2.0 / (double)uniprod;
            volatile double pr = (double)score * q;
            score = (int)pr - score;
        } else {
            volatile float q = (float)score / uniprod;
            score = (int)q;
        }
#endif
    }
    return score;
}

#ifndef MCU_HOST_EMU
// =========================================================================================
// kernels
// =========================================================================================
// HBM traffic: sizeof(K)+4 B read per rank (neighbours come from L1), one scattered 4-byte write per position.
template <typename K>
__global__ void __launch_bounds__(256) sol_count_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, u64 npos, u32* __restrict__ raw)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < npos; i += (u64)gridDim.x * blockDim.x) raw[vals[i]] = sol_run_length<K>(keys, npos, i);
}

// HBM traffic: 4 B read + 4 B written per position (the L-wide window is served by L1: consecutive lanes read consecutive words).
__global__ void __launch_bounds__(256) sol_smooth_kernel(const u32* __restrict__ raw, u64 npos, u64 n, int L, float* __restrict__ freq)
{
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) freq[k] = sol_frequency(raw, npos, n, L, k);
}

// one warp per match: lanes stride the columns, the integer column scores are summed exactly (the reference adds them into a
// double, :459-462; integers this small are exact there too).  16 B/column read (two bases, two frequencies), 8 B/match written.
__global__ void __launch_bounds__(256) anchor_score_kernel(const __grid_constant__ AnchorScoreArgs a)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    __shared__ int matrix[16];  // indexed by the bases: kept out of the (constant-bank) parameter block
    if (threadIdx.x < 16) matrix[threadIdx.x] = a.matrix[threadIdx.x];
    __syncthreads();
    for (u64 k = warp; k < a.n_rows; k += nwarps) {
        const mcu_match m = a.rows[k];
        const bool rev0 = m.start0 < 0, rev1 = m.start1 < 0;
        const u64 left0 = (u64)(rev0 ? -m.start0 : m.start0), left1 = (u64)(rev1 ? -m.start1 : m.start1), len = (u64)m.len;
        i64 sum = 0;
        for (u64 c = lane; c < len; c += 32) sum += anchor_col_score(a, matrix, left0, left1, len, rev0, rev1, c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
        if (lane == 0) a.match_score[k] = sum;
    }
}

// lcb_score = sum of its matches' scores, accumulated in a double in list order (:464)
__global__ void anchor_lcb_kernel(const i64* __restrict__ match_score, const u64* __restrict__ lcb_off, u64 n_lcb, double* __restrict__ lcb_score)
{
    for (u64 l = (u64)blockIdx.x * blockDim.x + threadIdx.x; l < n_lcb; l += (u64)gridDim.x * blockDim.x) {
        double s = 0;
        for (u64 k = lcb_off[l]; k < lcb_off[l + 1]; ++k) s = __dadd_rn(s, (double)match_score[k]);
        lcb_score[l] = s;
    }
}

static int sol_grid(u64 n, int block, int per_sm)
{
    const u64 want = div_up(n ? n : 1, (u64)block);
    const u64 cap = (u64)sm_count() * per_sm;
    return (int)std::max<u64>(1, std::min(want, cap));
}

// ---- host side ---------------------------------------------------------------------------------------------------------
// sorted list of `seq` (pack + seedgen + radix sort, anchor.cu) -> s.sol_freq[slot] (n floats on the device)
int sol_build_device(Session& s, const char* seq, u64 n, u64 seed, int slot)
{
    SeedParams sp;
    MCU_TRY(make_seed_params(seed, &sp));
    if (n == 0) return MCU_OK;
    u64 npos = 0;
    MCU_TRY(sml_build_device(s, seq, n, seed, nullptr, nullptr, nullptr, &npos));
    if (npos != s.sml_npos) { set_error("sol_build: sorted list state out of step"); return MCU_EINVAL; }
    MCU_TRY(s.sol_raw.reserve((n + 1) * sizeof(u32)));
    MCU_TRY(s.sol_freq[slot].reserve((n + 1) * sizeof(float)));
    if (npos) {
        if (s.sml_key_bytes == 4)
            sol_count_kernel<u32><<<sol_grid(npos, 256, 8), 256, 0, s.stream>>>((const u32*)s.sml_keys, s.sml_vals, npos, s.sol_raw.as<u32>());
        else
            sol_count_kernel<u64><<<sol_grid(npos, 256, 8), 256, 0, s.stream>>>((const u64*)s.sml_keys, s.sml_vals, npos, s.sol_raw.as<u32>());
        s.launches++;
    }
    sol_smooth_kernel<<<sol_grid(n, 256, 8), 256, 0, s.stream>>>(s.sol_raw.as<u32>(), npos, n, sp.L, s.sol_freq[slot].as<float>());
    s.launches++;
    MCU_CUDA(cudaGetLastError());
    return MCU_OK;
}

int sol_build(Session& s, const char* seq, u64 n, u64 seed, float* freq_out)
{
    MCU_TRY(sol_build_device(s, seq, n, seed, 0));
    if (n) MCU_CUDA(cudaMemcpyAsync(freq_out, s.sol_freq[0].p, n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    return MCU_OK;
}

static const int kHoxd[16] = {91, -114, -31, -123, -114, 100, -125, -31, -31, -125, 100, -114, -123, -31, -114, 91};  // LM/SubstitutionMatrix.h:23-33

int anchor_scores(Session& s, const char* seq0, u64 n0, const char* seq1, u64 n1, u64 seed, const float* freq0, const float* freq1,
                  const mcu_match* rows, u64 n_rows, const u64* lcb_off, u64 n_lcb, const int* matrix, int penalize_repeats,
                  double* lcb_score_out, i64* match_score_out)
{
    const char* seq[2] = {seq0, seq1};
    const u64 n[2] = {n0, n1};
    const float* freq[2] = {freq0, freq1};
    // rows are checked on the host: the kernel indexes the genomes with them
    for (u64 k = 0; k < n_rows; ++k) {
        const i64 len = rows[k].len;
        const u64 l0 = (u64)(rows[k].start0 < 0 ? -rows[k].start0 : rows[k].start0), l1 = (u64)(rows[k].start1 < 0 ? -rows[k].start1 : rows[k].start1);
        if (len < 0 || l0 < 1 || l1 < 1 || l0 - 1 + (u64)len > n0 || l1 - 1 + (u64)len > n1) {
            set_error("mcu_anchor_scores: row %llu lies outside the sequences", (unsigned long long)k);
            return MCU_EINVAL;
        }
    }
    for (u64 l = 0; l < n_lcb; ++l)
        if (lcb_off[l + 1] < lcb_off[l] || lcb_off[l + 1] > n_rows) { set_error("mcu_anchor_scores: bad LCB offsets at %llu", (unsigned long long)l); return MCU_EINVAL; }
    for (int g = 0; g < 2; ++g) {
        if (freq[g]) {
            MCU_TRY(s.sol_freq[g].reserve((n[g] + 1) * sizeof(float)));
            if (n[g]) MCU_CUDA(cudaMemcpyAsync(s.sol_freq[g].p, freq[g], n[g] * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        } else {
            MCU_TRY(sol_build_device(s, seq[g], n[g], seed, g));
        }
    }
    MCU_TRY(session_upload(s, seq0, n0, seq1, n1));
    MCU_TRY(s.as_rows.reserve((n_rows + 1) * sizeof(mcu_match)));
    MCU_TRY(s.as_match.reserve((n_rows + 1) * sizeof(i64)));
    MCU_TRY(s.as_off.reserve((n_lcb + 2) * sizeof(u64)));
    MCU_TRY(s.as_lcb.reserve((n_lcb + 1) * sizeof(double)));
    if (n_rows) MCU_CUDA(cudaMemcpyAsync(s.as_rows.p, rows, n_rows * sizeof(mcu_match), cudaMemcpyHostToDevice, s.stream));
    if (n_lcb) MCU_CUDA(cudaMemcpyAsync(s.as_off.p, lcb_off, (n_lcb + 1) * sizeof(u64), cudaMemcpyHostToDevice, s.stream));
    AnchorScoreArgs a;
    a.seq[0] = s.ascii[0].as<u8>();
    a.seq[1] = s.ascii[1].as<u8>();
    a.freq[0] = s.sol_freq[0].as<float>();
    a.freq[1] = s.sol_freq[1].as<float>();
    a.rows = s.as_rows.as<mcu_match>();
    a.n_rows = n_rows;
    for (int i = 0; i < 16; ++i) a.matrix[i] = matrix ? matrix[i] : kHoxd[i];
    a.penalize_repeats = penalize_repeats ? 1 : 0;
    a.match_score = s.as_match.as<i64>();
    if (n_rows) {
        anchor_score_kernel<<<sol_grid(n_rows * 32, 256, 8), 256, 0, s.stream>>>(a);
        s.launches++;
    }
    if (n_lcb) {
        anchor_lcb_kernel<<<sol_grid(n_lcb, 128, 8), 128, 0, s.stream>>>(s.as_match.as<i64>(), s.as_off.as<u64>(), n_lcb, s.as_lcb.as<double>());
        s.launches++;
        MCU_CUDA(cudaMemcpyAsync(lcb_score_out, s.as_lcb.p, n_lcb * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    }
    if (match_score_out && n_rows) MCU_CUDA(cudaMemcpyAsync(match_score_out, s.as_match.p, n_rows * sizeof(i64), cudaMemcpyDeviceToHost, s.stream));
    MCU_CUDA(cudaStreamSynchronize(s.stream));
    MCU_CUDA(cudaGetLastError());
    return MCU_OK;
}

#else  // MCU_HOST_EMU ----------------------------------------------------------------------------------------------------
}  // namespace mcu
// TEST-ONLY host drivers of the value functions above (tests/_emu.py builds this file with -DMCU_HOST_EMU into
// tests/_emu/libmcu_emu.so).  They replace the grid by a loop; nothing here is part of libmauve_cuda.so.
extern "C" {
__attribute__((visibility("default"))) void emu_sol(const uint64_t* keys, const uint32_t* vals, uint64_t npos, uint64_t n, int L, int key_bytes, float* freq_out)
{
    uint32_t* raw = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    if (key_bytes == 4) {
        uint32_t* k32 = (uint32_t*)malloc((npos + 1) * sizeof(uint32_t));
        for (uint64_t i = 0; i < npos; ++i) k32[i] = (uint32_t)keys[i];
        for (uint64_t i = 0; i < npos; ++i) raw[vals[i]] = mcu::sol_run_length<u32>(k32, npos, i);
        free(k32);
    } else {
        for (uint64_t i = 0; i < npos; ++i) raw[vals[i]] = mcu::sol_run_length<u64>(keys, npos, i);
    }
    for (uint64_t k = 0; k < n; ++k) freq_out[k] = mcu::sol_frequency(raw, npos, n, L, k);
    free(raw);
}
__attribute__((visibility("default"))) void emu_anchor_scores(const char* seq0, const char* seq1, const float* freq0, const float* freq1, const mcu_match* rows,
                                                              uint64_t n_rows, const int* matrix, int penalize_repeats, int64_t* match_score_out)
{
    mcu::AnchorScoreArgs a;
    a.seq[0] = (const u8*)seq0;
    a.seq[1] = (const u8*)seq1;
    a.freq[0] = freq0;
    a.freq[1] = freq1;
    a.rows = rows;
    a.n_rows = n_rows;
    for (int i = 0; i < 16; ++i) a.matrix[i] = matrix[i];
    a.penalize_repeats = penalize_repeats;
    a.match_score = match_score_out;
    for (uint64_t k = 0; k < n_rows; ++k) {
        const mcu_match m = rows[k];
        const bool rev0 = m.start0 < 0, rev1 = m.start1 < 0;
        const u64 left0 = (u64)(rev0 ? -m.start0 : m.start0), left1 = (u64)(rev1 ? -m.start1 : m.start1), len = (u64)m.len;
        i64 sum = 0;
        for (u64 c = 0; c < len; ++c) sum += mcu::anchor_col_score(a, a.matrix, left0, left1, len, rev0, rev1, c);
        match_score_out[k] = sum;
    }
}
}
namespace mcu {
#endif

}  // namespace mcu
