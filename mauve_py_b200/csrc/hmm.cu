// Pairwise HomologyHMM on the device (sm_100a).
//
// Replaces run() (LM/HomologyHMM/homologymain.cc:24-62): Forward (homology.cc:307-394), Backward
// (:400-547), posterior(homologous, i) = F_i(H) * B_i(H) / P and the 0.9 threshold (:48-50).
// Model (homology.xml / homology.cc): start -> {homologous, unrelated} with iStartHomologous /
// 1 - iStartHomologous; H -> U iGoUnrelated, U -> H iGoHomologous, H -> end iGoStopFromHomologous,
// U -> end iGoStopFromUnrelated, self transitions take the remainder; both states emit one of the
// eight column symbols '1'..'8' (encoder LM/Islands.h:90-155).
//
// The reference evaluates the recursions in `bfloat` (float mantissa, 2^104-radix exponent,
// algebras.h:41-80).  Here every per-column step is a 2x2 matrix A_x acting on (h, u) in double:
//   forward   f_i     = A_{x_i} f_{i-1}        (f_0 = diag(eH, eU) (start_H, start_U))
//   backward  b_{i-1} = A_{x_i}^T b_i          (b_{L-1} = (stop_H, stop_U))
// and the posterior is scale free:  post_i = f_i(H) b_i(H) / (f_i(H) b_i(H) + f_i(U) b_i(U)),
// so vectors are renormalised freely.  Parallel over columns by chunking: (1) one thread per
// 64-column chunk multiplies its matrices, (2) one thread per string chains the chunk products
// into the vector entering every chunk from the left (forward) and from the right (backward),
// (3) one thread per chunk replays its columns forward, parks f in a coalesced scratch, then walks
// back emitting posterior + H/N.  HBM traffic ~ 1 B/column in, 9 B/column out, 32 B/column scratch.
#include "hmm.cuh"

#include <math.h>

namespace mcu {

constexpr int HC = 64;  // columns per chunk
constexpr int HMM_SPEC = 8;  // columns per speculative group of the exact chain (hmm_exact_chain_warp_kernel)

struct HmmModel {
    double a[8][4];      // per symbol: {HH, UH, HU, UU} transition*emission, f' = (a0 h + a1 u, a2 h + a3 u)
    double first[8][2];  // per symbol: start * emission
    double stop[2];
};

// ---- parameters: getAdaptedHoxdMatrixParameters (LM/HomologyHMM/parameters.h:59-137) and
//      adaptToPercentIdentity (:140-159); same operation order, so the doubles are identical ----
int hmm_params(double gc, double go_homologous, double go_unrelated, double pct_identity, double* out)
{
    const double at = 1 - gc;
    const double gap_u[2] = {0.0483, 0.2535}, gap_h[2] = {0.004461, 0.050733};
    double* eh = out + 5;
    double* eu = out + 13;
    eu[0] = (at / 2) * (at / 2) + (at / 2) * (at / 2);
    eu[1] = (gc / 2) * (gc / 2) + (gc / 2) * (gc / 2);
    eu[2] = (at / 2) * (gc / 2) + (gc / 2) * (at / 2);
    eu[3] = eu[2];
    eu[4] = eu[0];
    eu[5] = eu[1];
    double nf = (1 - (gap_u[0] + gap_u[1])) / (eu[0] + eu[1] + eu[2] + eu[3] + eu[4] + eu[5]);
    for (int i = 0; i < 6; ++i) eu[i] = eu[i] * nf;
    eu[6] = gap_u[0];
    eu[7] = 1 - (eu[0] + eu[1] + eu[2] + eu[3] + eu[4] + eu[5] + eu[6]);
    // HOXD-derived pair frequencies, pre-normalised in the reference
    const double hoxd[6] = {0.1723 * 2, 0.1462 * 2, 0.0180 * 4, 0.0426 * 4, 0.0186 * 2, 0.0142 * 2};
    eh[0] = (at / 0.525) * hoxd[0];
    eh[1] = (gc / 0.475) * hoxd[1];
    eh[2] = hoxd[2];
    eh[3] = hoxd[3];
    eh[4] = (at / 0.525) * hoxd[4];
    eh[5] = (gc / 0.475) * hoxd[5];
    nf = (1 - (gap_h[0] + gap_h[1])) / (eh[0] + eh[1] + eh[2] + eh[3] + eh[4] + eh[5]);
    for (int i = 0; i < 6; ++i) eh[i] = eh[i] * nf;
    eh[6] = gap_h[0];
    eh[7] = 1 - (eh[0] + eh[1] + eh[2] + eh[3] + eh[4] + eh[5] + eh[6]);
    out[0] = 0.5;
    out[1] = 0.00001;
    out[2] = 0.0000001;
    out[3] = 0.0000001;
    out[4] = 0.0000001;
    if (go_homologous > 0) out[1] = go_homologous;  // CLI overrides, MA/progressiveMauve.cpp:236-237
    if (go_unrelated > 0) out[2] = go_unrelated;
    if (pct_identity != 0) {
        if (pct_identity < 0 || pct_identity > 1) { set_error("mcu_hmm_params: bad pct identity %g", pct_identity); return MCU_EINVAL; }
        const double target = pct_identity * (1.0 - eh[6] - eh[7]);
        const double ident = eh[0] + eh[1];
        const double diff = ident - target;
        const double rest = eh[2] + eh[3] + eh[4] + eh[5];
        for (int i = 2; i < 6; ++i) eh[i] += diff * eh[i] / rest;
        eh[0] -= diff * eh[0] / ident;
        eh[1] -= diff * eh[1] / ident;
    }
    return MCU_OK;
}

static void build_model(const double* p, HmmModel* m)
{
    const double tHU = p[2], tUH = p[1], tHE = p[4], tUE = p[3];
    const double tHH = 1.0 - tHU - tHE, tUU = 1.0 - tUH - tUE;
    const double* eh = p + 5;
    const double* eu = p + 13;
    for (int x = 0; x < 8; ++x) {
        m->a[x][0] = eh[x] * tHH;
        m->a[x][1] = eh[x] * tUH;
        m->a[x][2] = eu[x] * tHU;
        m->a[x][3] = eu[x] * tUU;
        m->first[x][0] = p[0] * eh[x];
        m->first[x][1] = (1.0 - p[0]) * eu[x];
    }
    m->stop[0] = tHE;
    m->stop[1] = tUE;
}

// string that owns global chunk c: largest s with chunk_first[s] <= c
__device__ __forceinline__ u32 find_string(const u64* __restrict__ chunk_first, u32 n, u64 c)
{
    u32 lo = 0, hi = n;  // invariant chunk_first[lo] <= c < chunk_first[hi]
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (chunk_first[mid] <= c) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ u32 sym_index(u8 c, u32& bad)
{
    u32 x = (u32)c - (u32)'1';
    bad |= x > 7u;
    return x & 7u;
}

// (1) chunk products P = A_{x_last} ... A_{x_first}, max-normalised
__global__ void __launch_bounds__(128) hmm_products_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, const u64* __restrict__ chunk_first,
                                                          u32 n, u64 nchunks, HmmModel m, double4* __restrict__ prod, u32* __restrict__ err)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const u32 s = find_string(chunk_first, n, c);
    const u64 lc = c - chunk_first[s];
    const u64 beg = off[s] + lc * HC;
    const u64 end = min(off[s + 1], beg + HC);
    u32 bad = 0;
    double p0 = 1, p1 = 0, p2 = 0, p3 = 1;  // row-major [[p0 p1][p2 p3]]
    for (u64 i = beg; i < end; ++i) {
        const u32 x = sym_index(__ldg(sym + i), bad);
        double a0, a1, a2, a3;
        if (lc == 0 && i == beg) { a0 = m.first[x][0]; a1 = 0; a2 = 0; a3 = m.first[x][1]; }  // column 0: diag(e) applied to the start vector
        else { a0 = m.a[x][0]; a1 = m.a[x][1]; a2 = m.a[x][2]; a3 = m.a[x][3]; }
        const double q0 = a0 * p0 + a1 * p2, q1 = a0 * p1 + a1 * p3;
        const double q2 = a2 * p0 + a3 * p2, q3 = a2 * p1 + a3 * p3;
        p0 = q0; p1 = q1; p2 = q2; p3 = q3;
        if (((i - beg) & 15) == 15) {
            const double sc = 1.0 / fmax(fmax(p0, p1), fmax(p2, p3));
            p0 *= sc; p1 *= sc; p2 *= sc; p3 *= sc;
        }
    }
    prod[c] = make_double4(p0, p1, p2, p3);
    if (bad) atomicOr(err, 1u);
}

// (2) per string: vector entering each chunk from the left (fin) and from the right (bin)
__global__ void __launch_bounds__(128) hmm_chain_kernel(const u64* __restrict__ chunk_first, u32 n, HmmModel m, const double4* __restrict__ prod,
                                                       double2* __restrict__ fin, double2* __restrict__ bin)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const u64 c0 = chunk_first[s], c1 = chunk_first[s + 1];
    double h = 1.0, u = 1.0;  // column 0's matrix already carries the start probabilities
    for (u64 c = c0; c < c1; ++c) {
        fin[c] = make_double2(h, u);
        const double4 p = prod[c];
        const double nh = p.x * h + p.y * u, nu = p.z * h + p.w * u;
        const double sc = 1.0 / (nh + nu);
        h = nh * sc; u = nu * sc;
    }
    h = m.stop[0]; u = m.stop[1];
    {
        const double sc = 1.0 / (h + u);
        h *= sc; u *= sc;
    }
    for (u64 c = c1; c-- > c0;) {
        bin[c] = make_double2(h, u);
        // the vector leaving chunk c to the left is P_c^T b -- except that column 0 of the string is not a transition
        const double4 p = prod[c];
        const double nh = p.x * h + p.z * u, nu = p.y * h + p.w * u;
        const double sc = 1.0 / (nh + nu);
        h = nh * sc; u = nu * sc;
    }
}

// (3) per chunk: replay forward, then walk back with the posterior
__global__ void __launch_bounds__(128) hmm_posterior_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, const u64* __restrict__ chunk_first,
                                                           u32 n, u64 nchunks, HmmModel m, const double2* __restrict__ fin,
                                                           const double2* __restrict__ bin, double2* __restrict__ scratch,
                                                           char* __restrict__ pred, double* __restrict__ post)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const u32 s = find_string(chunk_first, n, c);
    const u64 lc = c - chunk_first[s];
    const u64 beg = off[s] + lc * HC;
    const u64 end = min(off[s + 1], beg + HC);
    double2* sc = scratch + (u64)blockIdx.x * HC * blockDim.x + threadIdx.x;  // [column in chunk][thread]: coalesced
    u32 bad = 0;
    double2 f = fin[c];
    double h = f.x, u = f.y;
    for (u64 i = beg; i < end; ++i) {
        const u32 x = sym_index(__ldg(sym + i), bad);
        double nh, nu;
        if (lc == 0 && i == beg) { nh = m.first[x][0] * h; nu = m.first[x][1] * u; }
        else { nh = m.a[x][0] * h + m.a[x][1] * u; nu = m.a[x][2] * h + m.a[x][3] * u; }
        h = nh; u = nu;
        if (((i - beg) & 15) == 15) {
            const double r = 1.0 / (h + u);
            h *= r; u *= r;
        }
        sc[(i - beg) * blockDim.x] = make_double2(h, u);
    }
    double2 b = bin[c];
    double bh = b.x, bu = b.y;
    for (u64 i = end; i-- > beg;) {
        const double2 fv = sc[(i - beg) * blockDim.x];
        const double ph = fv.x * bh, pu = fv.y * bu;
        const double po = ph / (ph + pu);
        if (post) post[i] = po;
        pred[i] = po >= 0.9 ? 'H' : 'N';
        const u32 x = sym_index(__ldg(sym + i), bad);
        const double nh = m.a[x][0] * bh + m.a[x][2] * bu, nu = m.a[x][1] * bh + m.a[x][3] * bu;
        bh = nh; bu = nu;
        if (((i - beg) & 15) == 0) {
            const double r = 1.0 / (bh + bu);
            bh *= r; bu *= r;
        }
    }
}


// =========================================================================================
// bfloat-faithful evaluation (the default).
//
// The reference's numbers are `bfloat`s: a float32 mantissa kept in [1e-18, 1e18] and an int exponent of radix 2^104
// (algebras.h:41-80).  Its float32 rounding noise accumulates along the recursions, so against an evaluation in double
// the posteriors drift apart with the string length (measured on B200: 2e-6 relative at 10 k columns, 1e-5 at 30 k,
// 6.5e-5 at 3 M, with H/N flips at the 0.9 threshold), i.e. beyond the 1e-5 parity bar for LCB-sized strings.  These
// kernels therefore execute the reference's own operation sequence with IEEE round-to-nearest intrinsics (no FMA
// contraction, as the reference's x86-64 build has none):
//   Forward  (homology.cc:307-394):  F_U(k) = (T4*eU)*F_U(k-1) (+)= (T3*eU)*F_H(k-1);  F_H(k) = (T5*eH)*F_U(k-1) (+)= (T2*eH)*F_H(k-1)
//            with double*bfloat = bfloat_pr_double_product (algebras.h:225-231) and (+)= = bfloat_pr_sum_accum (:263-277);
//            P = (T7*1.0)*F_U(L) (+)= (T6*1.0)*F_H(L)
//   Backward (:400-547):             B_H(k) = (T2*eH')*B_H(k+1) (+)= (T3*eU')*B_U(k+1);  B_U(k) = (T4*eU')*B_U(k+1) (+)= (T5*eH')*B_H(k+1)
//            (eX' = emission of column k+1), B_H(L) = T6*1.0 * 1, B_U(L) = T7*1.0 * 1
//   posterior (homologymain.cc:48):  double(F_H(k) * B_H(k) / P)  with bfloat_pr_product / bfloat_pr_quotient and BFloat::Value
// The recursions are serial in k by construction (float rounding is order dependent): one thread per (string, direction),
// parallel over the strings of the batch; the posterior pass is parallel over columns.
// =========================================================================================
struct BF {
    float f;
    int e;
};
constexpr int BF_INF = 1000000000;  // cBFloatInfinity

struct HmmExactModel {
    double te[8][4];   // per symbol: T4*eU, T3*eU, T5*eH, T2*eH   (the doubles the generated code forms per column)
    double first[8][2];// per symbol: T1*eU, T0*eH
    double stop[2];    // T7*1.0, T6*1.0
    double range_sqrt, range_inv_sqrt;  // (double)(float)1e18, (double)(float)1e-18
    double value_tbl[50];               // BFloat::aDoubleConversionLookup (algebras.cc)
    double log_range;                   // (double)logcBFloatRange
};

__device__ __forceinline__ BF bf_dprod(BF a, double b, const HmmExactModel& m)  // bfloat_pr_double_product
{
    double x = __dmul_rn((double)a.f, b);
    int e = a.e;
    if (x <= 0.0) return BF{0.f, -BF_INF};
    while (x > m.range_sqrt) { x = __dmul_rn(x, 4.930380657631324e-32); ++e; }      // * 2^-104
    while (x < m.range_inv_sqrt) { x = __dmul_rn(x, 2.028240960365167e+31); --e; }   // * 2^104
    return BF{__double2float_rn(x), e};
}

__device__ __forceinline__ float bf_conv(int k) { return k == 0 ? 1.0f : (k == 1 ? 4.930380657631324e-32f : 0.0f); }  // aConversionLookup: 2^-104k in float

__device__ __forceinline__ void bf_sum_accum(BF& a, BF b)  // bfloat_pr_sum_accum
{
    if (a.e >= b.e) {
        if (a.e < b.e + 100) a.f = __fadd_rn(a.f, __fmul_rn(b.f, bf_conv(a.e - b.e)));
    } else if (a.e > b.e - 100) {
        a.f = __fadd_rn(b.f, __fmul_rn(a.f, bf_conv(b.e - a.e)));
        a.e = b.e;
    } else
        a = b;
}

__device__ __forceinline__ void bf_normalise(BF& a)  // BFloatNormalise
{
    if (a.f > 1.0e+18f) { a.f = __fmul_rn(a.f, 4.930380657631324e-32f); ++a.e; }
    else if (a.f < 1.0e-18f) {
        if (a.f == 0.0f) a.e = -BF_INF;
        else { a.f = __fmul_rn(a.f, 2.028240960365167e+31f); --a.e; }
    }
}

__device__ __forceinline__ double bf_value(BF a, const HmmExactModel& m)  // BFloat::Value
{
    const int ae = a.e < 0 ? -a.e : a.e;
    if (ae < 25) return __dmul_rn((double)a.f, m.value_tbl[a.e + 25]);
    if (a.e < 25) return 0.0;
    return (double)a.f * exp((double)a.e * m.log_range);
}

// thread t < n: forward chain of string t; thread n + t: its backward chain (so that the lanes of a warp walk the same way).
// fh / bh: the homologous-state value of every column.
__global__ void __launch_bounds__(64) hmm_exact_chain_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, u32 n, HmmExactModel m,
                                                            BF* __restrict__ fh, BF* __restrict__ bh, BF* __restrict__ total, u32* __restrict__ err)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2ull * n) return;
    const bool fwd = t < n;
    const u32 s = (u32)(fwd ? t : t - n);
    const u64 beg = off[s], end = off[s + 1];
    if (end == beg) return;
    u32 bad = 0;
    const BF one = BF{1.0f, 0};  // double2bfloat(1.0)
    if (fwd) {
        u32 x = sym_index(__ldg(sym + beg), bad);
        BF u = bf_dprod(one, m.first[x][0], m), h = bf_dprod(one, m.first[x][1], m);
        fh[beg] = h;
        for (u64 i = beg + 1; i < end; ++i) {
            x = sym_index(__ldg(sym + i), bad);
            BF nu = bf_dprod(u, m.te[x][0], m);
            bf_sum_accum(nu, bf_dprod(h, m.te[x][1], m));
            BF nh = bf_dprod(u, m.te[x][2], m);
            bf_sum_accum(nh, bf_dprod(h, m.te[x][3], m));
            u = nu;
            h = nh;
            fh[i] = h;
        }
        BF p = bf_dprod(u, m.stop[0], m);
        bf_sum_accum(p, bf_dprod(h, m.stop[1], m));
        total[s] = p;
    } else {
        BF h = bf_dprod(one, m.stop[1], m), u = bf_dprod(one, m.stop[0], m);
        bh[end - 1] = h;
        for (u64 i = end - 1; i > beg; --i) {  // value of column i - 1 from the emission of column i
            const u32 x = sym_index(__ldg(sym + i), bad);
            BF nh = bf_dprod(h, m.te[x][3], m);
            bf_sum_accum(nh, bf_dprod(u, m.te[x][1], m));
            BF nu = bf_dprod(u, m.te[x][0], m);
            bf_sum_accum(nu, bf_dprod(h, m.te[x][2], m));
            h = nh;
            u = nu;
            bh[i - 1] = h;
        }
    }
    if (bad) atomicOr(err, 1u);
}

// Few, long strings: one warp per (string, direction).  Lane 0 runs the serial recurrence; the other lanes keep memory
// latency out of its dependency chain: per block of 32 columns they fetch the symbols (one coalesced load, issued one
// block ahead), look up the four coefficients of every column into shared memory, and afterwards store the 32 results
// with one coalesced write.
__device__ __forceinline__ BF bf_dprod_fast(float af, int ae, double b, double hi, double lo)
{
    double x = __dmul_rn((double)af, b);
    if (x <= 0.0) return BF{0.f, -BF_INF};
    if (x > hi || x < lo) {  // rare: the mantissa left [1e-18, 1e18]
        while (x > hi) { x = __dmul_rn(x, 4.930380657631324e-32); ++ae; }
        while (x < lo) { x = __dmul_rn(x, 2.028240960365167e+31); --ae; }
    }
    return BF{__double2float_rn(x), ae};
}

__device__ __forceinline__ void bf_sum_accum_fast(BF& a, BF b)
{
    if (a.e == b.e) a.f = __fadd_rn(a.f, b.f);  // b.f * aConversionLookup[0] = b.f * 1.0f exactly
    else bf_sum_accum(a, b);
}

__global__ void __launch_bounds__(32) hmm_exact_chain_warp_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, u32 n, HmmExactModel m,
                                                                 BF* __restrict__ fh, BF* __restrict__ bh, BF* __restrict__ total, u32* __restrict__ err)
{
    __shared__ double te_s[8][4];
    __shared__ double4 coef[32];
    __shared__ BF res[32];
    const u32 lane = threadIdx.x;
    const u32 s = blockIdx.x >> 1;
    const bool fwd = (blockIdx.x & 1) == 0;
    const u64 beg = off[s], end = off[s + 1];
    if (end == beg) return;
    te_s[lane >> 2][lane & 3] = m.te[lane >> 2][lane & 3];
    __syncwarp();
    const double hi = m.range_sqrt, lo = m.range_inv_sqrt;
    const u64 len = end - beg;
    u32 bad = 0;
    const BF one = BF{1.0f, 0};
    BF h, u;
    // columns are visited in chain order: forward k = 1 .. len-1 uses the symbol of column k; backward step k (value of column
    // len-1-k) uses the symbol of column len-k
    if (fwd) {
        const u32 x = sym_index(__ldg(sym + beg), bad);
        u = bf_dprod(one, m.first[x][0], m);
        h = bf_dprod(one, m.first[x][1], m);
        if (lane == 0) fh[beg] = h;
    } else {
        h = bf_dprod(one, m.stop[1], m);
        u = bf_dprod(one, m.stop[0], m);
        if (lane == 0) bh[end - 1] = h;
    }
    const u64 steps = len - 1;
    auto sym_of_step = [&](u64 k) -> u64 { return fwd ? beg + 1 + k : end - 1 - k; };  // index of the symbol step k consumes
    u32 xn = 0;
    if (lane < steps) xn = sym_index(__ldg(sym + sym_of_step(lane)), bad);
    for (u64 k0 = 0; k0 < steps; k0 += 32) {
        const u32 cnt = (u32)min((u64)32, steps - k0);
        const u32 x = xn;
        if (k0 + 32 + lane < steps) xn = sym_index(__ldg(sym + sym_of_step(k0 + 32 + lane)), bad);  // next block, in flight during this one
        coef[lane] = make_double4(te_s[x][0], te_s[x][1], te_s[x][2], te_s[x][3]);
        __syncwarp();
        if (lane == 0) {
            // Common case, decided with integer tests on the high words of the four products: all of them positive and well
            // inside [1e-18, 1e18] (no renormalisation) and both states on the same exponent (aConversionLookup[0] = 1): the step
            // is 2 float->double conversions, 4 double products, 4 double->float roundings and 2 float additions.  Columns are
            // taken in groups of HMM_SPEC: the group is evaluated in that form without a branch per column (the range tests only
            // accumulate a flag, so the dependency chain of a column is conversion -> product -> rounding -> addition and nothing
            // else), and a group in which any column left the common case is evaluated again, operation by operation, from the
            // state it started with.
            const u32 w_lo = (u32)__double2hiint(lo) + 1u, w_span = (u32)__double2hiint(hi) - w_lo;
            auto exact_step = [&](u32 j) {
                const double4 c = coef[j];
                const double du = (double)u.f, dh = (double)h.f;
                // forward: U <- (c.x u) + (c.y h), H <- (c.z u) + (c.w h); backward: H <- (c.w h) + (c.y u), U <- (c.x u) + (c.z h)
                const double p0 = __dmul_rn(fwd ? du : dh, fwd ? c.x : c.w), p1 = __dmul_rn(fwd ? dh : du, c.y);
                const double p2 = __dmul_rn(du, fwd ? c.z : c.x), p3 = __dmul_rn(dh, fwd ? c.w : c.z);
                const bool in_range = ((u32)__double2hiint(p0) - w_lo < w_span) & ((u32)__double2hiint(p1) - w_lo < w_span) &
                                      ((u32)__double2hiint(p2) - w_lo < w_span) & ((u32)__double2hiint(p3) - w_lo < w_span);
                if (in_range && u.e == h.e) {
                    const float a = __fadd_rn(__double2float_rn(p0), __double2float_rn(p1));
                    const float b = __fadd_rn(__double2float_rn(p2), __double2float_rn(p3));
                    if (fwd) { u.f = a; h.f = b; } else { h.f = a; u.f = b; }
                } else if (fwd) {
                    BF nu = bf_dprod_fast(u.f, u.e, c.x, hi, lo);
                    bf_sum_accum_fast(nu, bf_dprod_fast(h.f, h.e, c.y, hi, lo));
                    BF nh = bf_dprod_fast(u.f, u.e, c.z, hi, lo);
                    bf_sum_accum_fast(nh, bf_dprod_fast(h.f, h.e, c.w, hi, lo));
                    u = nu;
                    h = nh;
                } else {
                    BF nh = bf_dprod_fast(h.f, h.e, c.w, hi, lo);
                    bf_sum_accum_fast(nh, bf_dprod_fast(u.f, u.e, c.y, hi, lo));
                    BF nu = bf_dprod_fast(u.f, u.e, c.x, hi, lo);
                    bf_sum_accum_fast(nu, bf_dprod_fast(h.f, h.e, c.z, hi, lo));
                    h = nh;
                    u = nu;
                }
                res[j] = h;
            };
            u32 j0 = 0;
            for (; j0 + HMM_SPEC <= cnt; j0 += HMM_SPEC) {
                bool ok = u.e == h.e;
                if (ok) {
                    float uf = u.f, hf = h.f;
                    u32 good = 1u;
#pragma unroll
                    for (int q = 0; q < HMM_SPEC; ++q) {
                        const double4 c = coef[j0 + q];
                        const double du = (double)uf, dh = (double)hf;
                        const double p0 = __dmul_rn(fwd ? du : dh, fwd ? c.x : c.w), p1 = __dmul_rn(fwd ? dh : du, c.y);
                        const double p2 = __dmul_rn(du, fwd ? c.z : c.x), p3 = __dmul_rn(dh, fwd ? c.w : c.z);
                        good &= ((u32)__double2hiint(p0) - w_lo < w_span) & ((u32)__double2hiint(p1) - w_lo < w_span) &
                                ((u32)__double2hiint(p2) - w_lo < w_span) & ((u32)__double2hiint(p3) - w_lo < w_span);
                        const float a = __fadd_rn(__double2float_rn(p0), __double2float_rn(p1));
                        const float b = __fadd_rn(__double2float_rn(p2), __double2float_rn(p3));
                        if (fwd) { uf = a; hf = b; } else { hf = a; uf = b; }
                        res[j0 + q] = BF{hf, h.e};
                    }
                    ok = good != 0u;
                    if (ok) { u.f = uf; h.f = hf; }
                }
                if (!ok)
                    for (u32 q = 0; q < HMM_SPEC; ++q) exact_step(j0 + q);
            }
            for (; j0 < cnt; ++j0) exact_step(j0);
        }
        __syncwarp();
        if (lane < cnt) {
            if (fwd) fh[beg + 1 + k0 + lane] = res[lane];
            else bh[end - 2 - k0 - lane] = res[lane];
        }
        __syncwarp();
    }
    if (fwd && lane == 0) {
        BF p = bf_dprod(u, m.stop[0], m);
        bf_sum_accum(p, bf_dprod(h, m.stop[1], m));
        total[s] = p;
    }
    if (bad) atomicOr(err, 1u);
}

__global__ void __launch_bounds__(256) hmm_exact_posterior_kernel(const u64* __restrict__ off, u32 n, u64 total_cols, HmmExactModel m,
                                                                 const BF* __restrict__ fh, const BF* __restrict__ bh, const BF* __restrict__ total,
                                                                 char* __restrict__ pred, double* __restrict__ post)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_cols) return;
    u32 lo = 0, hi = n;  // off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    const BF a = fh[i], b = bh[i], p = total[lo];
    BF q = BF{__fmul_rn(a.f, b.f), a.e + b.e};  // bfloat_pr_product
    bf_normalise(q);
    BF r = BF{__fdiv_rn(q.f, p.f), q.e - p.e};  // bfloat_pr_quotient
    bf_normalise(r);
    const double po = bf_value(r, m);
    if (post) post[i] = po;
    pred[i] = po >= 0.9 ? 'H' : 'N';
}

static void build_exact_model(const double* p, HmmExactModel* m)
{
    // iTransition[] of homology.cc:322-337
    const double T0 = p[0], T1 = 1.0 - p[0], T2 = 1.0 - p[2] - p[4], T3 = p[2], T4 = 1.0 - p[1] - p[3], T5 = p[1], T6 = p[4], T7 = p[3];
    const double* eh = p + 5;
    const double* eu = p + 13;
    for (int x = 0; x < 8; ++x) {
        m->te[x][0] = T4 * eu[x];
        m->te[x][1] = T3 * eu[x];
        m->te[x][2] = T5 * eh[x];
        m->te[x][3] = T2 * eh[x];
        m->first[x][0] = T1 * eu[x];
        m->first[x][1] = T0 * eh[x];
    }
    m->stop[0] = T7 * 1.0;
    m->stop[1] = T6 * 1.0;
    m->range_sqrt = (double)(float)1.0e+18;
    m->range_inv_sqrt = (double)(float)1.0e-18;
    const float range = 20282409603651670423947251286016.0f;  // cBFloatRange = 2^104
    const float log_range = logf(range);                      // logcBFloatRange is a BFMantissa (float)
    m->log_range = (double)log_range;
    // algebras.cc fills the table with exp((i - 25) * logcBFloatRange): int * float is a float product and std::exp(float) is expf
    for (int i = 0; i < 50; ++i) m->value_tbl[i] = (double)expf((float)(i - 25) * log_range);
}

struct HmmState {
    DevBuf sym, off, chunk_first, prod, fin, bin, scratch, pred, post, err, fh, bh, total;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};
static HmmState g_hmm;

int hmm_batch(u64 n, const char* sym, const u64* off, const double* params, char* pred_out, double* post_out, float* device_ms)
{
    HmmState& st = g_hmm;
    if (device_ms) *device_ms = 0.f;
    if (n == 0) return MCU_OK;
    if (!sym || !off || !params || !pred_out) { set_error("mcu_hmm_batch: NULL pointer"); return MCU_EINVAL; }
    if (n >= 0x7FFFFFFFull) { set_error("mcu_hmm_batch: too many strings"); return MCU_EINVAL; }
    for (int i = 0; i < 21; ++i)
        if (!(params[i] >= 0.0 && params[i] <= 1.0)) { set_error("mcu_hmm_batch: parameter %d is not a probability", i); return MCU_EINVAL; }
    const u64 total = off[n];
    u64* cf = (u64*)malloc((n + 1) * sizeof(u64));
    if (!cf) { set_error("out of host memory"); return MCU_ENOMEM; }
    u64 nchunks = 0;
    for (u64 s = 0; s < n; ++s) {
        if (off[s + 1] < off[s]) { free(cf); set_error("mcu_hmm_batch: offsets not monotone"); return MCU_EINVAL; }
        cf[s] = nchunks;
        nchunks += div_up(off[s + 1] - off[s], HC);
    }
    cf[n] = nchunks;
    if (total == 0) { free(cf); return MCU_OK; }
    if (!st.stream) {
        MCU_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        MCU_CUDA(cudaEventCreate(&st.e0));
        MCU_CUDA(cudaEventCreate(&st.e1));
    }
    cudaStream_t s = st.stream;
    HmmModel m;
    build_model(params, &m);
    const int block = 128;
    const u64 grid_c = div_up(nchunks, block);
    // MAUVE_CUDA_HMM_SCAN=1: the column-parallel evaluation in double (fast on a single long string, but only within
    // ~1e-5 of the reference for strings up to ~10 k columns); default: the bfloat-faithful chains
    const bool scan_mode = getenv("MAUVE_CUDA_HMM_SCAN") != nullptr;
    int r = MCU_OK;
    if ((r = st.sym.reserve(total + 16)) || (r = st.off.reserve((n + 1) * 8)) || (r = st.chunk_first.reserve((n + 1) * 8)) ||
        (r = st.prod.reserve(nchunks * sizeof(double4))) || (r = st.fin.reserve(nchunks * sizeof(double2))) ||
        (r = st.bin.reserve(nchunks * sizeof(double2))) || (r = st.scratch.reserve(scan_mode ? grid_c * block * HC * sizeof(double2) : 16)) ||
        (r = st.pred.reserve(total + 16)) || (r = st.post.reserve(total * 8 + 16)) || (r = st.err.reserve(16)) ||
        (!scan_mode && ((r = st.fh.reserve(total * 8 + 16)) || (r = st.bh.reserve(total * 8 + 16)) || (r = st.total.reserve(n * 8 + 16))))) {
        free(cf);
        return r;
    }
    cudaError_t e = cudaMemcpyAsync(st.chunk_first.p, cf, (n + 1) * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    free(cf);
    MCU_CUDA(e);
    MCU_CUDA(cudaMemcpyAsync(st.sym.p, sym, total, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemcpyAsync(st.off.p, off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    MCU_CUDA(cudaMemsetAsync(st.err.p, 0, 16, s));
    MCU_CUDA(cudaEventRecord(st.e0, s));
    if (scan_mode) {
        hmm_products_kernel<<<(unsigned)grid_c, block, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), st.chunk_first.as<u64>(), (u32)n, nchunks, m,
                                                               st.prod.as<double4>(), st.err.as<u32>());
        hmm_chain_kernel<<<(unsigned)div_up(n, block), block, 0, s>>>(st.chunk_first.as<u64>(), (u32)n, m, st.prod.as<double4>(), st.fin.as<double2>(),
                                                                      st.bin.as<double2>());
        hmm_posterior_kernel<<<(unsigned)grid_c, block, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), st.chunk_first.as<u64>(), (u32)n, nchunks, m,
                                                                st.fin.as<double2>(), st.bin.as<double2>(), st.scratch.as<double2>(),
                                                                st.pred.as<char>(), post_out ? st.post.as<double>() : nullptr);
    } else {
        HmmExactModel xm;
        build_exact_model(params, &xm);
        // few chains: a warp each (latency-optimised); many chains: a thread each (throughput)
        if (2 * n <= (u64)sm_count() * 64)
            hmm_exact_chain_warp_kernel<<<(unsigned)(2 * n), 32, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), (u32)n, xm, st.fh.as<BF>(), st.bh.as<BF>(),
                                                                         st.total.as<BF>(), st.err.as<u32>());
        else
            hmm_exact_chain_kernel<<<(unsigned)div_up(2 * n, 64), 64, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), (u32)n, xm, st.fh.as<BF>(), st.bh.as<BF>(),
                                                                              st.total.as<BF>(), st.err.as<u32>());
        hmm_exact_posterior_kernel<<<(unsigned)div_up(total, 256), 256, 0, s>>>(st.off.as<u64>(), (u32)n, total, xm, st.fh.as<BF>(), st.bh.as<BF>(),
                                                                                st.total.as<BF>(), st.pred.as<char>(),
                                                                                post_out ? st.post.as<double>() : nullptr);
    }
    MCU_CUDA(cudaEventRecord(st.e1, s));
    MCU_CUDA(cudaGetLastError());
    u32 err_flag = 0;
    MCU_CUDA(cudaMemcpyAsync(&err_flag, st.err.p, 4, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(pred_out, st.pred.p, total, cudaMemcpyDeviceToHost, s));
    if (post_out) MCU_CUDA(cudaMemcpyAsync(post_out, st.post.p, total * 8, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    if (err_flag) { set_error("mcu_hmm_batch: symbol outside '1'..'8'"); return MCU_EINVAL; }
    if (device_ms) cudaEventElapsedTime(device_ms, st.e0, st.e1);
    return MCU_OK;
}

}  // namespace mcu
