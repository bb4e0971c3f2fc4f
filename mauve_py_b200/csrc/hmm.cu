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
#ifndef MCU_HOST_EMU
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

#endif

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

// Everything that decides a VALUE of the bfloat-faithful evaluation is __host__ __device__ with explicitly rounded operations, so
// that tests/ can run exactly this code on the CPU (tests/_emu.py builds this file with -DMCU_HOST_EMU into a test-only
// library whose drivers are at the end of the file; the product library is built without the macro and has no host path).
#define HMM_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define H_FMUL(a, b) __fmul_rn(a, b)
#define H_FADD(a, b) __fadd_rn(a, b)
#define H_FMA(a, b, c) __fmaf_rn(a, b, c)
#define H_FDIV(a, b) __fdiv_rn(a, b)
#define H_DMUL(a, b) __dmul_rn(a, b)
#define H_D2F(x) __double2float_rn(x)
#define H_F2U(x) __float_as_uint(x)
#define H_U2F(x) __uint_as_float(x)
#else   // host (test-only build, compiled with -ffp-contract=off: every operator below is one IEEE operation)
static inline float h_fma(float a, float b, float c) { return fmaf(a, b, c); }
static inline u32 h_f2u(float x) { u32 r; memcpy(&r, &x, 4); return r; }
static inline float h_u2f(u32 x) { float r; memcpy(&r, &x, 4); return r; }
#define H_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define H_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define H_FMA(a, b, c) h_fma(a, b, c)
#define H_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#define H_DMUL(a, b) ((double)((double)(a) * (double)(b)))
#define H_D2F(x) ((float)(x))
#define H_F2U(x) h_f2u(x)
#define H_U2F(x) h_u2f(x)
#endif

struct HmmExactModel {
    double te[8][4];   // per symbol: T4*eU, T3*eU, T5*eH, T2*eH   (the doubles the generated code forms per column)
    double first[8][2];// per symbol: T1*eU, T0*eH
    double stop[2];    // T7*1.0, T6*1.0
    double range_sqrt, range_inv_sqrt;  // (double)(float)1e18, (double)(float)1e-18
    double value_tbl[50];               // BFloat::aDoubleConversionLookup (algebras.cc)
    double log_range;                   // (double)logcBFloatRange
};

HMM_HD BF bf_dprod(BF a, double b, const HmmExactModel& m)  // bfloat_pr_double_product
{
    double x = H_DMUL((double)a.f, b);
    int e = a.e;
    if (x <= 0.0) return BF{0.f, -BF_INF};
    while (x > m.range_sqrt) { x = H_DMUL(x, 4.930380657631324e-32); ++e; }      // * 2^-104
    while (x < m.range_inv_sqrt) { x = H_DMUL(x, 2.028240960365167e+31); --e; }   // * 2^104
    return BF{H_D2F(x), e};
}

HMM_HD float bf_conv(int k) { return k == 0 ? 1.0f : (k == 1 ? 4.930380657631324e-32f : 0.0f); }  // aConversionLookup: 2^-104k in float

HMM_HD void bf_sum_accum(BF& a, BF b)  // bfloat_pr_sum_accum
{
    if (a.e >= b.e) {
        if (a.e < b.e + 100) a.f = H_FADD(a.f, H_FMUL(b.f, bf_conv(a.e - b.e)));
    } else if (a.e > b.e - 100) {
        a.f = H_FADD(b.f, H_FMUL(a.f, bf_conv(b.e - a.e)));
        a.e = b.e;
    } else
        a = b;
}

HMM_HD void bf_normalise(BF& a)  // BFloatNormalise
{
    if (a.f > 1.0e+18f) { a.f = H_FMUL(a.f, 4.930380657631324e-32f); ++a.e; }
    else if (a.f < 1.0e-18f) {
        if (a.f == 0.0f) a.e = -BF_INF;
        else { a.f = H_FMUL(a.f, 2.028240960365167e+31f); --a.e; }
    }
}

HMM_HD double bf_value(BF a, const HmmExactModel& m)  // BFloat::Value
{
    const int ae = a.e < 0 ? -a.e : a.e;
    if (ae < 25) return H_DMUL((double)a.f, m.value_tbl[a.e + 25]);
    if (a.e < 25) return 0.0;
    return (double)a.f * exp((double)a.e * m.log_range);
}

// One column of either recursion, operation by operation as the reference performs it.  Coefficients in ROLE order:
// U' = (c0 u) (+)= (c1 h), H' = (c2 u) (+)= (c3 h) for Forward; Backward forms H' = (c3 h) (+)= (c2 u) first and then
// U' = (c0 u) (+)= (c1 h) with c1 = T5*eH, c2 = T3*eU (see hmm_role_coef): the accumulation ORDER differs between the two and is kept.
HMM_HD void hmm_exact_step(BF& u, BF& h, const double c[4], bool fwd, const HmmExactModel& m)
{
    if (fwd) {
        BF nu = bf_dprod(u, c[0], m);
        bf_sum_accum(nu, bf_dprod(h, c[1], m));
        BF nh = bf_dprod(u, c[2], m);
        bf_sum_accum(nh, bf_dprod(h, c[3], m));
        u = nu;
        h = nh;
    } else {
        BF nh = bf_dprod(h, c[3], m);
        bf_sum_accum(nh, bf_dprod(u, c[2], m));
        BF nu = bf_dprod(u, c[0], m);
        bf_sum_accum(nu, bf_dprod(h, c[1], m));
        h = nh;
        u = nu;
    }
}

// role-ordered coefficient k of symbol x: Forward (te0, te1, te2, te3); Backward (te0, te2, te1, te3)
HMM_HD double hmm_role_coef(const HmmExactModel& m, u32 x, int k, bool fwd) { return m.te[x][(fwd || k == 0 || k == 3) ? k : 3 - k]; }

// ---- the recurrence without the FP64 pipe ---------------------------------------------------------------------------------------
// On this part double-precision instructions (DMUL, F2F.F64) issue at about two lanes per clock and SM: the four
// bfloat_pr_double_products of a column, y = (float)((double)v * c), cost ~550 cycles of a one-lane chain.
//
// (1) The product.  Split c = ch + cl + (< 2^-48 |c|) with ch = (float)c, cl = (float)(c - ch).  Then
//         tl = RN32(v * cl),   r = RN32(v * ch + tl)          (one FMUL, one FFMA)
//     rounds a number S with |S - v c| < 2^-46 |r| to float, and the reference rounds X = RN64(v c), |X - v c| <= 2^-53 |r|.  RN32 is
//     monotone, so r != y needs a float rounding boundary (the midpoint of two neighbouring floats) within 2^-46 |r| <= 2^-22 ulp(r)
//     of the exact product.  dd = RN32(RN32(v * ch - r) + tl) measures (v c - r) to ~2^-22 ulp; the product is HAZARDOUS when |dd|
//     comes within 2^-15 ulp of ulp/2 or when r is a power of two (the boundary below it sits at ulp/4).
// (2) The renormalisation.  The reference multiplies X by 2^104 while X < 1e-18 (exponent - 1) and by 2^-104 while X > 1e18; powers of
//     two commute with both roundings, so the renormalised mantissa is r * 2^104 when r < 1e-18f.  Hazardous: r within 16 ulp of
//     1e-18f (X and r could sit on different sides), r near or above 1e18f, r below 1e-37 (a second renormalisation, or zero).
// (3) The sum.  a (+)= b adds the mantissas after multiplying the one with the smaller exponent by aConversionLookup[difference]
//     = 2^-104 (difference 1) or 0 (the table underflows from difference 2 on), and keeps the larger exponent; it is symmetric in a
//     and b.  With E the larger exponent, product i of a state with exponent e contributes r_i * M(e - E), M(k) = 2^(104 k) for
//     k in {1, 0, -1} and 0 below (when r_i was renormalised the two scalings collapse into one exact multiplication).
// hmm_float_step is (1)-(3) for one column; a hazardous column (about one in 2,000) is evaluated with hmm_exact_step instead.
struct HmmFastTab {
    float hi[8][4], lo[8][4];  // te = hi + lo, per symbol, in te order
};

HMM_HD float hmm_fprod(float v, float ch, float cl) { return H_FMA(v, ch, H_FMUL(v, cl)); }

// the rounding hazard of (1)
HMM_HD u32 hmm_fprod_hazard(float v, float ch, float cl, float r)
{
    const float tl = H_FMUL(v, cl);
    const float dd = H_FADD(H_FMA(v, ch, -r), tl);
    const u32 rb = H_F2U(r);
    const float half_ulp = H_U2F((rb & 0x7f800000u) - (24u << 23));
    const float dist = fabsf(fabsf(dd) - half_ulp);
    return (u32)(dist < half_ulp * 6.103515625e-05f) | (u32)((rb & 0x007fffffu) == 0u);
}

constexpr u32 BF_LO_BITS = 0x219392efu;  // 1e-18f
constexpr u32 BF_HI_BITS = 0x5d5e0b6bu;  // 1e+18f

// hazards of (1) and (2); low = the reference renormalises this product upwards (exponent - 1)
HMM_HD u32 hmm_prod_check(float v, float ch, float cl, float r, u32& low)
{
    const u32 rb = H_F2U(r);
    low = (u32)(rb < BF_LO_BITS);                       // positive floats order like their bit patterns
    return hmm_fprod_hazard(v, ch, cl, r) | (u32)(rb - (BF_LO_BITS - 16u) <= 32u) | (u32)(rb >= BF_HI_BITS - 16u) | (u32)!(r > 1.0e-37f);
}

HMM_HD float hmm_mexp(int k) { return k == 0 ? 1.0f : (k == 1 ? 2.028240960365167e+31f : (k == -1 ? 4.930380657631324e-32f : 0.0f)); }

// posterior(homologous, column) = double(F_H * B_H / P), homologymain.cc:48
HMM_HD double hmm_posterior_value(BF a, BF b, BF p, const HmmExactModel& m)
{
    BF q = BF{H_FMUL(a.f, b.f), a.e + b.e};  // bfloat_pr_product
    bf_normalise(q);
    BF r = BF{H_FDIV(q.f, p.f), q.e - p.e};  // bfloat_pr_quotient
    bf_normalise(r);
    return bf_value(r, m);
}

// One column in FP32, any exponents.  c / l: high and low parts of the coefficients in role order (U' = c0 u + c1 h, H' = c2 u + c3 h;
// the sum is symmetric, so Backward is Forward with c1 and c2 exchanged).  CHECK: returns false, leaving u and h alone, on a hazard.
template <bool CHECK>
HMM_HD bool hmm_float_step_t(BF& u, BF& h, const float c[4], const float l[4])
{
    const float r0 = hmm_fprod(u.f, c[0], l[0]), r1 = hmm_fprod(h.f, c[1], l[1]);
    const float r2 = hmm_fprod(u.f, c[2], l[2]), r3 = hmm_fprod(h.f, c[3], l[3]);
    u32 w0, w1, w2, w3;
    if (CHECK) {
        const u32 hz = hmm_prod_check(u.f, c[0], l[0], r0, w0) | hmm_prod_check(h.f, c[1], l[1], r1, w1) |
                       hmm_prod_check(u.f, c[2], l[2], r2, w2) | hmm_prod_check(h.f, c[3], l[3], r3, w3);
        if (hz) return false;
    } else {
        w0 = (u32)(H_F2U(r0) < BF_LO_BITS); w1 = (u32)(H_F2U(r1) < BF_LO_BITS);
        w2 = (u32)(H_F2U(r2) < BF_LO_BITS); w3 = (u32)(H_F2U(r3) < BF_LO_BITS);
    }
    const int e0 = u.e - (int)w0, e1 = h.e - (int)w1, e2 = u.e - (int)w2, e3 = h.e - (int)w3;
    const int eu = e0 > e1 ? e0 : e1, eh = e2 > e3 ? e2 : e3;
    const float nu = H_FADD(H_FMUL(r0, hmm_mexp(u.e - eu)), H_FMUL(r1, hmm_mexp(h.e - eu)));
    const float nh = H_FADD(H_FMUL(r2, hmm_mexp(u.e - eh)), H_FMUL(r3, hmm_mexp(h.e - eh)));
    u = BF{nu, eu};
    h = BF{nh, eh};
    return true;
}
HMM_HD bool hmm_float_step(BF& u, BF& h, const float c[4], const float l[4]) { return hmm_float_step_t<true>(u, h, c, l); }

// ---- the chain in a VIRTUAL mantissa domain ---------------------------------------------------------------------------------------
// Every scaling the bfloat operations perform is by a power of two (2^+-104) and commutes with both roundings; a term that the
// reference multiplies by aConversionLookup[>= 2] = 0, or by 2^-104 into the denormals, is far below half an ulp of the term it is
// added to (that one is >= 1e-18) and leaves the sum unchanged either way.  So the VALUES of the recursion do not depend on how the
// reference happens to split them into mantissa and exponent: with both states on ONE running scale (u', h' plain floats,
// value = x' * 2^(104 sc)) the step is U' = fl(r0 + r1), H' = fl(r2 + r3) in every regime -- no exponents in the chain at all, as long
// as u' and h' stay normal floats.  The scale follows the larger state (both are multiplied by 2^104 when it falls below 1e6, so it
// lives in [1e-13, 2e37] and a state up to 25 decades smaller is still normal); values never grow (every coefficient sum is < 1).
// What the reference's split IS needed for is the output -- BFloat::Value's table is filled with expf, so (f, e) and (f 2^104, e - 1)
// print as different doubles -- and it is recovered column by column when the block is re-examined: hmm_canon maps a virtual value to
// the form a value takes when it has only ever been renormalised downwards (mantissa in [1e-18, 1e-18 * 2^104)); the reference's own
// form differs from it only where a sum of two just-renormalised terms lands within a factor 2 above that window (about one column in
// 400), which the re-examination sees as "same value, other form" and hands on to the next column.
HMM_HD void hmm_vstep(float& u, float& h, const float c[4], const float l[4])
{
    const float r0 = hmm_fprod(u, c[0], l[0]), r1 = hmm_fprod(h, c[1], l[1]);
    const float r2 = hmm_fprod(u, c[2], l[2]), r3 = hmm_fprod(h, c[3], l[3]);
    u = H_FADD(r0, r1);
    h = H_FADD(r2, r3);
}

HMM_HD BF hmm_canon(float v, int e)
{
    const float L = 1.0e-18f, TOP = 1.0e-18f * 2.028240960365167e+31f;   // exact: a power-of-two multiple of the float 1e-18f
    if (!(v > 0.0f)) return BF{0.f, -BF_INF};
    while (v >= TOP) { v = H_FMUL(v, 4.930380657631324e-32f); ++e; }
    while (v < L) { v = H_FMUL(v, 2.028240960365167e+31f); --e; }
    return BF{v, e};
}

HMM_HD bool hmm_bf_same(BF a, BF b) { return H_F2U(a.f) == H_F2U(b.f) && a.e == b.e; }

// a bfloat pair on one scale: sc = the larger exponent, the other mantissa multiplied by M(its exponent - sc) (0 when two or more
// exponents below: such a state is invisible in every sum until it has been replaced)
HMM_HD void hmm_to_virtual(BF u, BF h, float& vu, float& vh, int& sc)
{
    sc = u.e > h.e ? u.e : h.e;
    vu = H_FMUL(u.f, hmm_mexp(u.e - sc));
    vh = H_FMUL(h.f, hmm_mexp(h.e - sc));
}

// ---- tables shared by the kernels and the test-only host drivers ----
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

static void build_fast_tab(const HmmExactModel& m, HmmFastTab* t)
{
    for (int x = 0; x < 8; ++x)
        for (int k = 0; k < 4; ++k) {
            const float hi = (float)m.te[x][k];
            t->hi[x][k] = hi;
            t->lo[x][k] = (float)(m.te[x][k] - (double)hi);  // the subtraction is exact
        }
}

#ifndef MCU_HOST_EMU
// thread t < n: forward chain of string t; thread n + t: its backward chain (so that the lanes of a warp walk the same way).
// fh / bh: the homologous-state value of every column.
__global__ void __launch_bounds__(64) hmm_exact_chain_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, u32 n, HmmExactModel m, HmmFastTab ft,
                                                            BF* __restrict__ fh, BF* __restrict__ bh, BF* __restrict__ total, u32* __restrict__ err)
{
    __shared__ float chi_s[2][8][4], clo_s[2][8][4];  // [forward / backward][symbol][role]
    __shared__ double te_s[2][8][4];
    for (u32 i = threadIdx.x; i < 64; i += blockDim.x) {
        const u32 d = i >> 5, x = (i >> 2) & 7, k = i & 3, src = (d == 0 || k == 0 || k == 3) ? k : 3 - k;
        chi_s[d][x][k] = ft.hi[x][src];
        clo_s[d][x][k] = ft.lo[x][src];
        te_s[d][x][k] = m.te[x][src];
    }
    __syncthreads();
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2ull * n) return;
    const bool fwd = t < n;
    const u32 s = (u32)(fwd ? t : t - n);
    const u64 beg = off[s], end = off[s + 1];
    if (end == beg) return;
    u32 bad = 0;
    const BF one = BF{1.0f, 0};  // double2bfloat(1.0)
    BF u, h;
    if (fwd) {
        const u32 x = sym_index(__ldg(sym + beg), bad);
        u = bf_dprod(one, m.first[x][0], m);
        h = bf_dprod(one, m.first[x][1], m);
        fh[beg] = h;
    } else {
        h = bf_dprod(one, m.stop[1], m);
        u = bf_dprod(one, m.stop[0], m);
        bh[end - 1] = h;
    }
    const int d = fwd ? 0 : 1;
    // forward step k = 1 .. len-1 takes the symbol of column k and gives column k; backward step gives column i - 1 from the symbol of column i
    for (u64 k = 1; k < end - beg; ++k) {
        const u64 i = fwd ? beg + k : end - k;
        const u32 x = sym_index(__ldg(sym + i), bad);
        if (!hmm_float_step(u, h, chi_s[d][x], clo_s[d][x])) hmm_exact_step(u, h, te_s[d][x], fwd, m);
        if (fwd) fh[i] = h; else bh[i - 1] = h;
    }
    if (fwd) {
        BF p = bf_dprod(u, m.stop[0], m);
        bf_sum_accum(p, bf_dprod(h, m.stop[1], m));
        total[s] = p;
    }
    if (bad) atomicOr(err, 1u);
}

// Few, long strings: one CTA of two warps per (string, direction), a block of 32 columns at a time, three blocks in flight.
//   warp 0, lane 0   runs the serial recurrence of block k in the VIRTUAL mantissa domain (hmm_vstep: 4 FMUL + 4 FFMA + 2 FADD per
//                    column, dependency chain FMUL -> FFMA -> FADD, no exponents, no regimes), eight columns per group with their
//                    coefficients in registers, parking (u', h', scale) behind every column in shared memory; once per group it
//                    looks at the scale (both states times 2^104 when the larger one has fallen below 1e6).
//   warp 1           meanwhile (a) re-examines block k - 1, one column per lane: the state parked in front of the column in the
//                    reference's form (hmm_canon; the lane behind a "same value, other form" column takes its neighbour's true
//                    form), one exact column from there (hmm_float_step, or hmm_exact_step when a product of it is hazardous), and
//                    the comparison with what is parked behind the column; (b) writes that block's 32 results -- the reference's
//                    (mantissa, exponent) pairs -- with one coalesced store; (c) looks up the coefficients of block k + 1 (symbols
//                    fetched two blocks ahead).
// A column whose parked VALUE is not what warp 1 gets (a rounding hazard that mattered: about one column in 20 million) is corrected;
// the chain is then taken up again behind it and block k, which started from the wrong state, is run again.
// MAUVE_CUDA_HMM_FP64=1 (tests): every column through hmm_exact_step.
struct HmmBlockBuf {
    float4 chi[40], clo[40];  // coefficients of the block's columns, role order, high and low parts; rows 32.. and the rows past a short
                              // block's end: identity (a group of eight may run past the end)
    float4 st[41];            // st[j] = the state in front of column j as a bfloat pair (u.f, h.f, bits of u.e, bits of h.e): the
                              // chain parks both mantissas on its running scale, a repair writes the reference's own form
    u8 xs[32];
};
struct HmmWarpSmem {
    double te[8][4];          // role order
    float4 chi_s[8], clo_s[8];
    HmmBlockBuf b[3];
    int bad;                  // first corrected column of the block re-examined in this iteration, or -1
};

__device__ __forceinline__ void hmm_lane0_exact(HmmWarpSmem& sm, HmmBlockBuf& bb, u32 cnt, bool fwd, const HmmExactModel& m)
{
    const float4 s0 = bb.st[0];
    BF u = BF{s0.x, __float_as_int(s0.z)}, h = BF{s0.y, __float_as_int(s0.w)};
    for (u32 j = 0; j < cnt; ++j) {
        hmm_exact_step(u, h, sm.te[bb.xs[j]], fwd, m);
        bb.st[j + 1] = make_float4(u.f, h.f, __int_as_float(u.e), __int_as_float(h.e));
    }
}

// warp 0, lane 0: columns [start, cnt) of the block in groups of eight (rows past the block's end hold identity coefficients: nothing
// happens there), from the virtual state (u, h, sc), which it carries on
// fault_every (tests only, MAUVE_CUDA_HMM_TEST_FAULT): every fault_every-th group ends with the last mantissa bit of u' flipped -- in the
// parked state and in the chain -- which is what a rounding hazard that mattered looks like to the re-examination
template <bool FAULT>
__device__ __forceinline__ void hmm_lane0_run(HmmBlockBuf& bb, u32 start, u32 cnt, float& u, float& h, int& sc, u32 fault_every = 0, u32* fault_ctr = nullptr)
{
    for (u32 j = start; j < cnt; j += 8) {
        float4 c[8], l[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            c[q] = bb.chi[j + q];
            l[q] = bb.clo[j + q];
        }
        if (fmaxf(u, h) < 1.0e6f) {   // the scale follows the larger state (exact: a power of two; neither state can overflow)
            u = __fmul_rn(u, 2.028240960365167e+31f);
            h = __fmul_rn(h, 2.028240960365167e+31f);
            --sc;
        }
        const float scf = __int_as_float(sc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float cc[4] = {c[q].x, c[q].y, c[q].z, c[q].w}, ll[4] = {l[q].x, l[q].y, l[q].z, l[q].w};
            hmm_vstep(u, h, cc, ll);
            bb.st[j + q + 1] = make_float4(u, h, scf, scf);
        }
        if (FAULT && j + 8 <= cnt && ++*fault_ctr % fault_every == 0) {
            u = __uint_as_float(__float_as_uint(u) ^ 1u);
            bb.st[j + 8] = make_float4(u, h, scf, scf);
        }
    }
    const float4 e = bb.st[cnt];   // the state behind column cnt - 1 (the last group may have run past it)
    u = e.x;
    h = e.y;
    sc = __float_as_int(e.z);
}

// the same from the state parked at st[start] (a bfloat pair in any form)
__device__ __forceinline__ void hmm_lane0_chain(HmmBlockBuf& bb, u32 start, u32 cnt)
{
    const float4 s0 = bb.st[start];
    float u, h;
    int sc;
    hmm_to_virtual(BF{s0.x, __float_as_int(s0.z)}, BF{s0.y, __float_as_int(s0.w)}, u, h, sc);
    hmm_lane0_run<false>(bb, start, cnt, u, h, sc);
}

// hmm_canon for a state the chain parked: one step either way covers its range (the scale keeps the larger state in [1e-13, 2e37])
__device__ __forceinline__ BF hmm_canon_parked(float v, int e)
{
    const float L = 1.0e-18f, TOP = 1.0e-18f * 2.028240960365167e+31f;
    if (v >= TOP) { v = __fmul_rn(v, 4.930380657631324e-32f); ++e; }
    else if (v < L) { v = __fmul_rn(v, 2.028240960365167e+31f); --e; }
    if (!(v >= L && v < TOP)) return hmm_canon(v, e);   // (zero, or a repair's state far from the scale)
    return BF{v, e};
}

// a verifier warp: columns [start, cnt) of a block, one per lane.  `first` = the state in front of column `start` in the reference's
// form (or the best guess at it: the caller compares and comes back).  out_u / out_h: the reference's state behind this lane's column
// (valid for start <= lane < cnt up to the returned column).  Returns the first column whose parked VALUE is not the reference's (its
// lane has written the reference's state behind it to st), or 32.
__device__ __forceinline__ u32 hmm_reexamine(HmmWarpSmem& sm, HmmBlockBuf& bb, u32 lane, u32 start, u32 cnt, BF first_u, BF first_h, bool fwd,
                                             const HmmExactModel& m, BF& out_u, BF& out_h, unsigned long long& n_exact)
{
    const bool mine = lane >= start && lane < cnt;
    BF want_u = first_u, want_h = first_h;
    float cc[4] = {1.f, 0.f, 0.f, 1.f}, ll[4] = {0.f, 0.f, 0.f, 0.f};
    if (mine) {
        const float4 w = bb.st[lane + 1];
        const float4 c = bb.chi[lane], l = bb.clo[lane];
        cc[0] = c.x; cc[1] = c.y; cc[2] = c.z; cc[3] = c.w;
        ll[0] = l.x; ll[1] = l.y; ll[2] = l.z; ll[3] = l.w;
        want_u = hmm_canon_parked(w.x, __float_as_int(w.z));
        want_h = hmm_canon_parked(w.y, __float_as_int(w.w));
    }
    // the state in front of a column = what is parked behind the column before it
    BF guess_u, guess_h;
    guess_u.f = __shfl_up_sync(0xffffffffu, want_u.f, 1); guess_u.e = __shfl_up_sync(0xffffffffu, want_u.e, 1);
    guess_h.f = __shfl_up_sync(0xffffffffu, want_h.f, 1); guess_h.e = __shfl_up_sync(0xffffffffu, want_h.e, 1);
    if (lane <= start) { guess_u = first_u; guess_h = first_h; }
    BF in_u = guess_u, in_h = guess_h;
    bool value_bad = false;
    for (int pass = 0; pass < 34; ++pass) {
        bool other_form = false;
        value_bad = false;
        if (mine) {
            out_u = in_u;
            out_h = in_h;
            if (!hmm_float_step(out_u, out_h, cc, ll)) {
                hmm_exact_step(out_u, out_h, sm.te[bb.xs[lane]], fwd, m);
                if (pass == 0) ++n_exact;
            }
            if (!(hmm_bf_same(out_u, want_u) && hmm_bf_same(out_h, want_h))) {
                const bool same_value = hmm_bf_same(hmm_canon(out_u.f, out_u.e), want_u) && hmm_bf_same(hmm_canon(out_h.f, out_h.e), want_h);
                other_form = same_value;
                value_bad = !same_value;
            }
        }
        if (__ballot_sync(0xffffffffu, other_form) == 0u) break;   // (the common case: nothing to hand on)
        // the lane behind a column that came out in another form than hmm_canon gives takes that form as its input
        const float puf = __shfl_up_sync(0xffffffffu, out_u.f, 1), phf = __shfl_up_sync(0xffffffffu, out_h.f, 1);
        const int pue = __shfl_up_sync(0xffffffffu, out_u.e, 1), phe = __shfl_up_sync(0xffffffffu, out_h.e, 1);
        const bool pform = __shfl_up_sync(0xffffffffu, other_form ? 1 : 0, 1) != 0;
        bool changed = false;
        if (mine && lane > start) {
            const BF nu = pform ? BF{puf, pue} : guess_u, nh = pform ? BF{phf, phe} : guess_h;
            if (!(hmm_bf_same(nu, in_u) && hmm_bf_same(nh, in_h))) {
                in_u = nu;
                in_h = nh;
                changed = true;
            }
        }
        if (__ballot_sync(0xffffffffu, changed) == 0u) break;
    }
    const u32 badmask = __ballot_sync(0xffffffffu, value_bad);
    if (badmask == 0u) return 32u;
    const u32 jb = (u32)__ffs((int)badmask) - 1u;
    if (lane == jb) bb.st[jb + 1] = make_float4(out_u.f, out_h.f, __int_as_float(out_u.e), __int_as_float(out_h.e));
    __syncwarp();
    return jb;
}

constexpr int HV = 6;        // verifier warps = blocks per iteration (measured on 5 M columns: 4 -> 123 ms)
constexpr int HR = 3 * HV;   // block buffers: the blocks being re-examined, chained and staged
struct HmmWarpSmem2 {
    HmmWarpSmem base;        // te, chi_s, clo_s, bad (b[0..2] of it are the first three buffers)
    HmmBlockBuf more[HR - 3];
    float4 exit_st[HV + 1];  // the reference's state behind block i of the ones re-examined in this iteration ([0]: behind the block before them)
    int bad_blk, bad_col;    // first column of this iteration whose parked value was not the reference's
};

__device__ __forceinline__ void hmm_bar_verifiers()
{
    asm volatile("bar.sync 1, %0;" ::"r"(32 * HV) : "memory");
}

template <bool FAULT>
__global__ void __launch_bounds__(32 * (1 + HV)) hmm_exact_chain_warp_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, u32 n, HmmExactModel m,
                                                                            HmmFastTab ft, int force_exact, int fault_every, BF* __restrict__ fh, BF* __restrict__ bh,
                                                                            BF* __restrict__ total, u32* __restrict__ err,
                                                                            unsigned long long* __restrict__ counters)
{
    __shared__ HmmWarpSmem2 sm2;
    HmmWarpSmem& sm = sm2.base;
    auto buf = [&](u64 blk) -> HmmBlockBuf& {
        const u32 i = (u32)(blk % HR);
        return i < 3 ? sm.b[i] : sm2.more[i - 3];
    };
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 s = blockIdx.x >> 1;
    const bool fwd = (blockIdx.x & 1) == 0;
    const u64 beg = off[s], end = off[s + 1];
    if (end == beg) return;
    if (warp == 0) {
        const u32 x = lane >> 2, k = lane & 3, src = (fwd || k == 0 || k == 3) ? k : 3 - k;
        sm.te[x][k] = m.te[x][src];
        reinterpret_cast<float*>(sm.chi_s)[lane] = ft.hi[x][src];
        reinterpret_cast<float*>(sm.clo_s)[lane] = ft.lo[x][src];
    }
    const u64 len = end - beg;
    u32 bad = 0;
    const BF one = BF{1.0f, 0};
    BF ent_u, ent_h;
    if (fwd) {
        const u32 x = sym_index(__ldg(sym + beg), bad);
        ent_u = bf_dprod(one, m.first[x][0], m);
        ent_h = bf_dprod(one, m.first[x][1], m);
    } else {
        ent_h = bf_dprod(one, m.stop[1], m);
        ent_u = bf_dprod(one, m.stop[0], m);
    }
    if (threadIdx.x == 0) {
        if (fwd) fh[beg] = ent_h; else bh[end - 1] = ent_h;
        sm.b[0].st[0] = make_float4(ent_u.f, ent_h.f, __int_as_float(ent_u.e), __int_as_float(ent_h.e));
        sm2.exit_st[0] = sm.b[0].st[0];
        sm2.bad_blk = 0x7fffffff;
        sm2.bad_col = 0;
    }
    const u64 steps = len - 1;
    const u64 nb = (steps + 31) / 32;
    auto sym_of_step = [&](u64 k) -> u64 { return fwd ? beg + 1 + k : end - 1 - k; };  // index of the symbol step k consumes
    auto cnt_of = [&](u64 k) -> u32 { return (u32)min((u64)32, steps - k * 32); };
    auto stage = [&](u64 blk, u32 x) {   // a verifier warp: the coefficient rows of a block (x: this lane's symbol index)
        HmmBlockBuf& nbuf = buf(blk);
        const bool live = blk * 32 + lane < steps;
        nbuf.chi[lane] = live ? sm.chi_s[x] : make_float4(1.f, 0.f, 0.f, 1.f);
        nbuf.clo[lane] = live ? sm.clo_s[x] : make_float4(0.f, 0.f, 0.f, 0.f);
        nbuf.xs[lane] = (u8)x;
        if (lane < 8) {
            nbuf.chi[32 + lane] = make_float4(1.f, 0.f, 0.f, 1.f);
            nbuf.clo[32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto load_sym = [&](u64 blk) -> u32 {
        return (blk < nb && blk * 32 + lane < steps) ? sym_index(__ldg(sym + sym_of_step(blk * 32 + lane)), bad) : 0u;
    };
    auto store_results = [&](u64 blk, u32 upto_lane, BF oh) {   // lanes 0 .. upto_lane of the block
        if (lane <= upto_lane && lane < cnt_of(blk)) {
            if (fwd) fh[beg + 1 + blk * 32 + lane] = oh;
            else bh[end - 2 - blk * 32 - lane] = oh;
        }
    };
    __syncthreads();
    if (warp >= 1) {   // the rows of the first 2 HV blocks
        for (int r = 0; r < 2; ++r) {
            const u64 blk = (u64)r * HV + (warp - 1);
            if (blk < nb) stage(blk, load_sym(blk));
        }
    }
    __syncthreads();
    unsigned long long n_rounds = 0, n_exact = 0;
    u32 fault_ctr = 0;
    u64 k = 0;       // first block the chain takes in this iteration
    u64 vk = 0;      // first block to re-examine, nv of them
    u32 nv = 0;
    while (k < nb || nv) {
        if (warp == 0) {
            // ---- the chain: blocks k .. k + HV - 1 ----
            if (lane == 0) {
                float cu = 0.f, ch = 0.f;   // the chain's state, carried from block to block
                int csc = 0;
                for (u32 i = 0; i < (u32)HV && k + i < nb; ++i) {
                    const u64 blk = k + i;
                    HmmBlockBuf& bb = buf(blk);
                    if (force_exact) {
                        if (blk) bb.st[0] = buf(blk - 1).st[cnt_of(blk - 1)];
                        hmm_lane0_exact(sm, bb, cnt_of(blk), fwd, m);
                    } else {
                        if (i == 0) {   // from what is parked behind the block before (after a repair: the reference's own form)
                            const float4 s0 = blk ? buf(blk - 1).st[cnt_of(blk - 1)] : bb.st[0];
                            hmm_to_virtual(BF{s0.x, __float_as_int(s0.z)}, BF{s0.y, __float_as_int(s0.w)}, cu, ch, csc);
                        }
                        bb.st[0] = make_float4(cu, ch, __int_as_float(csc), __int_as_float(csc));
                        hmm_lane0_run<FAULT>(bb, 0, cnt_of(blk), cu, ch, csc, (u32)fault_every, &fault_ctr);
                    }
                }
            }
        } else {
            // ---- verifier warp w: re-examine block vk + w, store its results, stage block k + HV + w ----
            const u32 w = warp - 1;
            const u64 sb = k + HV + w;
            const u32 sx = load_sym(sb);   // in flight during the re-examination
            const bool have = w < nv;
            const u64 vb = vk + w;
            BF ou = BF{0.f, 0}, oh = BF{0.f, 0};
            u32 jb = 32u, pc = 0;
            BF used_u = BF{0.f, 0}, used_h = BF{0.f, 0};
            if (have) {
                ++n_rounds;
                HmmBlockBuf& pb = buf(vb);
                pc = cnt_of(vb);
                if (force_exact) {
                    n_exact += lane < pc ? 1u : 0u;
                    if (lane < pc) {
                        const float4 r = pb.st[lane + 1];
                        ou = BF{r.x, __float_as_int(r.z)};
                        oh = BF{r.y, __float_as_int(r.w)};
                    }
                } else {
                    // the state in front of the block: known for the first verifier, for the others the form hmm_canon gives what the
                    // chain parked there (checked against the neighbour's result below)
                    const float4 e0 = w == 0 ? sm2.exit_st[0] : pb.st[0];
                    used_u = w == 0 ? BF{e0.x, __float_as_int(e0.z)} : hmm_canon_parked(e0.x, __float_as_int(e0.z));
                    used_h = w == 0 ? BF{e0.y, __float_as_int(e0.w)} : hmm_canon_parked(e0.y, __float_as_int(e0.w));
                    jb = hmm_reexamine(sm, pb, lane, 0, pc, used_u, used_h, fwd, m, ou, oh, n_exact);
                }
                const u32 lastl = jb < 32u ? jb : pc - 1;
                const float xuf = __shfl_sync(0xffffffffu, ou.f, lastl), xhf = __shfl_sync(0xffffffffu, oh.f, lastl);
                const int xue = __shfl_sync(0xffffffffu, ou.e, lastl), xhe = __shfl_sync(0xffffffffu, oh.e, lastl);
                if (lane == 0) sm2.exit_st[w + 1] = make_float4(xuf, xhf, __int_as_float(xue), __int_as_float(xhe));
            }
            // the blocks' entry states, in order: verifier w waits for w - 1 (w rounds), comes back when its guess was another form
            for (u32 r = 1; r < (u32)HV; ++r) {
                hmm_bar_verifiers();
                if (have && w == r && !force_exact) {
                    const float4 e0 = sm2.exit_st[w];
                    const BF tu = BF{e0.x, __float_as_int(e0.z)}, th = BF{e0.y, __float_as_int(e0.w)};
                    if (!(hmm_bf_same(tu, used_u) && hmm_bf_same(th, used_h))) {
                        HmmBlockBuf& pb = buf(vb);
                        jb = hmm_reexamine(sm, pb, lane, 0, pc, tu, th, fwd, m, ou, oh, n_exact);
                        const u32 lastl = jb < 32u ? jb : pc - 1;
                        const float xuf = __shfl_sync(0xffffffffu, ou.f, lastl), xhf = __shfl_sync(0xffffffffu, oh.f, lastl);
                        const int xue = __shfl_sync(0xffffffffu, ou.e, lastl), xhe = __shfl_sync(0xffffffffu, oh.e, lastl);
                        if (lane == 0) sm2.exit_st[w + 1] = make_float4(xuf, xhf, __int_as_float(xue), __int_as_float(xhe));
                    }
                }
            }
            if (have && jb < 32u && lane == 0) {
                const int prev = atomicMin(&sm2.bad_blk, (int)(vb & 0x3fffffff));
                (void)prev;
            }
            hmm_bar_verifiers();
            if (have && jb < 32u && lane == 0 && sm2.bad_blk == (int)(vb & 0x3fffffff)) sm2.bad_col = (int)jb;
            // what stands: every block in front of the first corrected column, and that block up to the column
            if (have) {
                const int bb_ = sm2.bad_blk;
                const int me = (int)(vb & 0x3fffffff);
                if (me < bb_) store_results(vb, 31u, oh);
                else if (me == bb_) store_results(vb, jb, oh);
            }
            if (sb < nb) stage(sb, sx);
        }
        __syncthreads();
        if (sm2.bad_blk != 0x7fffffff) {
            // ---- a parked value was not the reference's: repair that block behind the column (chain lane and verifier 0 take turns),
            //      then the chain starts again behind the block ----
            const u64 rb = (vk & ~(u64)0x3fffffff) | (u64)sm2.bad_blk;
            int jb = sm2.bad_col;
            HmmBlockBuf& pb = buf(rb);
            const u32 pc = cnt_of(rb);
            // the reference's state behind the corrected column: verifier (rb - vk) left it in exit_st
            if (warp == 1) {
                const float4 e0 = sm2.exit_st[(u32)(rb - vk) + 1];
                ent_u = BF{e0.x, __float_as_int(e0.z)};
                ent_h = BF{e0.y, __float_as_int(e0.w)};
            }
            __syncthreads();
            while (jb >= 0) {
                const u32 vstart = (u32)jb + 1u;
                if (vstart >= pc) break;   // the corrected column was the block's last
                if (threadIdx.x == 0) {
                    sm.bad = -1;
                    hmm_lane0_chain(pb, vstart, pc);
                }
                __syncthreads();
                if (warp == 1) {
                    ++n_rounds;
                    BF ou = ent_u, oh = ent_h;
                    const u32 j2 = hmm_reexamine(sm, pb, lane, vstart, pc, ent_u, ent_h, fwd, m, ou, oh, n_exact);
                    const u32 upto = j2 < 32u ? j2 : pc - 1;
                    if (lane >= vstart) store_results(rb, upto, oh);
                    ent_u.f = __shfl_sync(0xffffffffu, ou.f, upto); ent_u.e = __shfl_sync(0xffffffffu, ou.e, upto);
                    ent_h.f = __shfl_sync(0xffffffffu, oh.f, upto); ent_h.e = __shfl_sync(0xffffffffu, oh.e, upto);
                    if (lane == 0) sm.bad = j2 < 32u ? (int)j2 : -1;
                }
                __syncthreads();
                jb = sm.bad;
                __syncthreads();
            }
            if (warp == 1 && lane == 0) {
                const float4 e = make_float4(ent_u.f, ent_h.f, __int_as_float(ent_u.e), __int_as_float(ent_h.e));
                pb.st[pc] = e;          // the state the chain starts the next block from
                sm2.exit_st[0] = e;     // and the state the next re-examination starts from
                sm2.bad_blk = 0x7fffffff;
                sm2.bad_col = 0;
            }
            __syncthreads();
            k = rb + 1;
            vk = rb + 1;
            nv = 0;
            continue;
        }
        // ---- next iteration: re-examine what was just chained ----
        if (threadIdx.x == 32 && nv) sm2.exit_st[0] = sm2.exit_st[nv];
        vk = k;
        nv = (u32)min((u64)HV, nb > k ? nb - k : 0);
        k = min(nb, k + HV);
        __syncthreads();
    }
    if (fwd && threadIdx.x == 32) {
        float4 r = sm2.exit_st[0];   // the reference's state behind the last column
        if (force_exact && nb) r = buf(nb - 1).st[cnt_of(nb - 1)];
        BF u = BF{r.x, __float_as_int(r.z)}, h = BF{r.y, __float_as_int(r.w)};
        BF p = bf_dprod(u, m.stop[0], m);
        bf_sum_accum(p, bf_dprod(h, m.stop[1], m));
        total[s] = p;
    }
    if (counters && warp >= 1) {
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 16);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 8);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 4);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 2);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 1);
        if (lane == 0) {
            if (warp == 1) atomicAdd(counters, steps);
            atomicAdd(counters + 1, n_rounds);
            atomicAdd(counters + 2, n_exact);
        }
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
    const double po = hmm_posterior_value(fh[i], bh[i], total[lo], m);
    if (post) post[i] = po;
    pred[i] = po >= 0.9 ? 'H' : 'N';
}

struct HmmState {
    DevBuf sym, off, chunk_first, prod, fin, bin, scratch, pred, post, err, fh, bh, total;
    unsigned long long counters[3] = {0, 0, 0};  // last call of the warp chain: columns, chain rounds, columns through hmm_exact_step
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
        (r = st.pred.reserve(total + 16)) || (r = st.post.reserve(total * 8 + 16)) || (r = st.err.reserve(32)) ||
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
        HmmFastTab ft;
        build_fast_tab(xm, &ft);
        const int force_exact = getenv("MAUVE_CUDA_HMM_FP64") != nullptr;   // tests: every column operation by operation
        const int fault_every = getenv("MAUVE_CUDA_HMM_TEST_FAULT") ? atoi(getenv("MAUVE_CUDA_HMM_TEST_FAULT")) : 0;   // tests: see hmm_lane0_run
        MCU_CUDA(cudaMemsetAsync(st.err.as<u32>() + 2, 0, 24, s));
        // few chains: a warp each (latency-optimised); many chains: a thread each (throughput)
        if (2 * n <= (u64)sm_count() * 64) {
            if (fault_every > 0)
                hmm_exact_chain_warp_kernel<true><<<(unsigned)(2 * n), 32 * (1 + HV), 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), (u32)n, xm, ft, force_exact, fault_every > 0 ? fault_every : 0, st.fh.as<BF>(),
                                                                         st.bh.as<BF>(), st.total.as<BF>(), st.err.as<u32>(),
                                                                         reinterpret_cast<unsigned long long*>(st.err.as<u32>() + 2));
            else
                hmm_exact_chain_warp_kernel<false><<<(unsigned)(2 * n), 32 * (1 + HV), 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), (u32)n, xm, ft, force_exact, fault_every > 0 ? fault_every : 0, st.fh.as<BF>(),
                                                                         st.bh.as<BF>(), st.total.as<BF>(), st.err.as<u32>(),
                                                                         reinterpret_cast<unsigned long long*>(st.err.as<u32>() + 2));
        } else
            hmm_exact_chain_kernel<<<(unsigned)div_up(2 * n, 64), 64, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), (u32)n, xm, ft, st.fh.as<BF>(), st.bh.as<BF>(),
                                                                              st.total.as<BF>(), st.err.as<u32>());
        hmm_exact_posterior_kernel<<<(unsigned)div_up(total, 256), 256, 0, s>>>(st.off.as<u64>(), (u32)n, total, xm, st.fh.as<BF>(), st.bh.as<BF>(),
                                                                                st.total.as<BF>(), st.pred.as<char>(),
                                                                                post_out ? st.post.as<double>() : nullptr);
    }
    MCU_CUDA(cudaEventRecord(st.e1, s));
    MCU_CUDA(cudaGetLastError());
    u32 err_flag = 0;
    MCU_CUDA(cudaMemcpyAsync(&err_flag, st.err.p, 4, cudaMemcpyDeviceToHost, s));
    st.counters[0] = st.counters[1] = st.counters[2] = 0;
    if (!scan_mode) MCU_CUDA(cudaMemcpyAsync(st.counters, st.err.as<u32>() + 2, 24, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaMemcpyAsync(pred_out, st.pred.p, total, cudaMemcpyDeviceToHost, s));
    if (post_out) MCU_CUDA(cudaMemcpyAsync(post_out, st.post.p, total * 8, cudaMemcpyDeviceToHost, s));
    MCU_CUDA(cudaStreamSynchronize(s));
    if (err_flag) { set_error("mcu_hmm_batch: symbol outside '1'..'8'"); return MCU_EINVAL; }
    if (device_ms) cudaEventElapsedTime(device_ms, st.e0, st.e1);
    return MCU_OK;
}


void hmm_release()
{
    HmmState& st = g_hmm;
    DevBuf* bufs[] = {&st.sym, &st.off, &st.chunk_first, &st.prod, &st.fin, &st.bin, &st.scratch, &st.pred, &st.post, &st.err, &st.fh, &st.bh, &st.total};
    for (DevBuf* b : bufs) b->release();
}

void hmm_last_counters(u64* out3)
{
    out3[0] = g_hmm.counters[0];
    out3[1] = g_hmm.counters[1];
    out3[2] = g_hmm.counters[2];
}

#else  // MCU_HOST_EMU ----------------------------------------------------------------------------------------------------
}  // namespace mcu

// TEST-ONLY host drivers of the value functions above (tests/_emu.py builds this file with -DMCU_HOST_EMU into
// tests/_emu/libmcu_emu.so; they are not part of libmauve_cuda.so).
//
// emu_hmm_chain: one chain (forward or backward) of one string the way the re-examination and the thread-per-string kernel evaluate a
// column: hmm_float_step, hmm_exact_step at hazardous ones -- and, beside it, hmm_exact_step at EVERY column from the same state (the
// truth).  out_f / out_e: the homologous-state bfloat after every step (steps = n - 1, chain order).  counts[1] = columns evaluated by
// hmm_float_step, counts[2] = columns where it differs from the truth (must be 0), counts[3] = hazardous columns.
extern "C" void emu_hmm_chain(const unsigned char* sym, unsigned long long n, const double* params21, int fwd, float* out_f, int* out_e,
                              unsigned long long* counts)
{
    using namespace mcu;
    HmmExactModel m;
    build_exact_model(params21, &m);
    HmmFastTab ft;
    build_fast_tab(m, &ft);
    counts[0] = counts[1] = counts[2] = counts[3] = counts[4] = 0;
    if (n == 0) return;
    const BF one = BF{1.0f, 0};
    BF u, h;
    if (fwd) {
        const u32 x = (u32)(sym[0] - '1') & 7u;
        u = bf_dprod(one, m.first[x][0], m);
        h = bf_dprod(one, m.first[x][1], m);
    } else {
        h = bf_dprod(one, m.stop[1], m);
        u = bf_dprod(one, m.stop[0], m);
    }
    for (unsigned long long k = 1; k < n; ++k) {
        const u32 x = (u32)(sym[fwd ? k : n - k] - '1') & 7u;
        double c[4];
        float ch[4], cl[4];
        for (int r = 0; r < 4; ++r) {
            const int src = (fwd || r == 0 || r == 3) ? r : 3 - r;
            c[r] = m.te[x][src];
            ch[r] = ft.hi[x][src];
            cl[r] = ft.lo[x][src];
        }
        BF tu = u, th = h;
        hmm_exact_step(tu, th, c, fwd != 0, m);
        auto same = [&](BF a, BF b) { return a.e == tu.e && b.e == th.e && memcmp(&a.f, &tu.f, 4) == 0 && memcmp(&b.f, &th.f, 4) == 0; };
        BF gu = u, gh = h;
        const bool float_ok = hmm_float_step(gu, gh, ch, cl);
        if (float_ok && !same(gu, gh)) ++counts[2];
        if (float_ok) ++counts[1];
        else ++counts[3];
        u = tu;
        h = th;
        out_f[k - 1] = h.f;
        out_e[k - 1] = h.e;
    }
}

// emu_hmm_run: run() for one string the way hmm_batch evaluates it (both chains through emu_hmm_chain's hybrid, the posterior
// through hmm_posterior_value); counts as emu_hmm_chain, summed over the two chains.  Returns 0, or -1 when out of memory.
extern "C" int emu_hmm_run(const unsigned char* sym, unsigned long long n, const double* params21, char* pred, double* post, unsigned long long* counts)
{
    using namespace mcu;
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    if (n == 0) return 0;
    HmmExactModel m;
    build_exact_model(params21, &m);
    float* ff = (float*)malloc(n * 4), *bf_ = (float*)malloc(n * 4);
    int* fe = (int*)malloc(n * 4), *be = (int*)malloc(n * 4);
    if (!ff || !bf_ || !fe || !be) { free(ff); free(bf_); free(fe); free(be); return -1; }
    const BF one = BF{1.0f, 0};
    unsigned long long c4[5];
    // column 0 of the forward values and column n - 1 of the backward values are the chains' starting states
    {
        const u32 x = (u32)(sym[0] - '1') & 7u;
        const BF h0 = bf_dprod(one, m.first[x][1], m);
        ff[0] = h0.f; fe[0] = h0.e;
        const BF hl = bf_dprod(one, m.stop[1], m);
        bf_[n - 1] = hl.f; be[n - 1] = hl.e;
    }
    emu_hmm_chain(sym, n, params21, 1, ff + 1, fe + 1, c4);
    for (int i = 0; i < 4; ++i) counts[i] += c4[i];
    // the backward driver writes in chain order (column n-2 first): into a scratch, then reversed
    float* tf = (float*)malloc(n * 4);
    int* te = (int*)malloc(n * 4);
    if (!tf || !te) { free(ff); free(bf_); free(fe); free(be); free(tf); free(te); return -1; }
    emu_hmm_chain(sym, n, params21, 0, tf, te, c4);
    for (int i = 0; i < 4; ++i) counts[i] += c4[i];
    for (unsigned long long k = 1; k < n; ++k) { bf_[n - 1 - k] = tf[k - 1]; be[n - 1 - k] = te[k - 1]; }
    // P = (T7 * F_U(L)) (+)= (T6 * F_H(L)): needs the last unrelated-state value, which the chain driver does not return: redo the
    // forward chain's last state exactly (the hybrid equals it, that is what counts[2] == 0 says)
    BF u, h;
    {
        const u32 x = (u32)(sym[0] - '1') & 7u;
        u = bf_dprod(one, m.first[x][0], m);
        h = bf_dprod(one, m.first[x][1], m);
        for (unsigned long long k = 1; k < n; ++k) {
            const u32 xx = (u32)(sym[k] - '1') & 7u;
            hmm_exact_step(u, h, m.te[xx], true, m);
        }
    }
    BF p = bf_dprod(u, m.stop[0], m);
    bf_sum_accum(p, bf_dprod(h, m.stop[1], m));
    for (unsigned long long i = 0; i < n; ++i) {
        const double po = hmm_posterior_value(BF{ff[i], fe[i]}, BF{bf_[i], be[i]}, p, m);
        if (post) post[i] = po;
        pred[i] = po >= 0.9 ? 'H' : 'N';
    }
    free(ff); free(bf_); free(fe); free(be); free(tf); free(te);
    return 0;
}

// emu_hmm_vchain: the virtual-domain chain (hmm_vstep + the rescaling rule of the kernel) over a whole string, beside the exact chain.
// counts[0] = columns where hmm_canon(virtual) is the reference's very form, counts[1] = columns with the same value in another form
// (handed on by the re-examination), counts[2] = columns whose VALUE differs (the chain would be repaired there: hazards),
// counts[3] = rescalings, counts[4] = columns where a virtual state was below 1e-30 (close to the denormals).
extern "C" void emu_hmm_vchain(const unsigned char* sym, unsigned long long n, const double* params21, int fwd, unsigned long long* counts)
{
    using namespace mcu;
    HmmExactModel m;
    build_exact_model(params21, &m);
    HmmFastTab ft;
    build_fast_tab(m, &ft);
    for (int i = 0; i < 5; ++i) counts[i] = 0;
    if (n == 0) return;
    const BF one = BF{1.0f, 0};
    BF u, h;
    if (fwd) {
        const u32 x = (u32)(sym[0] - '1') & 7u;
        u = bf_dprod(one, m.first[x][0], m);
        h = bf_dprod(one, m.first[x][1], m);
    } else {
        h = bf_dprod(one, m.stop[1], m);
        u = bf_dprod(one, m.stop[0], m);
    }
    float vu, vh;
    int sc;
    hmm_to_virtual(u, h, vu, vh, sc);
    for (unsigned long long k = 1; k < n; ++k) {
        const u32 x = (u32)(sym[fwd ? k : n - k] - '1') & 7u;
        double c[4];
        float ch[4], cl[4];
        for (int r = 0; r < 4; ++r) {
            const int src = (fwd || r == 0 || r == 3) ? r : 3 - r;
            c[r] = m.te[x][src];
            ch[r] = ft.hi[x][src];
            cl[r] = ft.lo[x][src];
        }
        hmm_exact_step(u, h, c, fwd != 0, m);   // the truth
        if (((k - 1) & 7) == 0 && fmaxf(vu, vh) < 1.0e6f) {   // the kernel looks at the scale once per group of eight columns
            vu = H_FMUL(vu, 2.028240960365167e+31f);
            vh = H_FMUL(vh, 2.028240960365167e+31f);
            --sc;
            ++counts[3];
        }
        hmm_vstep(vu, vh, ch, cl);
        if (vu < 1.0e-30f || vh < 1.0e-30f) ++counts[4];
        const BF cu = hmm_canon(vu, sc), chh = hmm_canon(vh, sc);
        if (hmm_bf_same(cu, u) && hmm_bf_same(chh, h)) ++counts[0];
        else if (hmm_bf_same(cu, hmm_canon(u.f, u.e)) && hmm_bf_same(chh, hmm_canon(h.f, h.e))) ++counts[1];
        else {
            ++counts[2];
            hmm_to_virtual(u, h, vu, vh, sc);   // what the repair does: the chain goes on from the true state
        }
    }
}

// emu_hmm_fprod: y = (float)((double)v * c) for n pairs against the FP32 form; counts[0] = accepted, counts[1] = hazardous,
// counts[2] = accepted and different (must be 0), counts[3] = hazardous and different
extern "C" void emu_hmm_fprod(const float* v, const double* c, unsigned long long n, unsigned long long* counts)
{
    using namespace mcu;
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (unsigned long long i = 0; i < n; ++i) {
        const float ch = (float)c[i], cl = (float)(c[i] - (double)ch);
        const float y = (float)((double)v[i] * c[i]);
        const float r = hmm_fprod(v[i], ch, cl);
        const bool same = memcmp(&y, &r, 4) == 0;
        if (hmm_fprod_hazard(v[i], ch, cl, r)) { ++counts[1]; if (!same) ++counts[3]; }
        else { ++counts[0]; if (!same) ++counts[2]; }
    }
}

namespace mcu {
#endif  // MCU_HOST_EMU
}  // namespace mcu
