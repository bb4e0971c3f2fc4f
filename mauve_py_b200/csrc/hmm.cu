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
// the sum is symmetric, so Backward is Forward with c1 and c2 exchanged).  Returns false, leaving u and h alone, on a hazard.
HMM_HD bool hmm_float_step(BF& u, BF& h, const float c[4], const float l[4])
{
    const float r0 = hmm_fprod(u.f, c[0], l[0]), r1 = hmm_fprod(h.f, c[1], l[1]);
    const float r2 = hmm_fprod(u.f, c[2], l[2]), r3 = hmm_fprod(h.f, c[3], l[3]);
    u32 w0, w1, w2, w3;
    const u32 hz = hmm_prod_check(u.f, c[0], l[0], r0, w0) | hmm_prod_check(h.f, c[1], l[1], r1, w1) |
                   hmm_prod_check(u.f, c[2], l[2], r2, w2) | hmm_prod_check(h.f, c[3], l[3], r3, w3);
    if (hz) return false;
    const int e0 = u.e - (int)w0, e1 = h.e - (int)w1, e2 = u.e - (int)w2, e3 = h.e - (int)w3;
    const int eu = e0 > e1 ? e0 : e1, eh = e2 > e3 ? e2 : e3;
    const float nu = H_FADD(H_FMUL(r0, hmm_mexp(u.e - eu)), H_FMUL(r1, hmm_mexp(h.e - eu)));
    const float nh = H_FADD(H_FMUL(r2, hmm_mexp(u.e - eh)), H_FMUL(r3, hmm_mexp(h.e - eh)));
    u = BF{nu, eu};
    h = BF{nh, eh};
    return true;
}

// The chain's form of the column: exponents assumed to STAY as they are (ue, he; d = ue - he in {-1, 0, 1}), which fixes the
// multipliers of (3) to M(0), m1 = M(-d), m2 = M(d), M(0).  PLAIN: d = 0, no multiplications.  Dependency chain FMUL -> FFMA ->
// (FMUL ->) FADD.  hmm_regime_ok says afterwards whether the assumption and every product held for the column.
template <bool PLAIN>
HMM_HD void hmm_regime_step(float& u, float& h, const float c[4], const float l[4], float m1, float m2)
{
    const float r0 = hmm_fprod(u, c[0], l[0]), r1 = hmm_fprod(h, c[1], l[1]);
    const float r2 = hmm_fprod(u, c[2], l[2]), r3 = hmm_fprod(h, c[3], l[3]);
    if (PLAIN) {
        u = H_FADD(r0, r1);
        h = H_FADD(r2, r3);
    } else {
        u = H_FADD(r0, H_FMUL(r1, m1));
        h = H_FADD(H_FMUL(r2, m2), r3);
    }
}

HMM_HD bool hmm_regime_ok(float u, float h, int ue, int he, const float c[4], const float l[4])
{
    const float r0 = hmm_fprod(u, c[0], l[0]), r1 = hmm_fprod(h, c[1], l[1]);
    const float r2 = hmm_fprod(u, c[2], l[2]), r3 = hmm_fprod(h, c[3], l[3]);
    u32 w0, w1, w2, w3;
    const u32 hz = hmm_prod_check(u, c[0], l[0], r0, w0) | hmm_prod_check(h, c[1], l[1], r1, w1) |
                   hmm_prod_check(u, c[2], l[2], r2, w2) | hmm_prod_check(h, c[3], l[3], r3, w3);
    const int e0 = ue - (int)w0, e1 = he - (int)w1, e2 = ue - (int)w2, e3 = he - (int)w3;
    const int eu = e0 > e1 ? e0 : e1, eh = e2 > e3 ? e2 : e3;
    return hz == 0u && eu == ue && eh == he;
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

// Few, long strings: one warp per (string, direction), a block of 32 columns at a time.  Lane 0 runs the serial recurrence in
// its FP32 regime form (hmm_regime_step: 12 or 16 cycles of dependent latency per column) and parks the state after every column in
// shared memory; then the 32 lanes each re-examine one column from the state it started with (hmm_regime_ok).  The first column that
// did not hold (an exponent moves there, about once in 25 columns, or a product is hazardous) is evaluated by ITS lane with
// hmm_float_step, or hmm_exact_step on a hazard, and the chain resumes behind it.  The other lanes also keep memory out of the
// chain: symbols are fetched one block ahead with one coalesced load, the coefficients of every column are looked up into shared
// memory, and the 32 results leave with one coalesced write.  MAUVE_CUDA_HMM_FP64=1 (tests): every column through hmm_exact_step.
template <bool PLAIN>
__device__ __forceinline__ void hmm_chain_run(const float4* __restrict__ chi, const float4* __restrict__ clo, float2* __restrict__ st, u32 j0, u32 cnt,
                                              float uf, float hf, float m1, float m2)
{
    if (j0 == 0 && cnt == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float4 c = chi[j], l = clo[j];
            const float cc[4] = {c.x, c.y, c.z, c.w}, ll[4] = {l.x, l.y, l.z, l.w};
            hmm_regime_step<PLAIN>(uf, hf, cc, ll, m1, m2);
            st[j + 1] = make_float2(uf, hf);
        }
    } else {
        float4 c = chi[j0], l = clo[j0];
        for (u32 j = j0; j < cnt; ++j) {
            const float4 cn = chi[j + 1], ln = clo[j + 1];   // one column ahead (the arrays have a spare row)
            const float cc[4] = {c.x, c.y, c.z, c.w}, ll[4] = {l.x, l.y, l.z, l.w};
            hmm_regime_step<PLAIN>(uf, hf, cc, ll, m1, m2);
            st[j + 1] = make_float2(uf, hf);
            c = cn;
            l = ln;
        }
    }
}

__global__ void __launch_bounds__(32) hmm_exact_chain_warp_kernel(const u8* __restrict__ sym, const u64* __restrict__ off, u32 n, HmmExactModel m, HmmFastTab ft,
                                                                 int force_exact, BF* __restrict__ fh, BF* __restrict__ bh, BF* __restrict__ total,
                                                                 u32* __restrict__ err, unsigned long long* __restrict__ counters)
{
    __shared__ double te_s[8][4];          // role order
    __shared__ float4 chi_s[8], clo_s[8];  // role order
    __shared__ float4 chi[33], clo[33];    // per column of the block (+ a spare row for the look-ahead)
    __shared__ u8 xs[32];
    __shared__ float2 st[33];              // st[j] = (u, h) before column j of the block, st[j + 1] after it
    __shared__ BF res[32];
    const u32 lane = threadIdx.x;
    const u32 s = blockIdx.x >> 1;
    const bool fwd = (blockIdx.x & 1) == 0;
    const u64 beg = off[s], end = off[s + 1];
    if (end == beg) return;
    {
        const u32 x = lane >> 2, k = lane & 3, src = (fwd || k == 0 || k == 3) ? k : 3 - k;
        te_s[x][k] = m.te[x][src];
        reinterpret_cast<float*>(chi_s)[lane] = ft.hi[x][src];
        reinterpret_cast<float*>(clo_s)[lane] = ft.lo[x][src];
        if (lane == 0) chi[32] = clo[32] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const u64 len = end - beg;
    u32 bad = 0;
    const BF one = BF{1.0f, 0};
    BF h, u;   // every lane carries the state; it moves in lockstep (shared memory / shuffles below)
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
    unsigned long long n_rounds = 0, n_exact = 0;
    for (u64 k0 = 0; k0 < steps; k0 += 32) {
        const u32 cnt = (u32)min((u64)32, steps - k0);
        const u32 x = xn;
        if (k0 + 32 + lane < steps) xn = sym_index(__ldg(sym + sym_of_step(k0 + 32 + lane)), bad);  // next block, in flight during this one
        chi[lane] = chi_s[x];
        clo[lane] = clo_s[x];
        xs[lane] = (u8)x;
        __syncwarp();
        u32 j0 = 0;
        while (j0 < cnt) {
            ++n_rounds;
            const int d = u.e - h.e;
            u32 jbad;   // first column of [j0, cnt) that needs its own evaluation
            if (d >= -1 && d <= 1 && !force_exact) {
                if (lane == 0) {
                    st[j0] = make_float2(u.f, h.f);
                    if (d == 0) hmm_chain_run<true>(chi, clo, st, j0, cnt, u.f, h.f, 1.f, 1.f);
                    else hmm_chain_run<false>(chi, clo, st, j0, cnt, u.f, h.f, hmm_mexp(-d), hmm_mexp(d));
                }
                __syncwarp();
                bool ok = true;
                if (lane >= j0 && lane < cnt) {
                    const float2 v = st[lane];
                    const float4 c = chi[lane], l = clo[lane];
                    const float cc[4] = {c.x, c.y, c.z, c.w}, ll[4] = {l.x, l.y, l.z, l.w};
                    ok = hmm_regime_ok(v.x, v.y, u.e, h.e, cc, ll);
                }
                const u32 badmask = __ballot_sync(0xffffffffu, !ok);
                jbad = badmask ? (u32)__ffs((int)badmask) - 1u : cnt;
                if (lane >= j0 && lane < jbad) res[lane] = BF{st[lane + 1].y, h.e};
                if (jbad > j0) {
                    const float2 v = st[jbad];
                    u.f = v.x;
                    h.f = v.y;
                }
            } else
                jbad = j0;
            if (jbad < cnt) {   // the column's own lane evaluates it from the state in front of it, then everybody takes the result
                if (lane == jbad) {
                    const float4 c = chi[lane], l = clo[lane];
                    const float cc[4] = {c.x, c.y, c.z, c.w}, ll[4] = {l.x, l.y, l.z, l.w};
                    if (force_exact || !hmm_float_step(u, h, cc, ll)) {
                        hmm_exact_step(u, h, te_s[xs[lane]], fwd, m);
                        ++n_exact;
                    }
                    res[lane] = h;
                }
                u.f = __shfl_sync(0xffffffffu, u.f, jbad);
                u.e = __shfl_sync(0xffffffffu, u.e, jbad);
                h.f = __shfl_sync(0xffffffffu, h.f, jbad);
                h.e = __shfl_sync(0xffffffffu, h.e, jbad);
            }
            j0 = jbad + 1;
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
    if (counters) {
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 16);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 8);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 4);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 2);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, 1);
        if (lane == 0) {
            atomicAdd(counters, steps);
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
        MCU_CUDA(cudaMemsetAsync(st.err.as<u32>() + 2, 0, 24, s));
        // few chains: a warp each (latency-optimised); many chains: a thread each (throughput)
        if (2 * n <= (u64)sm_count() * 64)
            hmm_exact_chain_warp_kernel<<<(unsigned)(2 * n), 32, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), (u32)n, xm, ft, force_exact, st.fh.as<BF>(),
                                                                         st.bh.as<BF>(), st.total.as<BF>(), st.err.as<u32>(),
                                                                         reinterpret_cast<unsigned long long*>(st.err.as<u32>() + 2));
        else
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
// emu_hmm_chain: one chain (forward or backward) of one string, evaluated the way hmm_exact_chain_warp_kernel does: the regime step
// while hmm_regime_ok lets it stand, hmm_float_step at the columns where it does not, hmm_exact_step at hazardous ones -- and,
// beside it, hmm_exact_step at EVERY column from the same state (the truth).  out_f / out_e: the homologous-state bfloat after every
// step (steps = n - 1, chain order).  counts[0] = columns carried by the regime step, counts[1] = columns evaluated by hmm_float_step,
// counts[2] = columns where an accepted FP32 form (either) differs from the truth (must be 0), counts[3] = hazardous columns.
extern "C" void emu_hmm_chain(const unsigned char* sym, unsigned long long n, const double* params21, int fwd, float* out_f, int* out_e,
                              unsigned long long* counts)
{
    using namespace mcu;
    HmmExactModel m;
    build_exact_model(params21, &m);
    HmmFastTab ft;
    build_fast_tab(m, &ft);
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
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
        const int d = u.e - h.e;
        BF gu = u, gh = h;
        const bool float_ok = hmm_float_step(gu, gh, ch, cl);
        if (float_ok && !same(gu, gh)) ++counts[2];
        if (d >= -1 && d <= 1 && hmm_regime_ok(u.f, h.f, u.e, h.e, ch, cl)) {
            float uf = u.f, hf = h.f;
            if (d == 0) hmm_regime_step<true>(uf, hf, ch, cl, 1.f, 1.f);
            else hmm_regime_step<false>(uf, hf, ch, cl, hmm_mexp(-d), hmm_mexp(d));
            if (!same(BF{uf, u.e}, BF{hf, h.e})) ++counts[2];
            ++counts[0];
        } else if (float_ok)
            ++counts[1];
        else
            ++counts[3];
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
    unsigned long long c4[4];
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
