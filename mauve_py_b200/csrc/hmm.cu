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

struct HmmState {
    DevBuf sym, off, chunk_first, prod, fin, bin, scratch, pred, post, err;
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
    int r = MCU_OK;
    if ((r = st.sym.reserve(total + 16)) || (r = st.off.reserve((n + 1) * 8)) || (r = st.chunk_first.reserve((n + 1) * 8)) ||
        (r = st.prod.reserve(nchunks * sizeof(double4))) || (r = st.fin.reserve(nchunks * sizeof(double2))) ||
        (r = st.bin.reserve(nchunks * sizeof(double2))) || (r = st.scratch.reserve(grid_c * block * HC * sizeof(double2))) ||
        (r = st.pred.reserve(total + 16)) || (r = st.post.reserve(total * 8 + 16)) || (r = st.err.reserve(16))) {
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
    hmm_products_kernel<<<(unsigned)grid_c, block, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), st.chunk_first.as<u64>(), (u32)n, nchunks, m,
                                                           st.prod.as<double4>(), st.err.as<u32>());
    hmm_chain_kernel<<<(unsigned)div_up(n, block), block, 0, s>>>(st.chunk_first.as<u64>(), (u32)n, m, st.prod.as<double4>(), st.fin.as<double2>(),
                                                                  st.bin.as<double2>());
    hmm_posterior_kernel<<<(unsigned)grid_c, block, 0, s>>>(st.sym.as<u8>(), st.off.as<u64>(), st.chunk_first.as<u64>(), (u32)n, nchunks, m,
                                                            st.fin.as<double2>(), st.bin.as<double2>(), st.scratch.as<double2>(),
                                                            st.pred.as<char>(), post_out ? st.post.as<double>() : nullptr);
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
