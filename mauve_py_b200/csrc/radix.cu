// Host driver + small kernels of the onesweep radix sort (see radix.cuh).
#include "radix.cuh"

namespace mcu {

__global__ void rs_scan_hist_kernel(u64* hist)
{
    __shared__ u64 warp_tot[8];
    u64* h = hist + (size_t)blockIdx.x * RS_RADIX;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u64 v = h[tid], incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    u64 add = 0;
    for (u32 w = 0; w < warp; ++w) add += warp_tot[w];
    h[tid] = incl - v + add;
}

int radix_clear_hist(RadixScratch& sc, cudaStream_t stream)
{
    MCU_TRY(sc.hist.reserve(8 * RS_RADIX * sizeof(u64)));
    MCU_CUDA(cudaMemsetAsync(sc.hist.p, 0, 8 * RS_RADIX * sizeof(u64), stream));
    return MCU_OK;
}

template <typename K>
int radix_sort_pairs(RadixScratch& sc, K* keys_a, u32* vals_a, K* keys_b, u32* vals_b, u64 n, int bits, bool hist_ready,
                     cudaStream_t stream, bool* out_in_a, int* passes_out)
{
    *out_in_a = true;
    if (passes_out) *passes_out = 0;
    if (n == 0 || bits <= 0) return MCU_OK;
    if (bits > 8 * (int)sizeof(K)) bits = 8 * (int)sizeof(K);
    const int passes = (bits + 7) / 8;
    if (passes > 8) { set_error("radix_sort_pairs: too many passes"); return MCU_EINVAL; }
    constexpr int TILE = RsCfg<K>::TILE;
    const u64 tiles = div_up(n, TILE);

    static bool attr_set[2] = {false, false};
    const int ti = sizeof(K) == 4 ? 0 : 1;
    if (!attr_set[ti]) {
        MCU_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RsCfg<K>::SMEM));
        attr_set[ti] = true;
    }

    if (!hist_ready) {
        MCU_TRY(radix_clear_hist(sc, stream));
        int grid = (int)(tiles < (u64)(sm_count() * 8) ? tiles : (u64)(sm_count() * 8));
        if (grid < 1) grid = 1;
        rs_histogram_kernel<K><<<grid, 256, passes * RS_RADIX * sizeof(u32), stream>>>(keys_a, n, passes, 0, sc.hist.as<u64>());
        sc.launches++;
    }
    rs_scan_hist_kernel<<<passes, RS_RADIX, 0, stream>>>(sc.hist.as<u64>());
    sc.launches++;

    const size_t status_bytes = (size_t)tiles * RS_RADIX * sizeof(u64);
    if (status_bytes > sc.status.cap) {
        MCU_TRY(sc.status.reserve(status_bytes));
        MCU_CUDA(cudaMemsetAsync(sc.status.p, 0, sc.status.cap, stream));
        sc.epoch = 0;
    }
    if (sc.epoch + (u32)passes >= RS_EPOCH_MAX) {
        MCU_CUDA(cudaMemsetAsync(sc.status.p, 0, sc.status.cap, stream));
        sc.epoch = 0;
    }
    MCU_TRY(sc.counters.reserve(16 * sizeof(u32)));
    MCU_CUDA(cudaMemsetAsync(sc.counters.p, 0, 16 * sizeof(u32), stream));

    K* kin = keys_a; u32* vin = vals_a; K* kout = keys_b; u32* vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        ++sc.epoch;
        rs_onesweep_kernel<K><<<(unsigned)tiles, RS_BLOCK, RsCfg<K>::SMEM, stream>>>(
            kin, vin, kout, vout, n, 8 * p, sc.hist.as<u64>() + (size_t)p * RS_RADIX, (volatile u64*)sc.status.p,
            sc.counters.as<u32>() + p, sc.epoch);
        sc.launches++;
        K* tk = kin; kin = kout; kout = tk;
        u32* tv = vin; vin = vout; vout = tv;
    }
    MCU_CUDA(cudaGetLastError());
    *out_in_a = (kin == keys_a);
    if (passes_out) *passes_out = passes;
    return MCU_OK;
}

template int radix_sort_pairs<u32>(RadixScratch&, u32*, u32*, u32*, u32*, u64, int, bool, cudaStream_t, bool*, int*);
template int radix_sort_pairs<u64>(RadixScratch&, u64*, u32*, u64*, u32*, u64, int, bool, cudaStream_t, bool*, int*);

}  // namespace mcu
