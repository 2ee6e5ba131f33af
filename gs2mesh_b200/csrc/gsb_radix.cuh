// Stable LSD radix sort building block for (uint32 key, uint32 value) pairs, 8-bit digits.
//
// Replaces cub::DeviceRadixSort in the rasterizer's binning stage (the reference sorts R
// 64-bit tile|depth keys in ~6 passes, rasterizer_impl.cu:303-308).  Here the ordering work is
// split so that the R-sized stream is only touched for the tile bits:
//   1. the P Gaussians are sorted by view depth (4 passes over P, P << R);
//   2. instances are emitted in that order (tile id, Gaussian id);
//   3. a stable sort on the tile id alone (2 passes for <= 65536 tiles) groups them per tile while
//      preserving the (depth, index) order inside every tile.
// One pass = three kernels: per-block digit histograms -> row scan -> stable scatter.  The scatter
// ranks items with warp match-any votes (stable inside a warp's contiguous segment), combines
// the per-warp counts, stages the block's items in digit order in shared memory and writes runs.
#pragma once

#include "gsb_common.h"

namespace gsb {
namespace {

constexpr int kRdxThreads = 256;
constexpr int kRdxWarps = kRdxThreads / 32;
constexpr int kRdxItems = 16;                          // items per thread
constexpr int kRdxBlock = kRdxThreads * kRdxItems;     // 4096 items per CTA
constexpr int kRdxBins = 256;

// Number of items the pass works on: host value, or a device counter clamped by `capacity`
// (0 when the frame overflowed its workspace: counters[2] != 0).
__device__ __forceinline__ uint32_t radix_count(uint32_t n_host, const unsigned long long* counters, int64_t capacity) {
  if (counters == nullptr) return n_host;
  if (counters[2] != 0) return 0u;
  const unsigned long long n = counters[1];
  return (int64_t)n > capacity ? 0u : (uint32_t)n;
}

__global__ void __launch_bounds__(kRdxThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n_host,
                                                                const unsigned long long* __restrict__ counters,
                                                                int64_t capacity, int shift, uint32_t nblocks,
                                                                uint32_t* __restrict__ table) {
  __shared__ uint32_t hist[kRdxBins];
  const uint32_t n = radix_count(n_host, counters, capacity);
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kRdxBlock;
  if (base < n) {
#pragma unroll
    for (int k = 0; k < kRdxItems; ++k) {
      const uint32_t i = base + k * kRdxThreads + threadIdx.x;
      if (i < n) atomicAdd(&hist[(keys[i] >> shift) & (kRdxBins - 1)], 1u);
    }
  }
  __syncthreads();
  table[threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

// One CTA per digit: exclusive scan of the digit's per-block counts (in place), row total out.
__global__ void __launch_bounds__(kRdxThreads) radix_scan_rows_kernel(uint32_t* __restrict__ table, uint32_t nblocks,
                                                                     uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sums[kRdxWarps];
  __shared__ uint32_t carry;
  uint32_t* row = table + (size_t)blockIdx.x * nblocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nblocks; b0 += kRdxThreads) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t c = i < nblocks ? row[i] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kRdxWarps; ++w)
      if (w < warp) wbase += warp_sums[w];
    const uint32_t excl = carry + wbase + incl - c;
    if (i < nblocks) row[i] = excl;
    __syncthreads();
    if (threadIdx.x == kRdxThreads - 1) carry = excl + c;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// Stable scatter of one pass.  When `ranges` != NULL this is the last pass of the tile sort:
// keys are then fully sorted tile ids and the first / last instance of every tile seen by the
// block updates ranges[tile] = (start, end) with atomicMin / atomicMax.
__global__ void __launch_bounds__(kRdxThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                   const uint32_t* __restrict__ vals_in,
                                                                   uint32_t* __restrict__ keys_out,
                                                                   uint32_t* __restrict__ vals_out, uint32_t n_host,
                                                                   const unsigned long long* __restrict__ counters,
                                                                   int64_t capacity, int shift, uint32_t nblocks,
                                                                   const uint32_t* __restrict__ table,
                                                                   const uint32_t* __restrict__ totals,
                                                                   uint2* __restrict__ ranges) {
  __shared__ uint32_t s_key[kRdxBlock];
  __shared__ uint32_t s_val[kRdxBlock];
  __shared__ uint32_t warp_hist[kRdxWarps][kRdxBins];  // per-warp digit counts -> per-warp bases
  __shared__ uint32_t digit_start[kRdxBins];           // first local slot of each digit in this block
  __shared__ uint32_t global_off[kRdxBins];            // global output position of that slot
  __shared__ uint32_t scan_tmp[kRdxWarps];

  const uint32_t n = radix_count(n_host, counters, capacity);
  const uint32_t base = blockIdx.x * kRdxBlock;
  if (base >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t count = min((uint32_t)kRdxBlock, n - base);

#pragma unroll
  for (int w = 0; w < kRdxWarps; ++w) warp_hist[w][threadIdx.x] = 0;
  __syncthreads();

  // ---- phase 1: each warp walks its contiguous 512-item segment in order and ranks its items
  uint32_t key[kRdxItems], val[kRdxItems];
  uint32_t lrank[kRdxItems];  // rank among same-digit items of this warp's segment
  const uint32_t seg = base + warp * (kRdxBlock / kRdxWarps);
#pragma unroll
  for (int k = 0; k < kRdxItems; ++k) {
    const uint32_t i = seg + k * 32 + lane;
    const bool ok = i < n;
    key[k] = ok ? keys_in[i] : 0xffffffffu;
    val[k] = ok ? vals_in[i] : 0u;
    const uint32_t d = ok ? ((key[k] >> shift) & (kRdxBins - 1)) : (uint32_t)kRdxBins;  // 256 = "no item"
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t before = 0;
    if (ok) before = warp_hist[warp][d];
    __syncwarp();
    if (ok && (peers & lt_mask) == 0) warp_hist[warp][d] = before + __popc(peers);  // lowest lane of the group
    __syncwarp();
    lrank[k] = before + __popc(peers & lt_mask);
  }
  __syncthreads();

  // ---- phase 2: thread d owns digit d: per-warp bases, block digit starts, global offsets
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kRdxWarps; ++w) {
      const uint32_t c = warp_hist[w][d];
      warp_hist[w][d] = run;  // base of warp w inside digit d
      run += c;
    }
    // exclusive scan of `run` (digit totals of this block) over the 256 digits
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) scan_tmp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kRdxWarps; ++w)
      if (w < warp) wbase += scan_tmp[w];
    digit_start[d] = wbase + incl - run;
    __syncthreads();
    // global digit base = exclusive scan of the totals over digits, + this block's row-scanned count
    const uint32_t tot = totals[d];
    uint32_t tincl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, tincl, o);
      if (lane >= o) tincl += v;
    }
    if (lane == 31) scan_tmp[warp] = tincl;
    __syncthreads();
    uint32_t tbase = 0;
#pragma unroll
    for (int w = 0; w < kRdxWarps; ++w)
      if (w < warp) tbase += scan_tmp[w];
    global_off[d] = tbase + tincl - tot + table[(size_t)d * nblocks + blockIdx.x];
  }
  __syncthreads();

  // ---- phase 3: stage the block's items in digit order
#pragma unroll
  for (int k = 0; k < kRdxItems; ++k) {
    const uint32_t i = seg + k * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[k] >> shift) & (kRdxBins - 1);
      const uint32_t slot = digit_start[d] + warp_hist[warp][d] + lrank[k];
      s_key[slot] = key[k];
      s_val[slot] = val[k];
    }
  }
  __syncthreads();

  // ---- phase 4: write runs (consecutive slots of one digit go to consecutive addresses)
  for (uint32_t sidx = threadIdx.x; sidx < count; sidx += kRdxThreads) {
    const uint32_t kk = s_key[sidx];
    const uint32_t d = (kk >> shift) & (kRdxBins - 1);
    const uint32_t dst = global_off[d] + (sidx - digit_start[d]);
    keys_out[dst] = kk;
    vals_out[dst] = s_val[sidx];
    if (ranges != nullptr) {
      // slots are in ascending tile order inside the block (stable LSD => low digit sorted within high digit)
      if (sidx == 0 || s_key[sidx - 1] != kk) atomicMin(&ranges[kk].x, dst);
      if (sidx + 1 == count || s_key[sidx + 1] != kk) atomicMax(&ranges[kk].y, dst + 1);
    }
  }
}

struct RadixScratch {
  uint32_t* table;   // [256 * nblocks]
  uint32_t* totals;  // [256]
};

// Enqueues ceil(bits/8) stable passes sorting (keys, vals) by key bits [0, bits).  `a`/`b` are
// ping-pong buffers; returns which buffer holds the result (0 = a, 1 = b).  n is either the host
// value (counters == NULL) or read on the device from counters[1] (clamped by capacity).
inline int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n_host,
                            const unsigned long long* counters, int64_t capacity, size_t max_items, int bits,
                            const RadixScratch& sc, uint2* ranges_on_last_pass, cudaStream_t stream, uint64_t* launches) {
  const uint32_t nblocks = (uint32_t)((max_items + kRdxBlock - 1) / kRdxBlock);
  if (nblocks == 0) return 0;
  const int passes = (bits + 7) / 8;
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    const uint32_t* kin = cur ? keys_b : keys_a;
    const uint32_t* vin = cur ? vals_b : vals_a;
    uint32_t* kout = cur ? keys_a : keys_b;
    uint32_t* vout = cur ? vals_a : vals_b;
    radix_hist_kernel<<<nblocks, kRdxThreads, 0, stream>>>(kin, n_host, counters, capacity, 8 * p, nblocks, sc.table);
    radix_scan_rows_kernel<<<kRdxBins, kRdxThreads, 0, stream>>>(sc.table, nblocks, sc.totals);
    radix_scatter_kernel<<<nblocks, kRdxThreads, 0, stream>>>(kin, vin, kout, vout, n_host, counters, capacity, 8 * p, nblocks,
                                                              sc.table, sc.totals, p == passes - 1 ? ranges_on_last_pass : nullptr);
    *launches += 3;
    cur ^= 1;
  }
  return cur;
}

}  // namespace
}  // namespace gsb
