// Stable LSD radix sort building block for (uint32 key, uint32 value) pairs, 8-bit digits.
//
// Replaces cub::DeviceRadixSort in the rasterizer's binning stage (the reference sorts R
// 64-bit tile|depth keys in ~6 passes, rasterizer_impl.cu:303-308).  Here the ordering work is
// split so that the R-sized stream is only touched for the tile bits:
//   1. the P Gaussians are sorted by view depth (4 passes over P, P << R);
//   2. instances are emitted in that order (tile id, Gaussian id);
//   3. a stable sort on the tile id alone (2 passes for <= 65536 tiles) groups them per tile while
//      preserving the (depth, index) order inside every tile.
// A sort = one histogram kernel (all digits of all passes in a single read of the keys) + ONE
// kernel per pass: block-local stable ranking with warp match-any votes, block offsets by chained
// (decoupled look-back) prefix over per-block status words, items staged in digit order in shared
// memory and written out in runs.
#pragma once

#include "gsb_common.h"

namespace gsb {
namespace {

constexpr int kRdxThreads = 256;
constexpr int kRdxWarps = kRdxThreads / 32;
constexpr int kRdxItems = 16;                          // items per thread
constexpr int kRdxBlock = kRdxThreads * kRdxItems;     // 4096 items per CTA
constexpr int kRdxBins = 256;
constexpr int kRdxMaxPasses = 4;

// status word of (block, digit): [31:30] 0 = not ready, 1 = block aggregate, 2 = inclusive prefix
constexpr uint32_t kStAgg = 1u << 30, kStIncl = 2u << 30, kStVal = (1u << 30) - 1u;

// Number of items the sort works on: host value, or a device counter clamped by `capacity`
// (0 when the frame overflowed its workspace: counters[2] != 0).
__device__ __forceinline__ uint32_t radix_count(uint32_t n_host, const unsigned long long* counters, int64_t capacity) {
  if (counters == nullptr) return n_host;
  if (counters[2] != 0) return 0u;
  const unsigned long long n = counters[1];
  return (int64_t)n > capacity ? 0u : (uint32_t)n;
}

// Digit histograms of every pass in one read of the keys: ghist[pass][digit].
__global__ void __launch_bounds__(kRdxThreads) radix_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n_host,
                                                                     const unsigned long long* __restrict__ counters,
                                                                     int64_t capacity, int passes,
                                                                     uint32_t* __restrict__ ghist) {
  __shared__ uint32_t hist[kRdxMaxPasses][kRdxBins];
  const uint32_t n = radix_count(n_host, counters, capacity);
  const uint32_t base = blockIdx.x * kRdxBlock;
  if (base >= n) return;
  for (int p = 0; p < passes; ++p) hist[p][threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kRdxItems; ++k) {
    const uint32_t i = base + k * kRdxThreads + threadIdx.x;
    if (i < n) {
      const uint32_t key = keys[i];
      for (int p = 0; p < passes; ++p) atomicAdd(&hist[p][(key >> (8 * p)) & (kRdxBins - 1)], 1u);
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = hist[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * kRdxBins + threadIdx.x], c);
  }
}

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One stable pass.  `ghist` = this pass's 256 global digit counts, `status` = [nblocks][256] words and
// `ticket` = block-order counter, both zeroed before the sort.  When `ranges` != NULL this is the last
// pass of the tile sort: keys are then fully sorted tile ids and the first / last instance of every
// tile seen by the block updates ranges[tile] = (start, end) with atomicMin / atomicMax.
__global__ void __launch_bounds__(kRdxThreads, 3) radix_pass_kernel(const uint32_t* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out, uint32_t n_host,
                                                                const unsigned long long* __restrict__ counters,
                                                                int64_t capacity, int shift,
                                                                const uint32_t* __restrict__ ghist,
                                                                uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                                                                uint2* __restrict__ ranges) {
  __shared__ uint32_t s_key[kRdxBlock];
  __shared__ uint32_t s_val[kRdxBlock];
  __shared__ uint32_t warp_hist[kRdxWarps][kRdxBins];  // per-warp digit counts -> per-warp bases
  __shared__ uint32_t digit_start[kRdxBins];           // first local slot of each digit in this block
  __shared__ uint32_t global_off[kRdxBins];            // global output position of that slot
  __shared__ uint32_t scan_tmp[kRdxWarps];
  __shared__ uint32_t s_bid;

  const uint32_t n = radix_count(n_host, counters, capacity);
  // blocks are numbered in arrival order, so every predecessor of a block is already running
  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
#pragma unroll
  for (int w = 0; w < kRdxWarps; ++w) warp_hist[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t bid = s_bid;
  const uint32_t base = bid * kRdxBlock;
  if (base >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t count = min((uint32_t)kRdxBlock, n - base);

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

  // ---- phase 2: thread d owns digit d
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kRdxWarps; ++w) {
      const uint32_t c = warp_hist[w][d];
      warp_hist[w][d] = run;  // base of warp w inside digit d
      run += c;
    }
    // publish this block's count of digit d, then chain back over the predecessors
    uint32_t* my = status + (size_t)bid * kRdxBins + d;
    st_status(my, (bid == 0 ? kStIncl : kStAgg) | run);
    uint32_t prev = 0;
    if (bid > 0) {
      for (int64_t b = (int64_t)bid - 1; b >= 0; --b) {
        const uint32_t* sp = status + (size_t)b * kRdxBins + d;
        uint32_t sv = ld_status(sp);
        for (uint32_t spin = 0; (sv >> 30) == 0; ++spin) {
          if (spin > (1u << 24)) __trap();  // a predecessor never published: fail loudly instead of hanging
          sv = ld_status(sp);
        }
        prev += sv & kStVal;
        if ((sv >> 30) == 2) break;
      }
      st_status(my, kStIncl | (prev + run));
    }
    // local digit starts: exclusive scan of `run` over the 256 digits
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
    // global digit base: exclusive scan of the global histogram over the 256 digits
    const uint32_t tot = ghist[d];
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
    global_off[d] = tbase + tincl - tot + prev;
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

// Scratch words one sort needs for `max_items` items.
inline size_t radix_scratch_words(size_t max_items) {
  const size_t nblocks = (max_items + kRdxBlock - 1) / kRdxBlock;
  return (size_t)kRdxMaxPasses * (kRdxBins + nblocks * kRdxBins + 1) + 64;
}

// Scratch layout: ghist [kRdxMaxPasses][256] | tickets [64] | status [passes][nblocks][256]
inline uint32_t* radix_ghist(uint32_t* scratch) { return scratch; }

// Zeroes the histogram / ticket / status words one sort over `max_items` items with `bits` key bits needs.
inline void radix_prepare(uint32_t* scratch, size_t max_items, int bits, cudaStream_t stream) {
  const size_t nblocks = (max_items + kRdxBlock - 1) / kRdxBlock;
  const int passes = (bits + 7) / 8;
  const size_t words = (size_t)kRdxMaxPasses * kRdxBins + 64 + (size_t)passes * nblocks * kRdxBins;
  cudaMemsetAsync(scratch, 0, words * sizeof(uint32_t), stream);
}

// Enqueues ceil(bits/8) stable passes sorting (keys, vals) by key bits [0, bits).  `a`/`b` are
// ping-pong buffers; returns which buffer holds the result (0 = a, 1 = b).  n is either the host
// value (counters == NULL) or read on the device from counters[1] (clamped by capacity).
// histogram_ready: the caller already ran radix_prepare() and filled the digit histograms.
inline int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n_host,
                            const unsigned long long* counters, int64_t capacity, size_t max_items, int bits,
                            uint32_t* scratch, uint2* ranges_on_last_pass, cudaStream_t stream, uint64_t* launches,
                            bool histogram_ready) {
  const uint32_t nblocks = (uint32_t)((max_items + kRdxBlock - 1) / kRdxBlock);
  if (nblocks == 0) return 0;
  const int passes = (bits + 7) / 8;
  uint32_t* ghist = scratch;                                    // [kRdxMaxPasses][256]
  uint32_t* tickets = ghist + kRdxMaxPasses * kRdxBins;          // [64]
  uint32_t* status = tickets + 64;                               // [passes][nblocks][256]
  if (!histogram_ready) {
    radix_prepare(scratch, max_items, bits, stream);
    radix_histogram_kernel<<<nblocks, kRdxThreads, 0, stream>>>(keys_a, n_host, counters, capacity, passes, ghist);
    *launches += 1;
  }
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    const uint32_t* kin = cur ? keys_b : keys_a;
    const uint32_t* vin = cur ? vals_b : vals_a;
    uint32_t* kout = cur ? keys_a : keys_b;
    uint32_t* vout = cur ? vals_a : vals_b;
    radix_pass_kernel<<<nblocks, kRdxThreads, 0, stream>>>(kin, vin, kout, vout, n_host, counters, capacity, 8 * p,
                                                           ghist + p * kRdxBins, status + (size_t)p * nblocks * kRdxBins,
                                                           tickets + p, p == passes - 1 ? ranges_on_last_pass : nullptr);
    *launches += 1;
    cur ^= 1;
  }
  return cur;
}

}  // namespace
}  // namespace gsb
