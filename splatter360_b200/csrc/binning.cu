// binning.cu -- tile binning: hand-written LSD radix sort + scan + instance emission + tile ranges.
//
// Upstream sorts N (tile<<32 | depth) 64-bit keys with CUB (6 passes over every instance).  Here the
// order (tile, depth, id) is produced in two cheaper steps with identical result:
//   1. sort the P Gaussians once by (depth bits, id)              -- 4 passes over P pairs
//   2. emit instances in that order, then STABLE-sort them by tile -- ceil(log2(tiles)/8) passes over N
// Replaces cub::DeviceScan::InclusiveSum, duplicateWithKeys, cub::DeviceRadixSort::SortPairs and
// identifyTileRanges.
#include "common.cuh"

namespace s360 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per block
constexpr int RS_BINS = 256;

__device__ __forceinline__ int64_t effective_n(int64_t n_cap, const uint32_t* n_dev) {
  if (n_dev == nullptr) return n_cap;
  const int64_t nd = (int64_t)(*n_dev);
  return nd < n_cap ? nd : n_cap;
}

// ---- pass kernel A: per-block digit histogram, written bin-major: hist[bin * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS)
rs_histogram_kernel(const uint32_t* __restrict__ keys, int64_t n_cap, const uint32_t* __restrict__ n_dev,
                    int shift, uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t s_hist[RS_BINS];
  const int64_t n = effective_n(n_cap, n_dev);
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const int64_t j = base + i * RS_THREADS + threadIdx.x;
    uint32_t d = 0x100u | lane;  // unique pseudo digit for out-of-range lanes
    if (j < n) d = (keys[j] >> shift) & 0xffu;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d < 0x100u && lane == (__ffs(peers) - 1)) atomicAdd(&s_hist[d], (uint32_t)__popc(peers));
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

// ---- pass kernel B: one block per digit, exclusive scan over that digit's row of block counts
__global__ void __launch_bounds__(RS_THREADS)
rs_scan_rows_kernel(uint32_t* __restrict__ hist, int nblocks, uint32_t* __restrict__ bin_totals) {
  __shared__ uint32_t s_warp[RS_WARPS];
  __shared__ uint32_t s_carry;
  uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += RS_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t x = i < nblocks ? row[i] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) woff += (w < warp) ? s_warp[w] : 0u;
    const uint32_t carry = s_carry;
    if (i < nblocks) row[i] = carry + woff + incl - x;
    __syncthreads();
    if (threadIdx.x == RS_THREADS - 1) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) bin_totals[blockIdx.x] = s_carry;
}

// ---- pass kernel C: stable rank inside the block, exchange through shared memory, coalesced scatter
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n_cap,
                  const uint32_t* __restrict__ n_dev, int shift, const uint32_t* __restrict__ hist,
                  int nblocks, const uint32_t* __restrict__ bin_totals) {
  __shared__ uint32_t s_cnt[RS_WARPS][RS_BINS];   // per-warp digit counters -> exclusive warp prefixes
  __shared__ uint32_t s_keys[RS_TILE];
  __shared__ uint32_t s_vals[RS_TILE];
  __shared__ int64_t s_gbase[RS_BINS];            // global destination of local sorted slot 0 of each digit
  __shared__ uint32_t s_scan[RS_WARPS];
  __shared__ uint32_t s_scan2[RS_WARPS];
  const int64_t n = effective_n(n_cap, n_dev);
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  if (base >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) s_cnt[w][threadIdx.x] = 0;
  __syncthreads();

  // warp w owns the contiguous items [w*256, w*256+256) of this block, 8 rounds of 32
  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int64_t j = base + warp * (RS_ITEMS * 32) + r * 32 + lane;
    key[r] = 0; val[r] = 0;
    if (j < n) { key[r] = keys_in[j]; val[r] = vals_in[j]; }
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int64_t j = base + warp * (RS_ITEMS * 32) + r * 32 + lane;
    const bool valid = j < n;
    const uint32_t d = valid ? ((key[r] >> shift) & 0xffu) : (0x100u | lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = s_cnt[warp][d];
      s_cnt[warp][d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + (uint32_t)__popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // thread d: exclusive prefix of digit d over the warps, and block total of digit d
  const int d = threadIdx.x;
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) {
    const uint32_t c = s_cnt[w][d];
    s_cnt[w][d] = total;
    total += c;
  }
  // block-exclusive scan of `total` over digits -> first local slot of digit d;
  // block-exclusive scan of bin_totals over digits -> global base of digit d
  const uint32_t bt = bin_totals[d];
  uint32_t incl = total, incl2 = bt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    const uint32_t y2 = __shfl_up_sync(0xffffffffu, incl2, o);
    if (lane >= o) { incl += y; incl2 += y2; }
  }
  if (lane == 31) { s_scan[warp] = incl; s_scan2[warp] = incl2; }
  __syncthreads();
  uint32_t woff = 0, woff2 = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) {
    woff += (w < warp) ? s_scan[w] : 0u;
    woff2 += (w < warp) ? s_scan2[w] : 0u;
  }
  const uint32_t local_first = woff + incl - total;
  const uint32_t global_first = woff2 + incl2 - bt;
  s_gbase[d] = (int64_t)global_first + (int64_t)hist[(size_t)d * nblocks + blockIdx.x] - (int64_t)local_first;
  // s_cnt[w][d] += local_first, so that slot = s_cnt[warp][digit] + rank
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) s_cnt[w][d] += local_first;
  __syncthreads();

#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int64_t j = base + warp * (RS_ITEMS * 32) + r * 32 + lane;
    if (j < n) {
      const uint32_t dg = (key[r] >> shift) & 0xffu;
      const uint32_t slot = s_cnt[warp][dg] + rank[r];
      s_keys[slot] = key[r];
      s_vals[slot] = val[r];
    }
  }
  __syncthreads();
  const int nvalid = (int)((n - base) < (int64_t)RS_TILE ? (n - base) : (int64_t)RS_TILE);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const int slot = i * RS_THREADS + threadIdx.x;
    if (slot < nvalid) {
      const uint32_t k = s_keys[slot];
      const int64_t dst = s_gbase[(k >> shift) & 0xffu] + slot;
      keys_out[dst] = k;
      vals_out[dst] = s_vals[slot];
    }
  }
}

static inline int rs_blocks(int64_t n) { return (int)((n + RS_TILE - 1) / RS_TILE); }

size_t radix_scratch_bytes(int64_t n) {
  const size_t nb = (size_t)rs_blocks(n > 0 ? n : 1);
  return align_up(nb * RS_BINS * sizeof(uint32_t), 256) + align_up(RS_BINS * sizeof(uint32_t), 256);
}

int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n,
                     const uint32_t* n_dev, int nbits, void* scratch, cudaStream_t st, int* result_in_b) {
  *result_in_b = 0;
  if (n <= 0 || nbits <= 0) return 0;
  const int nblocks = rs_blocks(n);
  uint32_t* hist = (uint32_t*)scratch;
  uint32_t* bin_totals = (uint32_t*)((char*)scratch + align_up((size_t)nblocks * RS_BINS * sizeof(uint32_t), 256));
  uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  int flips = 0;
  for (int shift = 0; shift < nbits; shift += 8) {
    rs_histogram_kernel<<<nblocks, RS_THREADS, 0, st>>>(ki, n, n_dev, shift, hist, nblocks);
    rs_scan_rows_kernel<<<RS_BINS, RS_THREADS, 0, st>>>(hist, nblocks, bin_totals);
    rs_scatter_kernel<<<nblocks, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, n_dev, shift, hist, nblocks, bin_totals);
    count_launch(3);
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
    flips++;
  }
  *result_in_b = flips & 1;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of tiles-touched in depth order (K2).  2048 Gaussians per block.
__device__ __forceinline__ uint32_t rect_tiles(uint2 r) { return (r.x >> 16) * (r.y >> 16); }

__global__ void __launch_bounds__(RS_THREADS)
scan_block_sums_kernel(int P, const uint2* __restrict__ rect, const uint32_t* __restrict__ order,
                       uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s_w[RS_WARPS];
  uint32_t s = 0;
  const int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const int j = base + i * RS_THREADS + threadIdx.x;
    if (j < P) s += rect_tiles(rect[order[j]]);
  }
  s = __reduce_add_sync(0xffffffffu, s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) t += s_w[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024)
scan_sums_kernel(uint32_t* __restrict__ block_sums, int nblocks, S360Counters* counters) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t x = i < nblocks ? block_sums[i] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_w[w];
    const uint32_t carry = s_carry;
    if (i < nblocks) block_sums[i] = carry + woff + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { counters->num_rendered = s_carry; counters->overflow = 0; }
}

__global__ void __launch_bounds__(RS_THREADS)
scan_write_offsets_kernel(int P, const uint2* __restrict__ rect, const uint32_t* __restrict__ order,
                          const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ offsets) {
  __shared__ uint32_t s_w[RS_WARPS];
  // thread t owns the 8 consecutive items [base + 8t, base + 8t + 8)
  const int base = blockIdx.x * RS_TILE + threadIdx.x * RS_ITEMS;
  uint32_t c[RS_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const int j = base + i;
    c[i] = j < P ? rect_tiles(rect[order[j]]) : 0u;
    s += c[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) woff += (w < warp) ? s_w[w] : 0u;
  uint32_t run = block_sums[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const int j = base + i;
    if (j < P) offsets[j] = run;
    run += c[i];
  }
}

int launch_scan_offsets(const S360View& v, GeomState g, const uint32_t* depth_order, uint32_t* offsets,
                        S360Counters* counters, uint32_t* block_sums, cudaStream_t st) {
  const int nblocks = rs_blocks(v.P > 0 ? v.P : 1);
  if (v.P > 0) scan_block_sums_kernel<<<nblocks, RS_THREADS, 0, st>>>(v.P, g.rect, depth_order, block_sums);
  scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, v.P > 0 ? nblocks : 0, counters);
  if (v.P > 0) scan_write_offsets_kernel<<<nblocks, RS_THREADS, 0, st>>>(v.P, g.rect, depth_order, block_sums, offsets);
  count_launch(3);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3: emit (tile, gaussian) instances in depth order.  One block owns 2048 consecutive Gaussians of
// the depth order; its threads walk the block's contiguous instance range so that the writes are
// coalesced and the work is balanced no matter how many tiles a single Gaussian covers.
__global__ void __launch_bounds__(RS_THREADS)
emit_instances_kernel(int P, int gx, int mode, const uint2* __restrict__ rect, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ offsets, S360Counters* counters, int64_t capacity,
                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  __shared__ uint32_t s_off[RS_TILE];
  __shared__ uint32_t s_gid[RS_TILE];
  __shared__ uint2 s_rect[RS_TILE];
  const int base = blockIdx.x * RS_TILE;
  const int cnt = min(RS_TILE, P - base);
  for (int i = threadIdx.x; i < cnt; i += RS_THREADS) {
    const uint32_t gid = order[base + i];
    s_gid[i] = gid;
    s_rect[i] = rect[gid];
    s_off[i] = offsets[base + i];
  }
  __syncthreads();
  const uint32_t begin = s_off[0];
  const uint32_t end = s_off[cnt - 1] + rect_tiles(s_rect[cnt - 1]);
  bool overflow = false;
  for (uint32_t j = begin + threadIdx.x; j < end; j += RS_THREADS) {
    // last i with s_off[i] <= j
    int lo = 0, hi = cnt;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] <= j) lo = mid; else hi = mid;
    }
    const uint2 r = s_rect[lo];
    const uint32_t k = j - s_off[lo];
    const uint32_t nx = r.x >> 16;
    const uint32_t ky = k / nx, kx = k - ky * nx;
    int tx = (int)(int16_t)(r.x & 0xffffu) + (int)kx;
    if (mode == S360_MODE_ERP) { tx %= gx; if (tx < 0) tx += gx; }
    const uint32_t ty = (r.y & 0xffffu) + ky;
    if ((int64_t)j < capacity) {
      keys[j] = ty * (uint32_t)gx + (uint32_t)tx;
      vals[j] = s_gid[lo];
    } else {
      overflow = true;
    }
  }
  if (overflow) counters->overflow = 1;
}

int launch_emit(const S360View& v, GeomState g, const uint32_t* depth_order, const uint32_t* offsets,
                S360Counters* counters, int64_t capacity, uint32_t* keys, uint32_t* vals, cudaStream_t st) {
  if (v.P == 0) return 0;
  const int gx = (v.image_width + TILE - 1) / TILE;
  emit_instances_kernel<<<rs_blocks(v.P), RS_THREADS, 0, st>>>(v.P, gx, v.mode, g.rect, depth_order, offsets,
                                                               counters, capacity, keys, vals);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K5: [start, end) of each tile in the sorted instance list
__global__ void tile_ranges_kernel(const uint32_t* __restrict__ keys, const S360Counters* __restrict__ counters,
                                   int64_t capacity, uint2* __restrict__ ranges) {
  int64_t n = counters->num_rendered;
  if (n > capacity) n = capacity;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = keys[i];
  if (i == 0 || keys[i - 1] != t) ranges[t].x = (uint32_t)i;
  if (i == n - 1 || keys[i + 1] != t) ranges[t].y = (uint32_t)(i + 1);
}

int launch_tile_ranges(const S360View& v, const uint32_t* keys, const S360Counters* counters, int64_t capacity,
                       uint2* ranges, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
  cudaMemsetAsync(ranges, 0, (size_t)gx * gy * sizeof(uint2), st);
  if (capacity > 0) {
    tile_ranges_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, st>>>(keys, counters, capacity, ranges);
    count_launch();
  }
  return (int)cudaGetLastError();
}

}  // namespace s360
