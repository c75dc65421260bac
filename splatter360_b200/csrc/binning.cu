// binning.cu -- tile binning: hand-written onesweep radix sort + scan + instance emission + tile ranges.
//
// Upstream sorts N (tile<<32 | depth) 64-bit keys with CUB (6 passes over every instance).  Here the
// order (tile, depth, id) is produced in two cheaper steps with identical result:
//   1. sort the P Gaussians once by (depth bits, id)              -- 4 passes over P pairs
//   2. bucket the instances by tile in that order, stably: "matrix binning" (count matrix [chunk][tile] -> column scan ->
//      ranked scatter, further down) whenever the tile count allows it; otherwise emit (tile, id) and STABLE-sort by tile
//      -- ceil(log2(tiles)/8) passes over N
// Each pass is ONE kernel (onesweep: per-block digit counts are published to a status array and the
// exclusive prefix over preceding blocks is obtained by decoupled look-back), the digit histograms of
// all passes come from a single read of the keys (depth) or from the per-tile instance histogram that
// the emission kernel accumulates anyway (tiles); the tile ranges are a scan of that same histogram.
// Replaces cub::DeviceScan::InclusiveSum, duplicateWithKeys, cub::DeviceRadixSort::SortPairs and
// identifyTileRanges.
#include "common.cuh"

namespace s360 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef S360_RS_MINB
#define S360_RS_MINB 3
#endif
// keys per thread of a onesweep pass: 12 (3072 per CTA) in general; 10 (2560 per CTA) while all CTAs of a pass are still
// resident at once (148 SMs x 3 CTAs): more, shorter CTAs finish a latency-bound pass sooner (1M keys: 18.6 -> 17.5 us per
// pass; with 3M keys -- several waves either way -- the smaller tile is 2 % slower).  -DS360_RS_ITEMS_FORCE=10 or 12 pins one of the two (A/B).
constexpr int RS_ITEMS_SMALL = 10, RS_ITEMS_LARGE = 12;
constexpr int64_t RS_SMALL_MAX_KEYS = (int64_t)148 * 3 * RS_THREADS * RS_ITEMS_SMALL;
static inline int rs_items(int64_t n) {
#ifdef S360_RS_ITEMS_FORCE
  (void)n; return S360_RS_ITEMS_FORCE;
#else
  return n <= RS_SMALL_MAX_KEYS ? RS_ITEMS_SMALL : RS_ITEMS_LARGE;
#endif
}
constexpr int RS_BINS = 256;
constexpr uint32_t ST_AGG = 1u << 30, ST_PREFIX = 2u << 30, ST_MASK = (1u << 30) - 1u;

__device__ __forceinline__ int64_t effective_n(int64_t n_cap, const uint32_t* n_dev) {
  if (n_dev == nullptr) return n_cap;
  const int64_t nd = (int64_t)(*n_dev);
  return nd < n_cap ? nd : n_cap;
}

// ---- digit histograms of up to four 8-bit passes from one read of the keys --------------------
__global__ void __launch_bounds__(RS_THREADS)
rs_global_hist_kernel(const uint32_t* __restrict__ keys, int64_t n_cap, const uint32_t* __restrict__ n_dev, int npasses,
                      uint32_t* __restrict__ hist) {
  pdl_enter();
  __shared__ uint32_t s_hist[4][RS_BINS];
  const int64_t n = effective_n(n_cap, n_dev);
  for (int i = threadIdx.x; i < 4 * RS_BINS; i += RS_THREADS) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  // four keys per load (the key buffers are 256-byte aligned); the tail is handled key by key
  const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
  const int64_t n4 = n >> 2;
  const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
  for (int64_t j = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; j < n4; j += stride) {
    const uint4 k = keys4[j];
    for (int p = 0; p < npasses; p++) {
      atomicAdd(&s_hist[p][(k.x >> (8 * p)) & 0xffu], 1u); atomicAdd(&s_hist[p][(k.y >> (8 * p)) & 0xffu], 1u);
      atomicAdd(&s_hist[p][(k.z >> (8 * p)) & 0xffu], 1u); atomicAdd(&s_hist[p][(k.w >> (8 * p)) & 0xffu], 1u);
    }
  }
  for (int64_t j = (n4 << 2) + (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; j < n; j += stride) {
    const uint32_t k = keys[j];
    for (int p = 0; p < npasses; p++) atomicAdd(&s_hist[p][(k >> (8 * p)) & 0xffu], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npasses * RS_BINS; i += RS_THREADS) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// ---- one radix pass: rank in block, look back for the cross-block prefix, exchange, scatter ---
// hist: this pass's global digit counts [256].  status: [nblocks][256], zero on entry.
// counter: zero on entry; hands out block ids in scheduling order so that look-back cannot deadlock.
template <int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, S360_RS_MINB)
rs_onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n_cap,
                   const uint32_t* __restrict__ n_dev, int shift, const uint32_t* __restrict__ hist,
                   uint32_t* status, uint32_t* counter) {
  pdl_enter();
  constexpr int TILE = RS_THREADS * ITEMS;
  __shared__ uint32_t s_cnt[RS_WARPS][RS_BINS];   // per-warp digit counters -> exclusive warp prefixes
  __shared__ uint32_t s_keys[TILE];
  __shared__ uint32_t s_vals[TILE];
  __shared__ int64_t s_gbase[RS_BINS];            // global destination of local sorted slot 0 of each digit
  __shared__ uint32_t s_scan[RS_WARPS];
  __shared__ uint32_t s_scan2[RS_WARPS];
  __shared__ uint32_t s_bid;
  const int64_t n = effective_n(n_cap, n_dev);
  if (threadIdx.x == 0) s_bid = atomicAdd(counter, 1u);
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) s_cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t bid = s_bid;
  const int64_t base = (int64_t)bid * TILE;
  if (base >= n) return;   // every later block is out of range too: nobody will look back at this one
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  // warp w owns the contiguous items [w * 32 * ITEMS, (w + 1) * 32 * ITEMS) of this block: ITEMS rounds of 32
  uint32_t key[ITEMS], val[ITEMS], rank[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int64_t j = base + warp * (ITEMS * 32) + r * 32 + lane;
    key[r] = 0; val[r] = 0;
    if (j < n) { key[r] = keys_in[j]; val[r] = vals_in[j]; }
  }
#ifndef S360_RS_MATCH
  // lanes with the same digit, from eight ballots (one per digit bit) instead of MATCH.ANY: the match instruction's
  // latency grows with the number of distinct values in the warp (~30 of 256 digits here) and every round waits for it;
  // the ballots of all rounds are independent and are issued back to back before the serial counter updates
  unsigned peer_mask[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int64_t j = base + warp * (ITEMS * 32) + r * 32 + lane;
    const uint32_t d = (key[r] >> shift) & 0xffu;
    unsigned m = __ballot_sync(0xffffffffu, j < n);
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const bool bit = (d >> b) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, bit);
      m &= bit ? bal : ~bal;
    }
    peer_mask[r] = m;
  }
#endif
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int64_t j = base + warp * (ITEMS * 32) + r * 32 + lane;
    const bool valid = j < n;
    const uint32_t d = valid ? ((key[r] >> shift) & 0xffu) : (0x100u | lane);
#ifdef S360_RS_MATCH
    const unsigned peers = __match_any_sync(0xffffffffu, d);
#else
    const unsigned peers = valid ? peer_mask[r] : (1u << lane);
#endif
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
  // publish this block's count of digit d, then look back over the preceding blocks
  uint32_t* my_status = status + (size_t)bid * RS_BINS + d;
  volatile uint32_t* vstatus = status;
  if (bid == 0) {
    *reinterpret_cast<volatile uint32_t*>(my_status) = total | ST_PREFIX;
  } else {
    *reinterpret_cast<volatile uint32_t*>(my_status) = total | ST_AGG;
  }
  // Look-back over a WINDOW of predecessors per step: the loads of a window are independent and in flight together.  A
  // sort of a million keys is a single wave of CTAs that all publish their counts at about the same time, so the last
  // CTA has to walk over ~half the grid before it meets an inclusive prefix; one predecessor per L2 round trip made that
  // walk (not the ranking or the scatter) the critical path of every pass.
  uint32_t excl = 0;
  if (bid > 0) {
    constexpr int LB = 8;
    int64_t j = (int64_t)bid - 1;
    bool done = false;
    while (!done) {
      uint32_t sv[LB];
#pragma unroll
      for (int i = 0; i < LB; i++) {   // entries before block 0 read as an empty inclusive prefix
        sv[i] = 2u << 30;
        if (j - i >= 0) sv[i] = vstatus[(size_t)(j - i) * RS_BINS + d];
      }
      int used = 0;
#pragma unroll
      for (int i = 0; i < LB; i++) {
        const uint32_t flag = sv[i] & ~ST_MASK;
        if (done || used < i || flag == 0u) continue;   // stop at the first predecessor that has not published yet
        excl += sv[i] & ST_MASK;
        used = i + 1;
        if (flag == ST_PREFIX) done = true;
      }
      j -= used;   // used == 0: spin on the same window
    }
    *reinterpret_cast<volatile uint32_t*>(my_status) = (excl + total) | ST_PREFIX;
  }
  // block-exclusive scan of `total` over digits -> first local slot of digit d;
  // block-exclusive scan of the global histogram over digits -> global base of digit d
  const uint32_t bt = hist[d];
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
  s_gbase[d] = (int64_t)global_first + (int64_t)excl - (int64_t)local_first;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) s_cnt[w][d] += local_first;
  __syncthreads();

#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int64_t j = base + warp * (ITEMS * 32) + r * 32 + lane;
    if (j < n) {
      const uint32_t dg = (key[r] >> shift) & 0xffu;
      const uint32_t slot = s_cnt[warp][dg] + rank[r];
      s_keys[slot] = key[r];
      s_vals[slot] = val[r];
    }
  }
  __syncthreads();
  const int nvalid = (int)((n - base) < (int64_t)TILE ? (n - base) : (int64_t)TILE);
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int slot = i * RS_THREADS + threadIdx.x;
    if (slot < nvalid) {
      const uint32_t k = s_keys[slot];
      const int64_t dst = s_gbase[(k >> shift) & 0xffu] + slot;
      keys_out[dst] = k;
      vals_out[dst] = s_vals[slot];
    }
  }
}

static inline int rs_blocks(int64_t n) { const int tile = RS_THREADS * rs_items(n); return (int)((n + tile - 1) / tile); }

// scratch: [hist 4*256][counters 4 (padded)][status passes*nblocks*256]
static inline size_t rs_status_words(int64_t n, int passes) { return (size_t)passes * rs_blocks(n > 0 ? n : 1) * RS_BINS; }
size_t radix_scratch_bytes(int64_t n) {
  return align_up(4 * RS_BINS * 4, 256) + 256 + align_up(rs_status_words(n, 4) * 4, 256);
}

// hist_ready: the caller already filled hist[pass][256] (and the scratch was zeroed before that)
int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n,
                     const uint32_t* n_dev, int nbits, void* scratch, cudaStream_t st, int* result_in_b,
                     bool hist_ready, uint32_t* vals_final) {
  *result_in_b = 0;
  if (n <= 0 || nbits <= 0) return 0;
  const int passes = (nbits + 7) / 8;
  const int nblocks = rs_blocks(n);
  uint32_t* hist = (uint32_t*)scratch;
  uint32_t* counters = (uint32_t*)((char*)scratch + align_up(4 * RS_BINS * 4, 256));
  uint32_t* status = (uint32_t*)((char*)counters + 256);
  if (!hist_ready) {
    cudaMemsetAsync(scratch, 0, align_up(4 * RS_BINS * 4, 256) + 256 + rs_status_words(n, passes) * 4, st);
    const int hb = nblocks < 592 ? nblocks : 592;
    rs_global_hist_kernel<<<hb, RS_THREADS, 0, st>>>(keys_a, n, n_dev, passes, hist);
    count_launch();
  }
  uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  for (int p = 0; p < passes; p++) {
    uint32_t* vdst = (vals_final != nullptr && p == passes - 1) ? vals_final : vo;   // last pass may write elsewhere
    if (rs_items(n) == RS_ITEMS_SMALL)
      launch_pdl(rs_onesweep_kernel<RS_ITEMS_SMALL>, dim3(nblocks), dim3(RS_THREADS), 0, st, ki, vi, ko, vdst, n, n_dev, 8 * p,
                 hist + p * RS_BINS, status + (size_t)p * nblocks * RS_BINS, counters + p);
    else
      launch_pdl(rs_onesweep_kernel<RS_ITEMS_LARGE>, dim3(nblocks), dim3(RS_THREADS), 0, st, ki, vi, ko, vdst, n, n_dev, 8 * p,
                 hist + p * RS_BINS, status + (size_t)p * nblocks * RS_BINS, counters + p);
    count_launch();
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  *result_in_b = passes & 1;
  return (int)cudaGetLastError();
}

// zero the radix scratch and return the histogram pointer so that a producer kernel can fill it
uint32_t* radix_prepare_hist(void* scratch, int64_t n, int nbits, cudaStream_t st) {
  const int passes = (nbits + 7) / 8;
  cudaMemsetAsync(scratch, 0, align_up(4 * RS_BINS * 4, 256) + 256 + rs_status_words(n, passes > 0 ? passes : 1) * 4, st);
  return (uint32_t*)scratch;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of tiles-touched in depth order (K2).  2048 Gaussians per block.
__device__ __forceinline__ uint32_t rect_tiles(uint2 r) { return (r.x >> 16) * (r.y >> 16); }

constexpr int SC_ITEMS = 8, SC_TILE = RS_THREADS * SC_ITEMS;

// Single-pass chained scan (decoupled look-back): block b publishes its sum, adds the sums of its predecessors
// until it meets an inclusive prefix, and writes the exclusive offsets of its 2048 Gaussians.
// status: [nblocks] zero on entry; counter: zero on entry.
__global__ void __launch_bounds__(RS_THREADS)
scan_offsets_kernel(int P_cap, const uint32_t* __restrict__ n_dev, const uint2* __restrict__ rect,
                    const uint32_t* __restrict__ order, uint32_t* __restrict__ offsets, uint32_t* status,
                    uint32_t* counter, S360Counters* counters) {
  __shared__ uint32_t s_w[RS_WARPS];
  __shared__ uint32_t s_bid, s_excl;
  const int P = (int)effective_n(P_cap, n_dev);
  if (threadIdx.x == 0) s_bid = atomicAdd(counter, 1u);
  __syncthreads();
  const uint32_t bid = s_bid;
  const int base = (int)bid * SC_TILE + threadIdx.x * SC_ITEMS;   // thread t owns 8 consecutive items
  uint32_t c[SC_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) {
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
  uint32_t woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) { woff += (w < warp) ? s_w[w] : 0u; total += s_w[w]; }
  if (warp == 0) {
    // warp-wide look-back: 32 predecessors are inspected per step
    volatile uint32_t* vs = status;
    uint32_t excl = 0;
    if (bid == 0) {
      if (lane == 0) vs[0] = total | ST_PREFIX;
    } else {
      if (lane == 0) vs[bid] = total | ST_AGG;
      int64_t hi = (int64_t)bid - 1;   // nearest predecessor not yet accounted for
      while (true) {
        const int64_t j = hi - lane;
        uint32_t sv = ST_PREFIX;        // lanes before block 0 behave like an empty inclusive prefix
        if (j >= 0) sv = vs[j];
        const uint32_t flag = sv & ~ST_MASK;
        const unsigned notready = __ballot_sync(0xffffffffu, flag == 0u);
        const unsigned isprefix = __ballot_sync(0xffffffffu, flag == ST_PREFIX);
        // usable lanes: those nearer than the first not-ready lane, up to and including the first prefix
        const int first_nr = notready ? __ffs(notready) - 1 : 32;
        const int first_px = isprefix ? __ffs(isprefix) - 1 : 32;
        const int upto = first_px < first_nr ? first_px + 1 : first_nr;   // number of lanes to add
        const uint32_t add = (lane < upto) ? (sv & ST_MASK) : 0u;
        excl += __reduce_add_sync(0xffffffffu, add);
        if (first_px < first_nr) break;
        hi -= upto;                     // upto == 0 just spins on the same window
      }
      if (lane == 0) vs[bid] = (excl + total) | ST_PREFIX;
    }
    if (lane == 0) {
      s_excl = excl;
      // the block holding the last item publishes the total (block 0 when there is nothing); bit 1 of overflow
      // (pair buffers of the batched path) is not this stage's to clear
      if ((int)bid == max((P + SC_TILE - 1) / SC_TILE - 1, 0)) { counters->num_rendered = excl + total; counters->overflow &= 2u; }
    }
  }
  __syncthreads();
  uint32_t run = s_excl + woff + incl - s;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) {
    const int j = base + i;
    if (j < P) offsets[j] = run;
    run += c[i];
  }
}

__global__ void scan_empty_kernel(S360Counters* counters) { counters->num_rendered = 0; counters->overflow &= 2u; }

int launch_scan_offsets(int64_t n_items, const uint32_t* n_dev, GeomState g, const uint32_t* depth_order,
                        uint32_t* offsets, S360Counters* counters, uint32_t* block_sums, cudaStream_t st) {
  // block_sums scratch: [counter][status nblocks]; instance totals stay below 2^30 (status word payload)
  const int nblocks = (int)((n_items + SC_TILE - 1) / SC_TILE);
  if (n_items == 0) {
    scan_empty_kernel<<<1, 1, 0, st>>>(counters);
    count_launch();
    return (int)cudaGetLastError();
  }
  cudaMemsetAsync(block_sums, 0, (size_t)(nblocks + 1) * sizeof(uint32_t), st);
  scan_offsets_kernel<<<nblocks, RS_THREADS, 0, st>>>((int)n_items, n_dev, g.rect, depth_order, offsets, block_sums + 1, block_sums, counters);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3: emit (tile, gaussian) instances in depth order and count instances per tile.  A warp owns 32
// consecutive Gaussians of the depth order; their instances form one contiguous output range which the
// 32 lanes walk together (coalesced writes, balanced no matter how many tiles one Gaussian covers).
// The owner of output slot `pos` is found by a 5-step binary search over the lanes' start offsets.
constexpr int EMIT_THREADS = 256;
constexpr int TILE_HIST_COPIES = 16;   // replicated per-tile counters: spreads the L2 atomic traffic over more lines

__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(int P_cap, const uint32_t* __restrict__ n_dev, int gx, int ntiles, int mode,
                      const uint2* __restrict__ rect, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ offsets, S360Counters* counters, int64_t capacity,
                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ tile_count) {
  const int P = (int)effective_n(P_cap, n_dev);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * EMIT_THREADS + threadIdx.x;   // position in the depth order
  uint32_t gid = 0, off = 0xffffffffu, cnt = 0;
  uint2 r = make_uint2(0u, 0u);
  if (i < P) {
    gid = order[i];
    r = rect[gid];
    off = offsets[i];
    cnt = rect_tiles(r);
  }
  // [begin, end) of this warp's instances; lanes beyond P carry off = 0xffffffff and never own a slot
  const uint32_t begin = __shfl_sync(0xffffffffu, off, 0);
  uint32_t end = (i < P) ? off + cnt : 0u;
  end = __reduce_max_sync(0xffffffffu, end);
  if (begin == 0xffffffffu || end <= begin) return;
  bool overflow = false;
  for (uint32_t p0 = begin; p0 < end; p0 += 32) {   // warp-uniform trip count: every lane joins the shuffles
    const uint32_t pos = p0 + lane;
    // last lane j with off_j <= pos (zero-count lanes share their successor's offset and lose the tie)
    int lo = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const uint32_t v = __shfl_sync(0xffffffffu, off, (lo + s) & 31);
      if (lo + s < 32 && v <= pos) lo += s;
    }
    const uint32_t o_off = __shfl_sync(0xffffffffu, off, lo);
    const uint32_t o_rx = __shfl_sync(0xffffffffu, r.x, lo);
    const uint32_t o_ry = __shfl_sync(0xffffffffu, r.y, lo);
    const uint32_t o_gid = __shfl_sync(0xffffffffu, gid, lo);
    if (pos < end) {
      const uint32_t k = pos - o_off;
      const uint32_t nx = o_rx >> 16;
      const uint32_t ky = k / nx, kx = k - ky * nx;
      int tx = (int)(int16_t)(o_rx & 0xffffu) + (int)kx;
      if (mode == S360_MODE_ERP) { tx %= gx; if (tx < 0) tx += gx; }
      const uint32_t tile = ((o_ry & 0xffffu) + ky) * (uint32_t)gx + (uint32_t)tx;
      if ((int64_t)pos < capacity) {
        keys[pos] = tile;
        vals[pos] = o_gid;
        atomicAdd(&tile_count[(size_t)(blockIdx.x & (TILE_HIST_COPIES - 1)) * ntiles + tile], 1u);
      } else {
        overflow = true;
      }
    }
  }
  if (overflow) atomicOr(&counters->overflow, 1u);
}


// ================================================================================================
// Matrix binning: the stable "sort N instances by tile" step WITHOUT materialising (tile, id) keys.
// The depth order is cut into chunks of MB_CHUNK Gaussians.  (1) mb_count_kernel: one CTA per chunk expands the
// chunk's tile rectangles and counts instances per tile -> row [chunk][tile] of a count matrix.  (2) mb_colscan_kernel:
// exclusive prefix along the chunk axis per tile (in place) + tile totals.  (3) tile_scan_kernel: tile totals -> ranges
// + longest-first schedule.  (4) mb_scatter_kernel: every chunk expands its rectangles again, ranks each instance
// among the chunk's earlier instances of the same tile (warp w owns a contiguous slice of the chunk: per-warp u16
// counters, exclusive prefix over the warps, then in-order ranking with ballot-matched peers) and writes the Gaussian id
// straight to  ranges[tile].x + prefix[chunk][tile] + rank.  Result: point_list sorted by (tile, depth, id), bit-identical
// to emit + two stable radix passes, with 4 B written per instance instead of 8 B emitted + 2 x 16 B sorted, and no
// scan of per-Gaussian offsets at all.  Used whenever tiles <= MB_MAX_TILES and the matrix stays small.
#ifndef S360_MB_CHUNK
#define S360_MB_CHUNK 2048
#endif
constexpr int MB_CHUNK = S360_MB_CHUNK;   // a multiple of 512 (16 warps x 32 Gaussians per round)
static_assert(MB_CHUNK % 512 == 0 && MB_CHUNK >= 512 && MB_CHUNK <= 8192, "MB_CHUNK: whole rounds of 16 warps, u16 counters");
constexpr int MB_MAX_TILES = 8192;    // shared-memory tile histogram of the count kernel: 32 KB
constexpr int MB_BAND_TILES = 2048;   // tiles per band of the scatter kernel (and the widest supported tile row)

bool matrix_binning_ok(int64_t n_items, int64_t tiles, int gx) {
  if (tiles <= 0 || tiles > MB_MAX_TILES || gx > MB_BAND_TILES || n_items <= 0) return false;
  const int64_t chunks = (n_items + MB_CHUNK - 1) / MB_CHUNK;
  return chunks * tiles * 4 <= (256ll << 20);
}
size_t matrix_scratch_bytes(int64_t n_items, int64_t tiles) {
  const int64_t chunks = (n_items + MB_CHUNK - 1) / MB_CHUNK;
  return align_up((size_t)chunks * tiles * 4, 256) + align_up((size_t)tiles * 4, 256);
}

// tile id of the k-th cell (row-major) of a packed tile rectangle; erp wraps the column.  The matrix path holds at most
// MB_MAX_TILES cells per rectangle, so row = floor((k + 0.5) / nx) is exact in float arithmetic (distance of the quotient
// to the next integer >= 0.5 / nx, error of the approximate reciprocal <= 8192 / nx * 3e-7) and the erp column needs at
// most one wrap (x0 >= -gx, x0 + nx <= 2 gx).
__device__ __forceinline__ uint32_t rect_tile(uint32_t rx, uint32_t ry, uint32_t k, int gx, int mode) {
  const uint32_t nx = rx >> 16;
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"((float)nx));
  const uint32_t ky = (uint32_t)(((float)k + 0.5f) * inv), kx = k - ky * nx;
  int tx = (int)(int16_t)(rx & 0xffffu) + (int)kx;
  if (mode == S360_MODE_ERP) { if (tx < 0) tx += gx; else if (tx >= gx) tx -= gx; }
  return ((ry & 0xffffu) + ky) * (uint32_t)gx + (uint32_t)tx;
}

// A warp expands the rectangles of 32 consecutive Gaussians of the depth order (lane j holds Gaussian j) into their
// instances, 32 at a time, IN ORDER (Gaussian, then cell): f(tile, gid, valid) is called by all lanes in lock-step.
// Balanced no matter how many tiles one Gaussian covers (pole-sized rectangles).  The owner of output slot `pos` is the
// last Gaussian-lane whose first slot is <= pos: Gaussians without tiles sort to the END of the depth order (their key is
// 0xFFFFFFFF), so the lanes with tiles form a prefix and  owner = #heads before the batch + #heads inside it up to pos - 1
// (one OR-reduction of the head bits + one ballot per batch); any other pattern takes a 5-step shuffle search.
template <class F>
__device__ __forceinline__ void expand_group(uint32_t gid, uint2 r, uint32_t cnt, int lane, int gx, int mode, F&& f) {
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const uint32_t off = incl - cnt;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  const unsigned nz = __ballot_sync(0xffffffffu, cnt > 0u);
  const bool prefix = (nz & (nz + 1u)) == 0u;
  for (uint32_t p0 = 0; p0 < total; p0 += 32) {
    const uint32_t pos = p0 + lane;
    int lo;
    if (prefix) {
      const uint32_t rel = off - p0;   // wraps for heads before the batch: not < 32
      const unsigned heads = __reduce_or_sync(0xffffffffu, (cnt > 0u && rel < 32u) ? (1u << rel) : 0u);
      const int before = __popc(__ballot_sync(0xffffffffu, cnt > 0u && off < p0));
      lo = before + __popc(heads & ((2u << lane) - 1u)) - 1;
    } else {
      lo = 0;   // last lane j with off_j <= pos (zero-count lanes share their successor's offset and lose the tie)
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, off, (lo + s) & 31);
        if (lo + s < 32 && v <= pos) lo += s;
      }
    }
    lo &= 31;
    const uint32_t o_off = __shfl_sync(0xffffffffu, off, lo);
    const uint32_t o_rx = __shfl_sync(0xffffffffu, r.x, lo);
    const uint32_t o_ry = __shfl_sync(0xffffffffu, r.y, lo);
    const uint32_t o_gid = __shfl_sync(0xffffffffu, gid, lo);
    const bool valid = pos < total;
    const uint32_t tile = valid ? rect_tile(o_rx, o_ry, pos - o_off, gx, mode) : 0xffffffffu;
    f(tile, o_gid, valid);
  }
}

// Order-free expansion (counting only): every lane walks the cells of its own rectangle -- no search, no division; a
// round whose largest rectangle is pole-sized falls back to the balanced expansion above.
template <class F>
__device__ __forceinline__ void expand_unordered(uint32_t gid, uint2 r, uint32_t cnt, int lane, int gx, int mode, F&& f) {
  const uint32_t mx = __reduce_max_sync(0xffffffffu, cnt);
  if (mx > 48u) {
    expand_group(gid, r, cnt, lane, gx, mode, [&](uint32_t tile, uint32_t, bool valid) { if (valid) f(tile); });
    return;
  }
  const int nx = (int)(r.x >> 16), x0 = (int)(int16_t)(r.x & 0xffffu);
  uint32_t rowbase = (r.y & 0xffffu) * (uint32_t)gx;
  int cx = 0;
  for (uint32_t k = 0; k < cnt; k++) {
    int tx = x0 + cx;
    if (mode == S360_MODE_ERP) { if (tx < 0) tx += gx; else if (tx >= gx) tx -= gx; }
    f(rowbase + (uint32_t)tx);
    if (++cx == nx) { cx = 0; rowbase += (uint32_t)gx; }
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32)
mb_count_kernel(int n_cap, const uint32_t* __restrict__ n_dev, int gx, int ntiles, int mode,
                const uint2* __restrict__ rect, const uint32_t* __restrict__ order, uint32_t* __restrict__ matrix) {
  pdl_enter();
  extern __shared__ uint32_t s_cnt[];   // [ntiles]
  constexpr int ROUNDS = MB_CHUNK / (NW * 32);
  const int n = (int)effective_n(n_cap, n_dev);
  const int base = (int)blockIdx.x * MB_CHUNK;
  if (base >= n) return;   // rows of empty chunks are never read (the column scan stops at the same bound)
  for (int t = threadIdx.x; t < ntiles; t += NW * 32) s_cnt[t] = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t gid[ROUNDS];
  uint2 rc[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int i = base + warp * (ROUNDS * 32) + r * 32 + lane;
    gid[r] = i < n ? order[i] : 0u;
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int i = base + warp * (ROUNDS * 32) + r * 32 + lane;
    rc[r] = i < n ? rect[gid[r]] : make_uint2(0u, 0u);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; r++)
    expand_unordered(gid[r], rc[r], rect_tiles(rc[r]), lane, gx, mode, [&](uint32_t tile) { atomicAdd(&s_cnt[tile], 1u); });
  __syncthreads();
  uint32_t* row = matrix + (size_t)blockIdx.x * ntiles;
  for (int t = threadIdx.x; t < ntiles; t += NW * 32) row[t] = s_cnt[t];
}

// exclusive prefix over the chunks for every tile (in place) + instances per tile.  Block = 32 tiles x 32 chunk segments.
__global__ void __launch_bounds__(1024)
mb_colscan_kernel(int n_cap, const uint32_t* __restrict__ n_dev, int ntiles, uint32_t* matrix, uint32_t* __restrict__ tile_total) {
  pdl_enter();
  __shared__ uint32_t s_seg[32][33];
  const int n = (int)effective_n(n_cap, n_dev);
  const int n_chunks = (n + MB_CHUNK - 1) / MB_CHUNK;
  const int tx = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int t = (int)blockIdx.x * 32 + tx;
  const int seg_len = (n_chunks + 31) / 32;
  const int c0 = min(seg * seg_len, n_chunks), c1 = min(c0 + seg_len, n_chunks);
  uint32_t sum = 0;
  if (t < ntiles) {
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = matrix[(size_t)(c + k) * ntiles + t];
#pragma unroll
      for (int k = 0; k < 8; k++) sum += v[k];
    }
    for (; c < c1; c++) sum += matrix[(size_t)c * ntiles + t];
  }
  s_seg[seg][tx] = sum;
  __syncthreads();
  if (seg == 0) {
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) { const uint32_t v = s_seg[k][tx]; s_seg[k][tx] = run; run += v; }
    if (t < ntiles) tile_total[t] = run;
  }
  __syncthreads();
  uint32_t run = s_seg[seg][tx];
  if (t < ntiles) {
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = matrix[(size_t)(c + k) * ntiles + t];
#pragma unroll
      for (int k = 0; k < 8; k++) { matrix[(size_t)(c + k) * ntiles + t] = run; run += v[k]; }
    }
    for (; c < c1; c++) { const uint32_t v = matrix[(size_t)c * ntiles + t]; matrix[(size_t)c * ntiles + t] = run; run += v; }
  }
}

// blockIdx.y = band of tile rows [band * band_rows, (band + 1) * band_rows): every Gaussian's rectangle is clipped to the
// band, so the per-warp counters cover band_rows * gx <= MB_BAND_TILES tiles whatever the image size (instances of
// different tiles never interact, so bands are independent); small images have one band (BANDED = false).  With bands,
// every warp first compacts its 128 Gaussians to those that reach the band (order preserved, staged in shared memory):
// the rounds of the two passes then run over the survivors only, not over four times as many mostly empty rounds.
template <int NW, bool BANDED>
__global__ void __launch_bounds__(NW * 32)
mb_scatter_kernel(int n_cap, const uint32_t* __restrict__ n_dev, int gx, int gy, int band_rows, int ntiles, int mode,
                  const uint2* __restrict__ rect, const uint32_t* __restrict__ order, const uint32_t* __restrict__ prefix,
                  const uint2* __restrict__ ranges, int64_t capacity, uint32_t* __restrict__ point_list,
                  S360Counters* counters) {
  pdl_enter();
  extern __shared__ uint32_t s_mem[];
  constexpr int ROUNDS = MB_CHUNK / (NW * 32);
  const int y_lo = (int)blockIdx.y * band_rows, y_hi = min(y_lo + band_rows, gy);
  const int btiles = (y_hi - y_lo) * gx;         // tiles of this band
  const uint32_t tile0 = (uint32_t)(y_lo * gx);  // first global tile id of the band
  const int half = (band_rows * gx + 1) >> 1;    // u16 counters, two per word
  uint32_t* s_base = s_mem;                      // [band tiles]  first slot of this chunk's instances of the tile
  uint32_t* s_cnt32 = s_mem + band_rows * gx;    // [NW][half] per-warp counters -> exclusive prefixes over the warps
  uint32_t* s_list = s_cnt32 + NW * half;        // BANDED: [NW][ROUNDS * 32][3] compacted (gid, rect.x, rect.y) per warp
  const int n = (int)effective_n(n_cap, n_dev);
  const int base = (int)blockIdx.x * MB_CHUNK;
  if (base >= n) return;
  for (int j = threadIdx.x; j < NW * half; j += NW * 32) s_cnt32[j] = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t gid[ROUNDS];
  uint2 rc[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int i = base + warp * (ROUNDS * 32) + r * 32 + lane;
    gid[r] = i < n ? order[i] : 0u;
  }
  int nrounds = ROUNDS;
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int i = base + warp * (ROUNDS * 32) + r * 32 + lane;
    rc[r] = i < n ? rect[gid[r]] : make_uint2(0u, 0u);
  }
  if (BANDED) {
    uint32_t* my_list = s_list + warp * (ROUNDS * 32 * 3);
    int kept = 0;
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      // clip the rows to the band (an empty intersection has no cells)
      const uint2 q = rc[r];
      const int r0 = (int)(q.y & 0xffffu), r1 = r0 + (int)(q.y >> 16);
      const int c0 = max(r0, y_lo), c1 = min(r1, y_hi);
      const bool keep = c1 > c0 && (q.x >> 16) != 0u;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        uint32_t* e = my_list + 3 * (kept + __popc(m & ((1u << lane) - 1u)));
        e[0] = gid[r]; e[1] = q.x; e[2] = (uint32_t)c0 | ((uint32_t)(c1 - c0) << 16);
      }
      kept += __popc(m);
    }
    nrounds = (kept + 31) >> 5;
    __syncwarp();
    // re-read as rounds of 32 survivors (the tail of the last round has no cells)
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      const int k = r * 32 + lane;
      const bool in = k < kept;
      gid[r] = in ? my_list[3 * k] : 0u;
      rc[r] = in ? make_uint2(my_list[3 * k + 1], my_list[3 * k + 2]) : make_uint2(0u, 0u);
    }
  }
  {
    const uint32_t* prow = prefix + (size_t)blockIdx.x * ntiles + tile0;
    for (int t = threadIdx.x; t < btiles; t += NW * 32) s_base[t] = ranges[tile0 + t].x + prow[t];
  }
  __syncthreads();
  // pass 1: instances per (warp, tile); packed u16 pairs take native 32-bit shared-memory adds
  uint32_t* my32 = s_cnt32 + warp * half;
#pragma unroll
  for (int r = 0; r < ROUNDS; r++)
    if (r < nrounds)
      expand_unordered(gid[r], rc[r], rect_tiles(rc[r]), lane, gx, mode,
                       [&](uint32_t tile) { tile -= tile0; atomicAdd(&my32[tile >> 1], 1u << ((tile & 1u) * 16u)); });
  __syncthreads();
  // exclusive prefix over the warps, both halves of a word at once (a chunk has at most MB_CHUNK instances per tile)
  for (int j = threadIdx.x; j < half; j += NW * 32) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) { const uint32_t c = s_cnt32[w * half + j]; s_cnt32[w * half + j] = run; run += c; }
  }
  __syncthreads();
  // pass 2: in-order ranking.  Lanes of one batch that hit the same tile are found with MATCH.ANY (eleven ballots, one per
  // tile-id bit, measured the same: the kernel is latency-bound); all of them read the warp's counter of that tile
  // (shared-memory broadcast), the lowest one then advances it by the size of the group.
  unsigned short* my16 = reinterpret_cast<unsigned short*>(my32);
  const unsigned lt_mask = (1u << lane) - 1u;
  bool overflow = false;
#pragma unroll
  for (int r = 0; r < ROUNDS; r++)
    if (r < nrounds)
      expand_group(gid[r], rc[r], rect_tiles(rc[r]), lane, gx, mode, [&](uint32_t tile, uint32_t g, bool valid) {
        tile -= tile0;   // lanes past the end carry 0xffffffff - tile0: still equal to each other, never dereferenced
        const unsigned peers = __match_any_sync(0xffffffffu, tile);
        uint32_t old = 0;
        if (valid) old = my16[tile];
        __syncwarp();
        if (valid && (peers & lt_mask) == 0u) my16[tile] = (unsigned short)(old + (uint32_t)__popc(peers));
        __syncwarp();
        if (valid) {
          const uint32_t dst = s_base[tile] + old + (uint32_t)__popc(peers & lt_mask);
          if ((int64_t)dst < capacity) point_list[dst] = g; else overflow = true;
        }
      });
  if (overflow) atomicOr(&counters->overflow, 1u);
}

// scratch: [matrix chunks x tiles][tile totals]; the three launchers below run in this order with the tile scan
// (launch_tile_scan with copies = 1 on mb_tile_totals) between the column scan and the scatter
uint32_t* mb_tile_totals(void* scratch, int64_t n_items, int ntiles) {
  const int64_t chunks = (n_items + MB_CHUNK - 1) / MB_CHUNK;
  return (uint32_t*)((char*)scratch + align_up((size_t)chunks * ntiles * 4, 256));
}

int launch_mb_count(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, GeomState g,
                    const uint32_t* depth_order, void* scratch, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = NV * ((v.image_height + TILE - 1) / TILE);
  const int ntiles = gx * gy;
  const int chunks = (int)((n_items + MB_CHUNK - 1) / MB_CHUNK);
  if (chunks == 0) return 0;
  // 16 warps per chunk: the per-warp work (rounds of 32 Gaussians, each a chain of shuffles and shared-memory atomics) is
  // latency-bound, so more and shorter chains win (8 warps: 25 us, measured)
  launch_pdl(mb_count_kernel<16>, dim3(chunks), dim3(512), (size_t)ntiles * 4, st, (int)n_items, n_dev, gx, ntiles, v.mode, g.rect,
             depth_order, (uint32_t*)scratch);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_mb_colscan(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, void* scratch, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = NV * ((v.image_height + TILE - 1) / TILE);
  const int ntiles = gx * gy;
  launch_pdl(mb_colscan_kernel, dim3((ntiles + 31) / 32), dim3(1024), 0, st, (int)n_items, n_dev, ntiles, (uint32_t*)scratch,
             mb_tile_totals(scratch, n_items, ntiles));
  count_launch();
  return (int)cudaGetLastError();
}

int launch_mb_scatter(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, GeomState g,
                      const uint32_t* depth_order, S360Counters* counters, int64_t capacity, uint32_t* point_list,
                      const uint2* ranges, void* scratch, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = NV * ((v.image_height + TILE - 1) / TILE);
  const int ntiles = gx * gy;
  const int chunks = (int)((n_items + MB_CHUNK - 1) / MB_CHUNK);
  if (chunks == 0) return 0;
  // 16 warps per chunk; tile rows are cut into bands of at most MB_BAND_TILES tiles so that the per-warp u16 counters stay
  // at 4 KB per warp (72 KB per CTA, three CTAs per SM) whatever the image size -- 1024x2048 (8192 tiles) = four bands
  const int band_rows = max(1, MB_BAND_TILES / gx);
  const int bands = (gy + band_rows - 1) / band_rows;
  const int btiles = band_rows * gx;
  const size_t smem = (size_t)btiles * 4 + (size_t)16 * ((btiles + 1) / 2) * 4 + (bands > 1 ? (size_t)MB_CHUNK * 3 * 4 : 0);
  const uint32_t* prefix = (const uint32_t*)scratch;
#define S360_SCATTER(BANDED_) do { \
    if (smem > 48 * 1024) cudaFuncSetAttribute(mb_scatter_kernel<16, BANDED_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    launch_pdl(mb_scatter_kernel<16, BANDED_>, dim3(chunks, bands), dim3(512), smem, st, (int)n_items, n_dev, gx, gy, band_rows, ntiles, \
               v.mode, g.rect, depth_order, prefix, ranges, capacity, point_list, counters); } while (0)
  if (bands > 1) S360_SCATTER(true); else S360_SCATTER(false);
#undef S360_SCATTER
  count_launch();
  return (int)cudaGetLastError();
}

int tile_hist_copies() { return TILE_HIST_COPIES; }

// One block: exclusive scan of the per-tile instance counts -> tile ranges, plus the digit histograms of
// the (up to two) tile-sort passes, which are just partial sums of the same counts.
// copies: replicated histograms to sum (TILE_HIST_COPIES after the emission kernel, 1 after the matrix column scan);
// capacity: ranges are clipped to the instances that were actually stored (overflow of a caller-chosen capacity).
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int ntiles, const uint32_t* __restrict__ tile_count, int copies, uint32_t capacity, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ order, uint32_t* __restrict__ work, uint32_t* __restrict__ hist, int npasses) {
  pdl_enter();
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_hist[3][RS_BINS];
  __shared__ uint32_t s_bucket[1024];   // longest-first schedule: counting sort of the tiles by instances / 32
  s_bucket[threadIdx.x] = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  for (int i = threadIdx.x; i < 3 * RS_BINS; i += 1024) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    uint32_t x = 0;
    if (i < ntiles) {
      for (int c = 0; c < copies; c++) x += tile_count[(size_t)c * ntiles + i];
    }
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
    if (i < ntiles) {
      const uint32_t start = carry + woff + incl - x;
      ranges[i] = x ? make_uint2(min(start, capacity), min(start + x, capacity)) : make_uint2(0u, 0u);
      atomicAdd(&s_bucket[1023u - min(x >> 5, 1023u)], 1u);   // bucket 0 = heaviest
      if (x) {
        atomicAdd(&s_hist[0][i & 0xff], x);
        if (npasses > 1) atomicAdd(&s_hist[1][(i >> 8) & 0xff], x);
        if (npasses > 2) atomicAdd(&s_hist[2][(i >> 16) & 0xff], x);
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (hist != nullptr)
    for (int i = threadIdx.x; i < npasses * RS_BINS; i += 1024) hist[i] = (&s_hist[0][0])[i];
  // exclusive scan of the 1024 buckets (one per thread), then every tile claims a slot in its bucket
  {
    const uint32_t x = s_bucket[threadIdx.x];
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_w[w];
    s_bucket[threadIdx.x] = woff + incl - x;
    __syncthreads();
    for (int i = threadIdx.x; i < ntiles; i += 1024) {
      uint32_t c = 0;
      for (int k = 0; k < copies; k++) c += tile_count[(size_t)k * ntiles + i];
      const uint32_t pos = atomicAdd(&s_bucket[1023u - min(c >> 5, 1023u)], 1u);
      order[pos] = (uint32_t)i;
      work[i] = 0u;
    }
  }
}

// longest-first schedule of the backward pass from the work the forward pass measured per tile
__global__ void __launch_bounds__(1024)
tile_order_kernel(int ntiles, const uint32_t* __restrict__ work, uint32_t* __restrict__ order) {
  pdl_enter();
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_bucket[1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s_bucket[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < ntiles; i += 1024) atomicAdd(&s_bucket[1023u - min(work[i] >> 5, 1023u)], 1u);
  __syncthreads();
  const uint32_t x = s_bucket[threadIdx.x];
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
  __syncthreads();
  s_bucket[threadIdx.x] = woff + incl - x;
  __syncthreads();
  for (int i = threadIdx.x; i < ntiles; i += 1024) {
    const uint32_t pos = atomicAdd(&s_bucket[1023u - min(work[i] >> 5, 1023u)], 1u);
    order[pos] = (uint32_t)i;
  }
}

int launch_tile_order(const S360View& v, int NV, const uint32_t* work, uint32_t* order, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = NV * ((v.image_height + TILE - 1) / TILE);
  launch_pdl(tile_order_kernel, dim3(1), dim3(1024), 0, st, gx * gy, work, order);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_emit(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, GeomState g,
                const uint32_t* depth_order, const uint32_t* offsets, S360Counters* counters, int64_t capacity,
                uint32_t* keys, uint32_t* vals, uint32_t* tile_count, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = NV * ((v.image_height + TILE - 1) / TILE);
  cudaMemsetAsync(tile_count, 0, (size_t)gx * gy * TILE_HIST_COPIES * sizeof(uint32_t), st);
  if (n_items == 0) return 0;
  emit_instances_kernel<<<(int)((n_items + EMIT_THREADS - 1) / EMIT_THREADS), EMIT_THREADS, 0, st>>>(
      (int)n_items, n_dev, gx, gx * gy, v.mode, g.rect, depth_order, offsets, counters, capacity, keys, vals, tile_count);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_tile_scan(const S360View& v, int NV, const uint32_t* tile_count, int copies, int64_t capacity, uint2* ranges,
                     uint32_t* order, uint32_t* work, uint32_t* hist, int npasses, cudaStream_t st) {
  const int gx = (v.image_width + TILE - 1) / TILE, gy = NV * ((v.image_height + TILE - 1) / TILE);
  const uint32_t cap = (uint32_t)(capacity < 0xffffffffll ? capacity : 0xffffffffll);
  launch_pdl(tile_scan_kernel, dim3(1), dim3(1024), 0, st, gx * gy, tile_count, copies, cap, ranges, order, work, hist, npasses);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace s360
