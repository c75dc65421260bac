// render.cu -- per-tile compositing kernels.
//   K6 render_forward_kernel : front-to-back alpha compositing            
//   K7 render_backward_kernel: back-to-front gradient pass                 
// One CTA = one 16x16 tile = 4 warps; a warp owns an 8x8 pixel block and a lane the two pixels (x, y) and
// (x, y + 4), whose state is packed in float2 and advanced with Blackwell's two-wide FP32 instructions
// (FFMA2 / FMUL2 / FADD2).  The kernels are warp-autonomous (no CTA barrier): every warp streams the tile's
// depth-sorted instance list itself in chunks of 32 -- lane j gathers instance j (its id was prefetched one chunk
// earlier, the 48-B record comes from L2/L1, which the four warps of a tile share) -- tests its own instance
// against the warp's pixel rectangle with an exact "can the alpha >= 1/255 ellipse reach it" test, and the ballot
// survivors write their staged parameters COMPACTED, in list order, into a per-warp shared-memory slice (conic
// pre-multiplied by -0.5*log2(e), log2(opacity), colour, centre relative to the warp block, list position).  The
// evaluation loop then just walks that slice (3 LDS.128 per survivor, no ffs / shuffle bookkeeping):
// alpha = ex2(A'dx^2 + C'dy^2 + B'dxdy + log2 o).  Blending is branch-free: a pixel that does not take an instance
// runs the recurrences with alpha = 0.  erp: the per-pixel seam wrap is needed only by Gaussians wider than half the
// panorama; whether a chunk holds one is decided once per chunk and selects a second instantiation of the loop.
// The CTAs pick their tile from a longest-first schedule; V stacked views (batched path) are just more tiles.
//
// Backward: per (pixel, instance) only the moments q, q*dx, q*dy, q*dx^2, q*dxdy, q*dy^2 (q = G * dL/dalpha) and the
// three colour terms are formed; they are summed over the warp with a transposed butterfly (14 shuffles) and nine
// lanes issue one fire-and-forget RED.ADD.F32 each into the per-Gaussian accumulator.  Moments are converted to
// dL/d{mean2D, conic, opacity} once per Gaussian in preprocess_backward_kernel.  Replaces upstream renderCUDA
// (forward.cu / backward.cu) behind /root/reference/src/model/decoder/cuda_splatting.py:113-124.
#include <type_traits>

#include "common.cuh"
#include "render_cull.cuh"

namespace s360 {

template <int MODE>
__device__ __forceinline__ float wrap_dx(float dx, float W, float halfW) {
  if (MODE == S360_MODE_ERP) {
    if (dx > halfW) dx -= W;
    else if (dx < -halfW) dx += W;
  }
  return dx;
}

// nal if (p <= 0 && -nal >= amin) else 0 -- one predicate chain and one select (the compiler's own lowering of the
// conditional spends two selects per pixel)
__device__ __forceinline__ float take_alpha(float p, float nal, float amin) {
  float r;
  asm("{\n\t.reg .pred q, t;\n\tsetp.le.f32 q, %1, 0f00000000;\n\tsetp.ge.and.f32 t, %2, %3, q;\n\t"
      "selp.f32 %0, %4, 0f00000000, t;\n\t}" : "=f"(r) : "f"(p), "f"(-nal), "f"(amin), "f"(nal));
  return r;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifndef S360_FWD_PREFETCH
#define S360_FWD_PREFETCH 0   // 1: next chunk's records travel in registers during compositing; 0: only its ids do (measured faster: fewer registers)
#endif
#ifndef S360_BWD_PREFETCH
#define S360_BWD_PREFETCH 0
#endif
#ifndef S360_FWD_MINB
#define S360_FWD_MINB 6
#endif
#ifndef S360_BWD_MINB
#define S360_BWD_MINB 7
#endif

#ifndef S360_FWD_UNROLL
#define S360_FWD_UNROLL 2   // survivors evaluated per loop trip of the forward kernel (measured: 1 -> 2 = -8 %)
#endif
#define S360_PRAGMA_(x) _Pragma(#x)
#define S360_PRAGMA(x) S360_PRAGMA_(x)

constexpr int NWARPS = RT / 32;

// Instrumented build (-DS360_COUNTERS=1, tools/counters.py): how much of the evaluated work is useful.
//   [0] fwd warp-chunks  [1] fwd (warp, instance) tests  [2] fwd survivors  [3] fwd survivors some pixel takes
//   [4] fwd (pixel, instance) pairs taken   [5] fwd instances in the tile lists (range sizes)
//   [8] bwd warp-chunks  [9] bwd tests      [10] bwd survivors   [11] bwd survivors some pixel takes   [12] bwd pairs taken
#ifndef S360_COUNTERS
#define S360_COUNTERS 0
#endif
#if S360_COUNTERS
__device__ unsigned long long g_counters[16];
#define S360_COUNT(var, x) var += (x)
#else
#define S360_COUNT(var, x)
#endif
int read_counters(unsigned long long* out, int reset, cudaStream_t st) {
#if S360_COUNTERS
  cudaStreamSynchronize(st);
  int rc = (int)cudaMemcpyFromSymbol(out, g_counters, sizeof(unsigned long long) * 16);
  if (!rc && reset) { unsigned long long z[16] = {0}; rc = (int)cudaMemcpyToSymbol(g_counters, z, sizeof(z)); }
  return rc;
#else
  (void)reset; (void)st;
  for (int i = 0; i < 16; i++) out[i] = 0;
  return S360_ERR_UNSUPPORTED;
#endif
}

// Warp-autonomous streaming of a tile's instance list: lane j of every warp gathers instance
// (chunk*32 + j) itself (the four warps of a tile hit the same lines in L1); the ids of the next two chunks are
// always in flight (S360_*_PREFETCH=1 additionally keeps the next chunk's records in registers), and no CTA-wide
// barrier exists in the kernel.
struct ChunkRegs {
  float4 r0, r1, r2;
  uint32_t gid;
};

__device__ __forceinline__ void load_records(ChunkRegs& c, const float4* __restrict__ rec, bool valid) {
  if (valid) {
    c.r0 = __ldg(rec + 3 * (size_t)c.gid);
    c.r1 = __ldg(rec + 3 * (size_t)c.gid + 1);
    c.r2 = __ldg(rec + 3 * (size_t)c.gid + 2);
  }
}

// S360_ASYNC_STAGE=1 (A/B variant, north_star's "async-copy staging of sorted Gaussian attributes into shared memory"):
// the next chunk's records travel global -> shared by cp.async (LDGSTS, no registers held) while the current chunk is
// composited, double-buffered per warp; every lane copies and later reads only its own record, so no barrier is needed.
#ifndef S360_ASYNC_STAGE
#define S360_ASYNC_STAGE 0
#endif
#if S360_ASYNC_STAGE
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void issue_records(float4 (*buf)[32], uint32_t gid, const float4* __restrict__ rec, bool valid, int lane) {
  if (valid) {
    cp_async16(&buf[0][lane], rec + 3 * (size_t)gid);
    cp_async16(&buf[1][lane], rec + 3 * (size_t)gid + 1);
    cp_async16(&buf[2][lane], rec + 3 * (size_t)gid + 2);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fetch_records(ChunkRegs& c, float4 (*buf)[32], int lane) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  c.r0 = buf[0][lane]; c.r1 = buf[1][lane]; c.r2 = buf[2][lane];
}
#endif

template <int MODE, bool DEPTH>
__global__ void __launch_bounds__(RT, S360_FWD_MINB)
render_forward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                      const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ order, uint32_t* __restrict__ work,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                      float* __restrict__ out_depth, const DepthSpec dspec) {
  pdl_enter();
  // survivors of the current chunk, compacted in list order: [0] = A', B', C' (log2-scaled conic), log2(opacity);
  // [1] = -r, -g, -b, -depth value; [2] = centre relative to the warp block (x, y), wide flag, position in the tile list
  __shared__ float4 s_sv[NWARPS][3][32];
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int tile = (int)order[blockIdx.x];   // heaviest tiles first (longest-processing-time schedule)
  // batched path: the views are stacked on a virtual image, tile row = view * gy + ty (one view: view = 0)
  const int tx = tile % gx, tyv = tile / gx, view = tyv / gy, ty = tyv - view * gy;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py0 = wy0 + (lane >> 3), py1 = py0 + 4;
  const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
  const float pxf = (float)px, pyf = (float)py0;
  const float wcx = (float)wx0 + HALF_W, wcy = (float)wy0 + HALF_H;
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];

  // packed state of the lane's two pixels (.x = row py0, .y = row py0 + 4); FFMA2/FMUL2/FADD2 are Blackwell's
  // two-wide FP32 instructions
  float2 T = make_float2(1.f, 1.f), Cr = make_float2(0.f, 0.f), Cg = Cr, Cb = Cr, Cd = Cr;
  uint32_t last0 = 0, last1 = 0;
  const float INF = __int_as_float(0x7f800000);
  float amin0 = in0 ? ALPHA_MIN : INF, amin1 = in1 ? ALPHA_MIN : INF;   // +inf once the pixel is finished
  // pixel offsets from the warp-block centre: instance centres are broadcast relative to that centre, which keeps
  // dx, dy accurate to an ulp of the (small) distance even at coordinates in the thousands
  const float offx = pxf - wcx;
  const float2 npy = make_float2(-(pyf - wcy), -(pyf + 4.f - wcy));

  ChunkRegs nx;
  nx.r0 = nx.r1 = nx.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
  nx.gid = 0;
  if (range.x + lane < range.y) nx.gid = point_list[range.x + lane];
#if S360_ASYNC_STAGE
  __shared__ float4 s_rec[NWARPS][2][3][32];
  int rbuf = 0;
  issue_records(s_rec[warp][0], nx.gid, rec, range.x + lane < range.y, lane);
#elif S360_FWD_PREFETCH
  load_records(nx, rec, range.x + lane < range.y);
#endif
  uint32_t gid2 = (range.x + 32 + lane < range.y) ? point_list[range.x + 32 + lane] : 0u;
#if S360_COUNTERS
  unsigned long long c_chunks = 0, c_tests = 0, c_surv = 0, c_hit = 0, c_pairs = 0;
#endif

  for (uint32_t base = range.x; base < range.y; base += 32) {
    if (__all_sync(0xffffffffu, amin0 == INF && amin1 == INF)) break;
    const bool valid = base + lane < range.y;
    S360_COUNT(c_chunks, 1); S360_COUNT(c_tests, __popc(__ballot_sync(0xffffffffu, valid)));
#if S360_ASYNC_STAGE
    fetch_records(nx, s_rec[warp][rbuf], lane);
#elif !S360_FWD_PREFETCH
    load_records(nx, rec, valid);
#endif
    float4 cull, ev, col;
    stage_instance(nx.r0, nx.r1, nx.r2, cull, ev, col);
    const float thr = col.w;
    const bool huge = MODE == S360_MODE_ERP && record_is_wide(nx.r1, nx.r2, halfW);
    const float dval = DEPTH ? -depth_value(dspec, nx.r2.w) : 0.f;   // per-Gaussian value of the fused depth channel (negated like rgb)
    // keep the pipeline full: records of the next chunk, ids of the one after
    nx.gid = gid2;
#if S360_ASYNC_STAGE
    rbuf ^= 1;
    issue_records(s_rec[warp][rbuf], nx.gid, rec, base + 32 + lane < range.y, lane);
#elif S360_FWD_PREFETCH
    load_records(nx, rec, base + 32 + lane < range.y);
#endif
    gid2 = (base + 64 + lane < range.y) ? point_list[base + 64 + lane] : 0u;

    // centre relative to this warp's block centre (erp: nearest periodic copy)
    const float ddx = wrap_dx<MODE>(cull.x - wcx, Wf, halfW), ddy = cull.y - wcy;
    const bool hit = valid && (huge || rect_can_contribute(cull, ev, thr, ddx, ddy));
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      // survivors are written compacted, in list order: the evaluation loop just walks the slice
      const int slot = __popc(mask & ((1u << lane) - 1u));
      s_sv[warp][0][slot] = ev;
      s_sv[warp][1][slot] = make_float4(-col.x, -col.y, -col.z, dval);   // negated: the loop carries -alpha
      s_sv[warp][2][slot] = make_float4(ddx, ddy, 0.f, __uint_as_float(base - range.x + (uint32_t)lane + 1u));
    }
    // erp: does any survivor of this chunk need the per-pixel wrap (wider than half the panorama)?  Almost never.
    const bool any_wide = MODE == S360_MODE_ERP && __any_sync(0xffffffffu, hit && huge);
    __syncwarp();
    const int nsv = __popc(mask);
    S360_COUNT(c_surv, nsv);
    const float4* sv = &s_sv[warp][0][0];
    auto composite = [&](auto wide_tag) {
      constexpr bool WIDE = decltype(wide_tag)::value;
      S360_PRAGMA(unroll S360_FWD_UNROLL)
      for (int k = 0; k < nsv; k++) {
        const float4 e = sv[k];
        const float4 c = sv[32 + k];
        const float4 g = sv[64 + k];
        float dx = g.x - offx;
        if (WIDE) dx = wrap_dx<MODE>(dx, Wf, halfW);
        const float2 dy = __fadd2_rn(make_float2(g.y, g.y), npy);
        const float u = e.y * dx, pb = e.x * dx * dx;
        const float2 p = __ffma2_rn(__ffma2_rn(make_float2(e.z, e.z), dy, make_float2(u, u)), dy, make_float2(pb, pb));
        const float2 a = __fadd2_rn(p, make_float2(e.w, e.w));          // power * log2(e) + log2(opacity)
        const float nal0 = fmaxf(-ALPHA_MAX, -ex2_approx(a.x)), nal1 = fmaxf(-ALPHA_MAX, -ex2_approx(a.y));   // -alpha
        // -alpha if this pixel takes the Gaussian (power <= 0, alpha >= 1/255, pixel not finished), else 0
        float2 nae = make_float2(take_alpha(p.x, nal0, amin0), take_alpha(p.y, nal1, amin1));
        const float2 tt = __ffma2_rn(T, nae, T);                          // T (1 - alpha); == T when not taken
        if (tt.x < T_EPS) { nae.x = 0.f; amin0 = INF; }                   // would saturate: not blended, pixel done
        if (tt.y < T_EPS) { nae.y = 0.f; amin1 = INF; }
        const float2 nw = __fmul2_rn(nae, T);                             // -alpha T
        Cr = __ffma2_rn(make_float2(c.x, c.x), nw, Cr);                   // c holds -rgb
        Cg = __ffma2_rn(make_float2(c.y, c.y), nw, Cg);
        Cb = __ffma2_rn(make_float2(c.z, c.z), nw, Cb);
        if (DEPTH) Cd = __ffma2_rn(make_float2(c.w, c.w), nw, Cd);
        T = __ffma2_rn(T, nae, T);
        const uint32_t pos = __float_as_uint(g.w);
        if (nae.x < 0.f) last0 = pos;
        if (nae.y < 0.f) last1 = pos;
#if S360_COUNTERS
        { const unsigned b0 = __ballot_sync(0xffffffffu, nae.x < 0.f), b1 = __ballot_sync(0xffffffffu, nae.y < 0.f);
          c_hit += (b0 | b1) ? 1 : 0; c_pairs += __popc(b0) + __popc(b1); }
#endif
      }
    };
    if (any_wide) composite(std::true_type{}); else composite(std::false_type{});
    __syncwarp();
  }
#if S360_COUNTERS
  if (lane == 0) {
    atomicAdd(&g_counters[0], c_chunks); atomicAdd(&g_counters[1], c_tests); atomicAdd(&g_counters[2], c_surv);
    atomicAdd(&g_counters[3], c_hit); atomicAdd(&g_counters[4], c_pairs);
    if (warp == 0) atomicAdd(&g_counters[5], (unsigned long long)(range.y - range.x));
  }
#endif
  // how far this warp had to walk: the backward pass replays at most that much of the tile's list
  {
    const uint32_t walked = __reduce_max_sync(0xffffffffu, max(last0, last1));
    if (lane == 0 && walked) atomicMax(&work[tile], walked);
  }
  const size_t plane = (size_t)H * W;
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  final_T += (size_t)view * plane; n_contrib += (size_t)view * plane;   // pixel state and outputs of this view
  out_color += (size_t)view * 3 * plane;
  if (DEPTH) out_depth += (size_t)view * plane;
  if (in0) {
    const size_t pid = (size_t)py0 * W + px;
    final_T[pid] = T.x; n_contrib[pid] = last0;
    out_color[pid] = Cr.x + T.x * b0; out_color[plane + pid] = Cg.x + T.x * b1; out_color[2 * plane + pid] = Cb.x + T.x * b2;
    if (DEPTH) out_depth[pid] = Cd.x;
  }
  if (in1) {
    const size_t pid = (size_t)py1 * W + px;
    final_T[pid] = T.y; n_contrib[pid] = last1;
    out_color[pid] = Cr.y + T.y * b0; out_color[plane + pid] = Cg.y + T.y * b1; out_color[2 * plane + pid] = Cb.y + T.y * b2;
    if (DEPTH) out_depth[pid] = Cd.y;
  }
}

int launch_render_forward(const S360View& v, int NV, GeomState g, const uint32_t* point_list, ImageState img,
                          float* out_color, float* out_depth, int depth_mode, float depth_near, float depth_far,
                          cudaStream_t st) {
  const int W = v.image_width, H = v.image_height;
  const int tiles = NV * ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  if (tiles == 0) return 0;
  DepthSpec ds;
  ds.mode = depth_mode; ds.inv_scale = 1.f / v.scene_scale; ds.near = depth_near; ds.far = depth_far;
#define S360_LAUNCH_FWD(MODE_, DEPTH_) launch_pdl(render_forward_kernel<MODE_, DEPTH_>, dim3(tiles), dim3(RT), 0, st, \
      W, H, v.bg, g.rec, point_list, img.ranges, img.order, img.work, img.final_T, img.n_contrib, out_color, out_depth, ds)
  if (v.mode == S360_MODE_PINHOLE) { if (out_depth) S360_LAUNCH_FWD(S360_MODE_PINHOLE, true); else S360_LAUNCH_FWD(S360_MODE_PINHOLE, false); }
  else { if (out_depth) S360_LAUNCH_FWD(S360_MODE_ERP, true); else S360_LAUNCH_FWD(S360_MODE_ERP, false); }
#undef S360_LAUNCH_FWD
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Sum eight per-lane values across the warp with a transposed butterfly: after the call every lane
// holds the complete sum of value index ((lane >> 2) & 7).  9 shuffles instead of 40.  The "which half do I keep" choices
// are bit-selects with per-lane masks (m16 / m8 / m4 = all ones where that lane-id bit is set): one LOP3 each and no
// predicate -- with predicates the compiler re-derived three of them per survivor (ISETP), the loop has only seven.
__device__ __forceinline__ float bsel(float a, float b, uint32_t m) {   // m ? a : b
  return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m));
}
__device__ __forceinline__ float warp_reduce8(float v0, float v1, float v2, float v3, float v4, float v5,
                                              float v6, float v7, uint32_t m16, uint32_t m8, uint32_t m4) {
  const unsigned F = 0xffffffffu;
  float w0 = bsel(v4, v0, m16) + __shfl_xor_sync(F, bsel(v0, v4, m16), 16);
  float w1 = bsel(v5, v1, m16) + __shfl_xor_sync(F, bsel(v1, v5, m16), 16);
  float w2 = bsel(v6, v2, m16) + __shfl_xor_sync(F, bsel(v2, v6, m16), 16);
  float w3 = bsel(v7, v3, m16) + __shfl_xor_sync(F, bsel(v3, v7, m16), 16);
  float u0 = bsel(w2, w0, m8) + __shfl_xor_sync(F, bsel(w0, w2, m8), 8);
  float u1 = bsel(w3, w1, m8) + __shfl_xor_sync(F, bsel(w1, w3, m8), 8);
  float s = bsel(u1, u0, m4) + __shfl_xor_sync(F, bsel(u0, u1, m4), 4);
  s += __shfl_xor_sync(F, s, 2);
  s += __shfl_xor_sync(F, s, 1);
  return s;  // value index = 4*bit4 + 2*bit3 + bit2 of the lane id = (lane >> 2) & 7
}

constexpr int NACC = 9;         // colour x3, q*dx, q*dy, q*dx^2, q*dxdy, q*dy^2, q

// running state of the lane's pixel pair in the back-to-front pass, packed for FFMA2/FMUL2/FADD2
// (.x = row py0, .y = row py0 + 4).  Upstream carries accum_rec[3] (the colour behind the current instance) and forms
// dL/dalpha = sum_c (c_c - accum_rec_c) T dL/dC_c.  Here the colour behind instance i enters only through the scalar
//   B_i = sum_{j behind i} w_j alpha_j T_j + T_final sum_c bg_c dL/dC_c,   w_j = sum_c c_{j,c} dL/dC_c,
// and dL/dalpha_i = T_i w_i - B_i / (1 - alpha_i): one running value per pixel instead of three, no "previous
// instance" bookkeeping.  nB = -B.
struct PairB {
  float2 T, nB;
  float2 dp0, dp1, dp2;
  uint32_t lastc0, lastc1;
};

__device__ __forceinline__ void pair_init(PairB& p, bool in0, bool in1, size_t pid0, size_t pid1, size_t plane,
                                          const float* final_T, const uint32_t* n_contrib, const float* dL,
                                          const float* bg) {
  const float T0 = in0 ? final_T[pid0] : 0.f, T1 = in1 ? final_T[pid1] : 0.f;
  p.lastc0 = in0 ? n_contrib[pid0] : 0u;
  p.lastc1 = in1 ? n_contrib[pid1] : 0u;
  p.dp0 = make_float2(in0 ? dL[pid0] : 0.f, in1 ? dL[pid1] : 0.f);
  p.dp1 = make_float2(in0 ? dL[plane + pid0] : 0.f, in1 ? dL[plane + pid1] : 0.f);
  p.dp2 = make_float2(in0 ? dL[2 * plane + pid0] : 0.f, in1 ? dL[2 * plane + pid1] : 0.f);
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  p.nB = make_float2(-T0 * (b0 * p.dp0.x + b1 * p.dp1.x + b2 * p.dp2.x), -T1 * (b0 * p.dp0.y + b1 * p.dp1.y + b2 * p.dp2.y));
  p.T = make_float2(T0, T1);
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE, bool DEPTH>
__global__ void __launch_bounds__(RT, S360_BWD_MINB)
render_backward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                       const uint32_t* __restrict__ order,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth, const DepthSpec dspec,
                       float* __restrict__ acc) {
  pdl_enter();
  // survivors of the current chunk, compacted back-to-front: [0] = A', B', C', log2(opacity); [1] = r, g, b, bits of the
  // Gaussian id; [2] = centre relative to the warp block (x, y), wide flag, position in the tile list
  __shared__ float4 s_sv[NWARPS][3][32];
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int tile = (int)order[blockIdx.x];   // heaviest tiles first (longest-processing-time schedule)
  // batched path: the views are stacked on a virtual image, tile row = view * gy + ty (one view: view = 0)
  const int tx = tile % gx, tyv = tile / gx, view = tyv / gy, ty = tyv - view * gy;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py0 = wy0 + (lane >> 3), py1 = py0 + 4;
  const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
  const float pxf = (float)px, pyf = (float)py0;
  const float wcx = (float)wx0 + HALF_W, wcy = (float)wy0 + HALF_H;
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];
  const size_t plane = (size_t)H * W;

  PairB S;
  pair_init(S, in0, in1, (size_t)py0 * W + px, (size_t)py1 * W + px, plane, final_T + (size_t)view * plane,
            n_contrib + (size_t)view * plane, dL_dcolor + (size_t)view * 3 * plane, bg);
  // fused depth channel with gradient (DEPTH): a fourth blended channel without background
  float2 dpd = make_float2(0.f, 0.f);
  if (DEPTH) {
    const float* gd = dL_ddepth + (size_t)view * plane;
    dpd = make_float2(in0 ? gd[(size_t)py0 * W + px] : 0.f, in1 ? gd[(size_t)py1 * W + px] : 0.f);
  }
  // pixel offsets from the warp-block centre: instance centres are broadcast relative to that centre, which keeps
  // dx, dy accurate to an ulp of the (small) distance even at coordinates in the thousands
  const float offx = pxf - wcx;
  const float2 npy = make_float2(-(pyf - wcy), -(pyf + 4.f - wcy));
  // lane roles of the per-survivor reduction: butterfly masks, and which accumulator slot this lane adds to
  // (lanes 0, 4, .., 28 hold sums 0..7, lane 1 the ninth, lane 17 the depth channel's tenth; -1: none)
  const uint32_t m16 = (lane & 16) ? 0xffffffffu : 0u, m8 = (lane & 8) ? 0xffffffffu : 0u, m4 = (lane & 4) ? 0xffffffffu : 0u;
  const uint32_t m_tail = (lane == 1 || lane == 17) ? 0xffffffffu : 0u;
  const int red_slot = (lane & 3) == 0 ? ((lane >> 2) & 7) : lane == 1 ? 8 : (DEPTH && lane == 17) ? 9 : -1;
  // instances [0, todo) of this tile's list can matter to this warp's 64 pixels
  const uint32_t todo = __reduce_max_sync(0xffffffffu, max(S.lastc0, S.lastc1));
  const int nchunks = (int)((todo + 31u) >> 5);

  ChunkRegs nx;
  nx.r0 = nx.r1 = nx.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
  nx.gid = 0;
  uint32_t gid2 = 0;
#if S360_ASYNC_STAGE
  __shared__ float4 s_rec[NWARPS][2][3][32];
  int rbuf = 0;
#endif
  if (nchunks > 0) {
    const uint32_t p = (uint32_t)(nchunks - 1) * 32u + lane;
    if (p < todo) nx.gid = point_list[range.x + p];
#if S360_ASYNC_STAGE
    issue_records(s_rec[warp][0], nx.gid, rec, p < todo, lane);
#elif S360_BWD_PREFETCH
    load_records(nx, rec, p < todo);
#endif
    if (nchunks > 1) gid2 = point_list[range.x + p - 32u];
  }
#if S360_COUNTERS
  unsigned long long c_chunks = 0, c_tests = 0, c_surv = 0, c_hit = 0, c_pairs = 0;
#endif
  for (int ci = nchunks - 1; ci >= 0; --ci) {
    const uint32_t pos0 = (uint32_t)ci * 32u;
    const bool valid = pos0 + lane < todo;
    S360_COUNT(c_chunks, 1); S360_COUNT(c_tests, __popc(__ballot_sync(0xffffffffu, valid)));
#if S360_ASYNC_STAGE
    fetch_records(nx, s_rec[warp][rbuf], lane);
#elif !S360_BWD_PREFETCH
    load_records(nx, rec, valid);
#endif
    float4 cull, ev, col;
    stage_instance(nx.r0, nx.r1, nx.r2, cull, ev, col);
    const float thr = col.w;
    const bool huge = MODE == S360_MODE_ERP && record_is_wide(nx.r1, nx.r2, halfW);
    col.w = __uint_as_float(nx.gid);
    const float nx_depth = nx.r2.w;      // sort depth of this lane's instance (fused depth channel)
    nx.gid = gid2;                       // chunks below the last one are always full
#if S360_ASYNC_STAGE
    rbuf ^= 1;
    issue_records(s_rec[warp][rbuf], nx.gid, rec, ci > 0, lane);
#elif S360_BWD_PREFETCH
    load_records(nx, rec, ci > 0);
#endif
    gid2 = (ci > 1) ? point_list[range.x + pos0 - 64u + lane] : 0u;

    const float ddx = wrap_dx<MODE>(cull.x - wcx, Wf, halfW), ddy = cull.y - wcy;
    const bool hit = valid && (huge || rect_can_contribute(cull, ev, thr, ddx, ddy));
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      // compacted back-to-front: slot = number of surviving lanes above this one
      const int slot = __popc(mask & ~((2u << lane) - 1u));
      s_sv[warp][0][slot] = ev;
      s_sv[warp][1][slot] = col;
      s_sv[warp][2][slot] = make_float4(ddx, ddy, DEPTH ? depth_value(dspec, nx_depth) : 0.f, __uint_as_float(pos0 + (uint32_t)lane));
    }
    const bool any_wide = MODE == S360_MODE_ERP && __any_sync(0xffffffffu, hit && huge);
    __syncwarp();
    const int nsv = __popc(mask);
    S360_COUNT(c_surv, nsv);
    const float4* sv = &s_sv[warp][0][0];
    auto replay = [&](auto wide_tag) {
    constexpr bool WIDE = decltype(wide_tag)::value;
    for (int k = 0; k < nsv; k++) {
      const float4 e = sv[k];
      const float4 g = sv[64 + k];
      const uint32_t pos = __float_as_uint(g.w);
      float dx = g.x - offx;
      if (WIDE) dx = wrap_dx<MODE>(dx, Wf, halfW);
      const float2 dy = __fadd2_rn(make_float2(g.y, g.y), npy);
      const float u = e.y * dx, pb = e.x * dx * dx;
      const float2 p = __ffma2_rn(__ffma2_rn(make_float2(e.z, e.z), dy, make_float2(u, u)), dy, make_float2(pb, pb));
      // same expressions as the forward pass, so that T / (1 - alpha) undoes exactly what it applied
      const float2 a = __fadd2_rn(p, make_float2(e.w, e.w));
      const float e0 = ex2_approx(a.x), e1 = ex2_approx(a.y);                             // unclamped alpha = o G
      const float nal0 = fmaxf(-ALPHA_MAX, -e0), nal1 = fmaxf(-ALPHA_MAX, -e1);
      const bool ok0 = (pos < S.lastc0) && (p.x <= 0.f) && (-nal0 >= ALPHA_MIN);
      const bool ok1 = (pos < S.lastc1) && (p.y <= 0.f) && (-nal1 >= ALPHA_MIN);
      if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;
      S360_COUNT(c_hit, 1); S360_COUNT(c_pairs, __popc(__ballot_sync(0xffffffffu, ok0)) + __popc(__ballot_sync(0xffffffffu, ok1)));
      const float4 c = sv[32 + k];
      // A pixel that does not take this instance runs the same recurrences with alpha = 0: T and B stay as they are.
      const float2 nae = make_float2(ok0 ? nal0 : 0.f, ok1 ? nal1 : 0.f);                  // -alpha or 0
      const float2 om = __fadd2_rn(make_float2(1.f, 1.f), nae);                            // 1 - alpha
      const float2 inv = make_float2(rcp_approx(om.x), rcp_approx(om.y));
      S.T = __fmul2_rn(S.T, inv);
      const float2 nw = __fmul2_rn(nae, S.T);                                              // -alpha T
      // w = sum_c c_c dL/dC_c ;  dL/dalpha = T w - B / (1 - alpha) ;  then B += w alpha T  (nB = -B, nw = -alpha T)
      float2 w = __fmul2_rn(make_float2(c.x, c.x), S.dp0);
      w = __ffma2_rn(make_float2(c.y, c.y), S.dp1, w);
      w = __ffma2_rn(make_float2(c.z, c.z), S.dp2, w);
      if (DEPTH) w = __ffma2_rn(make_float2(g.z, g.z), dpd, w);
      const float2 dLda = __ffma2_rn(S.T, w, __fmul2_rn(S.nB, inv));
      S.nB = __ffma2_rn(w, nw, S.nB);
#if S360_BWD_QPRIME
      float2 q = __fmul2_rn(make_float2(e0, e1), dLda);   // o G dL/dalpha: the opacity factor is divided out once per Gaussian in K8
#else
      float2 q = __fmul2_rn(make_float2(ex2_approx(p.x), ex2_approx(p.y)), dLda);          // G dL/dalpha
#endif
      q.x = ok0 ? q.x : 0.f; q.y = ok1 ? q.y : 0.f;
      // the lane's two pixels share dx: sum them first, then apply the common factor
      const float2 qy = __fmul2_rn(q, dy);
      const float2 t0 = __fmul2_rn(nw, S.dp0), t1 = __fmul2_rn(nw, S.dp1), t2 = __fmul2_rn(nw, S.dp2);
      const float2 t7 = __fmul2_rn(qy, dy);
      float v[NACC];
      v[0] = -(t0.x + t0.y); v[1] = -(t1.x + t1.y); v[2] = -(t2.x + t2.y);
      v[8] = q.x + q.y;
      v[4] = qy.x + qy.y;
      v[3] = v[8] * dx; v[5] = v[3] * dx; v[6] = v[4] * dx; v[7] = t7.x + t7.y;
      const float s8 = warp_reduce8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], m16, m8, m4);
      float v8 = v[8];
      if (DEPTH) {
        // tenth sum: dL/d(depth value) = sum alpha T dL/dD; shares the butterfly of the ninth (upper half-warp)
        const float2 t3 = __fmul2_rn(nw, dpd);
        const float v9 = -(t3.x + t3.y);
        v8 = bsel(v9, v8, m16) + __shfl_xor_sync(0xffffffffu, bsel(v8, v9, m16), 16);
      } else
      v8 += __shfl_xor_sync(0xffffffffu, v8, 16);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 8);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 4);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 2);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
      // nine moment sums go straight to the per-Gaussian accumulator (fire-and-forget RED);
      // lanes 0,4,..,28 hold sums 0..7, lane 1 holds sum 8.  Conversion to gradients happens once per
      // Gaussian in preprocess_backward_kernel.
      float* dst = acc + (size_t)__float_as_uint(c.w) * ACC_STRIDE;
      if (red_slot >= 0) atomicAdd(dst + red_slot, bsel(v8, s8, m_tail));   // one RED per role lane
    }
    };
    if (any_wide) replay(std::true_type{}); else replay(std::false_type{});
    __syncwarp();
  }
#if S360_COUNTERS
  if (lane == 0) {
    atomicAdd(&g_counters[8], c_chunks); atomicAdd(&g_counters[9], c_tests); atomicAdd(&g_counters[10], c_surv);
    atomicAdd(&g_counters[11], c_hit); atomicAdd(&g_counters[12], c_pairs);
  }
#endif
}

int launch_render_backward(const S360View& v, int NV, GeomState g, const uint32_t* point_list, ImageState img,
                           const float* dL_dcolor, const float* dL_ddepth, int depth_mode, float depth_near,
                           float depth_far, float* acc, cudaStream_t st) {
  const int W = v.image_width, H = v.image_height;
  const int tiles = NV * ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  if (tiles == 0) return 0;
  DepthSpec ds;
  ds.mode = depth_mode; ds.inv_scale = 1.f / v.scene_scale; ds.near = depth_near; ds.far = depth_far;
#define S360_LAUNCH_BWD(MODE_, DEPTH_) launch_pdl(render_backward_kernel<MODE_, DEPTH_>, dim3(tiles), dim3(RT), 0, st, \
      W, H, v.bg, g.rec, point_list, img.ranges, img.order_bwd, img.final_T, img.n_contrib, dL_dcolor, dL_ddepth, ds, acc)
  if (v.mode == S360_MODE_PINHOLE) { if (dL_ddepth) S360_LAUNCH_BWD(S360_MODE_PINHOLE, true); else S360_LAUNCH_BWD(S360_MODE_PINHOLE, false); }
  else { if (dL_ddepth) S360_LAUNCH_BWD(S360_MODE_ERP, true); else S360_LAUNCH_BWD(S360_MODE_ERP, false); }
#undef S360_LAUNCH_BWD
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace s360
