// api.cu -- extern "C" entry points of libsplatter360.so (see include/splatter360.h).
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "render_cull.cuh"

namespace s360 {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("S360_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// ---- optional per-stage CUDA-event timing (off by default; bench.py turns it on) -------------
static std::atomic<int> g_profile{0};
static std::mutex g_prof_mu;
struct ProfRec { int stage; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
struct StageTimer {
  int stage; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; bool on;
  StageTimer(int stage_, cudaStream_t st_) : stage(stage_), st(st_), on(g_profile.load(std::memory_order_relaxed) != 0) {
    if (!on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    auto get = [&]() { cudaEvent_t e; if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    a = get(); b = get();
    cudaEventRecord(a, st);
  }
  ~StageTimer() {
    if (!on) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back({stage, a, b});
  }
};

// scratch layout of s360_forward_preprocess: [depth keys a | ids a | depth keys b | ids b | radix | block sums]
struct PreScratch {
  uint32_t *keys_a, *ids_a, *keys_b, *ids_b;
  void* radix;
  uint32_t* block_sums;
  uint32_t* k1_status;   // batched path: [ticket][status per K1 block]
};
// n = items sorted (P Gaussians, or the pair capacity of the batched path); P_status = Gaussians of the batched K1
static size_t pre_scratch_carve(void* buf, int64_t n, PreScratch* out, int P_status = 0) {
  char* p = (char*)buf;
  const size_t arr = align_up((size_t)(n > 0 ? n : 1) * 4, 256);
  const size_t rad = radix_scratch_bytes(n);
  const size_t sums = align_up(((size_t)(n > 0 ? n : 1) / 2048 + 4) * 4, 256);
  const size_t k1 = P_status > 0 ? align_up(((size_t)P_status / 128 + 4) * 4, 256) : 0;
  if (out) {
    out->keys_a = (uint32_t*)p; out->ids_a = (uint32_t*)(p + arr);
    out->keys_b = (uint32_t*)(p + 2 * arr); out->ids_b = (uint32_t*)(p + 3 * arr);
    out->radix = p + 4 * arr;
    out->block_sums = (uint32_t*)(p + 4 * arr + rad);
    out->k1_status = (uint32_t*)(p + 4 * arr + rad + sums);
  }
  return 4 * arr + rad + sums + k1;
}

// scratch layout of s360_forward_render: [tile keys a | tile keys b | vals b | tile counts | radix]
struct BinScratch {
  uint32_t *keys_a, *keys_b, *vals_b, *tile_count;
  void* radix;
};
// S360_FORCE_RADIX_BINNING=1 (environment, read once): always take the emit + radix-sort path (tests, A/B)
static bool force_radix() {
  static const bool f = [] { const char* e = getenv("S360_FORCE_RADIX_BINNING"); return e && e[0] == '1'; }();
  return f;
}
static int64_t tiles_of(int H, int W, int V = 1) { return (int64_t)V * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE); }
// n_items: Gaussians (or pair capacity) the binning stage walks
static bool use_matrix(int64_t n_items, int H, int W, int V = 1) {
  return !force_radix() && matrix_binning_ok(n_items, tiles_of(H, W, V), (W + TILE - 1) / TILE);
}

static size_t bin_scratch_carve(void* buf, int64_t cap, int H, int W, BinScratch* out, int V = 1) {
  char* p = (char*)buf;
  const size_t arr = align_up((size_t)(cap > 0 ? cap : 1) * 4, 256);
  const size_t tiles = (size_t)V * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
  const size_t tc = align_up((tiles > 0 ? tiles : 1) * 4 * (size_t)tile_hist_copies(), 256);
  const size_t rad = radix_scratch_bytes(cap);
  if (out) {
    out->keys_a = (uint32_t*)p; out->keys_b = (uint32_t*)(p + arr); out->vals_b = (uint32_t*)(p + 2 * arr);
    out->tile_count = (uint32_t*)(p + 3 * arr);
    out->radix = p + 3 * arr + tc;
  }
  return 3 * arr + tc + rad;
}

static int tile_bits(int H, int W, int V = 1) {
  const int tiles = V * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
  int b = 0;
  while ((1 << b) < tiles) b++;
  return b;
}

static bool view_ok(const S360View* v) {
  return v && v->P >= 0 && v->image_height >= 0 && v->image_width >= 0 &&
         (v->mode == S360_MODE_PINHOLE || v->mode == S360_MODE_ERP) && v->scene_scale > 0.f &&
         (v->sh_layout == 0 || v->sh_layout == 1) && (v->cov_layout == 0 || v->cov_layout == 1) && v->viewmatrix && v->campos && v->bg &&
         (v->mode == S360_MODE_ERP || v->projmatrix) &&
         (v->mode != S360_MODE_ERP || v->image_width % TILE == 0) &&
         ((v->image_width + TILE - 1) / TILE) < 32768 && ((v->image_height + TILE - 1) / TILE) < 65536;
}
static bool batch_ok(const S360View* v, int V, int64_t pair_capacity) {
  return view_ok(v) && V >= 1 && V <= S360_MAX_VIEWS && pair_capacity >= 0 && pair_capacity < (1ll << 30) &&
         (int64_t)V * ((v->image_height + TILE - 1) / TILE) < 65536 &&
         (int64_t)V * ((v->image_height + TILE - 1) / TILE) * ((v->image_width + TILE - 1) / TILE) < (1ll << 24);
}

// ---- stage bodies shared by the single-view and the batched entry points -----------------------------------
// n_items entries of the (Gaussian- or pair-indexed) geometry state; n_dev optionally overrides the count on device
static int order_impl(int64_t n_items, const uint32_t* n_dev, GeomState g, const PreScratch& s, uint32_t* depth_order,
                      uint32_t* inst_offsets, S360Counters* counters, bool matrix, cudaStream_t st) {
  int rc, in_b = 0;
  { StageTimer t(S360_STAGE_DEPTH_SORT, st);
    // the last of the four passes writes the sorted ids straight into depth_order.  (Accumulating the digit histograms
    // inside K1 instead of the histogram kernel was measured and rejected: K1 0.077 -> 0.119 ms for 0.015 ms less sort.)
    rc = radix_sort_pairs(s.keys_a, s.ids_a, s.keys_b, s.ids_b, n_items, n_dev, 32, s.radix, st, &in_b, false, depth_order);
    if (rc) return rc; }
  if (matrix) return 0;   // matrix binning needs no per-Gaussian offsets; K1 already counted the instances
  StageTimer t(S360_STAGE_SCAN, st);
  return launch_scan_offsets(n_items, n_dev, g, depth_order, inst_offsets, counters, s.block_sums, st);
}

static int render_impl(const S360View* view, int V, int64_t n_items, const uint32_t* n_dev, GeomState g,
                       const uint32_t* depth_order, const uint32_t* inst_offsets, S360Counters* counters,
                       int64_t instance_capacity, uint32_t* point_list, void* image_state, float* out_color,
                       float* out_depth, int32_t depth_mode, float depth_near, float depth_far, void* scratch,
                       cudaStream_t st) {
  const int H = view->image_height, W = view->image_width;
  if (depth_mode < S360_DEPTH_DEPTH || depth_mode > S360_DEPTH_LOG) return S360_ERR_BAD_ARGUMENT;
  ImageState img = carve_image(image_state, H, W, V);
  if (use_matrix(n_items, H, W, V)) {
    // matrix binning (binning.cu): count per (chunk, tile) -> column scan -> tile ranges -> ranked scatter
    int rc;
    { StageTimer t(S360_STAGE_EMIT, st);
      rc = launch_mb_count(*view, V, n_items, n_dev, g, depth_order, scratch, st); }
    if (rc) return rc;
    { StageTimer t(S360_STAGE_SCAN, st);
      rc = launch_mb_colscan(*view, V, n_items, n_dev, scratch, st); }
    if (rc) return rc;
    { StageTimer t(S360_STAGE_TILE_RANGES, st);
      rc = launch_tile_scan(*view, V, mb_tile_totals(scratch, n_items, (int)tiles_of(H, W, V)), 1, instance_capacity,
                            img.ranges, img.order, img.work, nullptr, 0, st); }
    if (rc) return rc;
    { StageTimer t(S360_STAGE_TILE_SORT, st);
      rc = launch_mb_scatter(*view, V, n_items, n_dev, g, depth_order, counters, instance_capacity, point_list, img.ranges,
                             scratch, st); }
    if (rc) return rc;
    StageTimer t(S360_STAGE_RENDER_FWD, st);
    return launch_render_forward(*view, V, g, point_list, img, out_color, out_depth, depth_mode, depth_near, depth_far, st);
  }
  BinScratch s;
  bin_scratch_carve(scratch, instance_capacity, H, W, &s, V);
  const int nbits = tile_bits(H, W, V);
  const int passes = (nbits + 7) / 8;
  if (passes > 3) return S360_ERR_UNSUPPORTED;
  // the sorted ids must land in point_list: start in (keys_a, point_list) for an even number of passes,
  // in (keys_b, vals_b) for an odd number
  uint32_t *k0 = s.keys_a, *v0 = point_list, *k1 = s.keys_b, *v1 = s.vals_b;
  if (passes & 1) { k0 = s.keys_b; v0 = s.vals_b; k1 = s.keys_a; v1 = point_list; }
  int rc;
  uint32_t* hist;
  { StageTimer t(S360_STAGE_EMIT, st);
    hist = radix_prepare_hist(s.radix, instance_capacity, nbits, st);
    rc = launch_emit(*view, V, n_items, n_dev, g, depth_order, inst_offsets, counters, instance_capacity, k0, v0, s.tile_count, st); }
  if (rc) return rc;
  { StageTimer t(S360_STAGE_TILE_RANGES, st);
    rc = launch_tile_scan(*view, V, s.tile_count, tile_hist_copies(), instance_capacity, img.ranges, img.order, img.work, hist,
                          passes > 0 ? passes : 1, st); }
  if (rc) return rc;
  int in_b = 0;
  { StageTimer t(S360_STAGE_TILE_SORT, st);
    rc = radix_sort_pairs(k0, v0, k1, v1, instance_capacity, &counters->num_rendered, nbits, s.radix, st, &in_b, true); }
  if (rc) return rc;
  StageTimer t(S360_STAGE_RENDER_FWD, st);
  return launch_render_forward(*view, V, g, point_list, img, out_color, out_depth, depth_mode, depth_near, depth_far, st);
}
}  // namespace s360

using namespace s360;

extern "C" {

int s360_abi_version(void) { return S360_ABI_VERSION; }

int s360_profile_enable(int on) { return g_profile.exchange(on ? 1 : 0); }

int s360_profile_read(double* ms, uint64_t* counts, int reset) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < S360_NUM_STAGES; i++) { if (ms) ms[i] = 0.0; if (counts) counts[i] = 0; }
  int rc = 0;
  for (auto& r : g_prof_recs) {
    float t = 0.f;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) { rc = (int)e; continue; }
    if (r.stage >= 0 && r.stage < S360_NUM_STAGES) { if (ms) ms[r.stage] += (double)t; if (counts) counts[r.stage] += 1; }
  }
  if (reset) {
    for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof_recs.clear();
  }
  return rc;
}

int s360_debug_counters(uint64_t* out, int reset, void* stream) {
  if (!out) return S360_ERR_BAD_ARGUMENT;
  return read_counters((unsigned long long*)out, reset, (cudaStream_t)stream);
}

uint64_t s360_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* s360_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case S360_ERR_BAD_ARGUMENT: return "splatter360: bad argument";
    case S360_ERR_WORKSPACE_OVERFLOW: return "splatter360: instance workspace overflow";
    case S360_ERR_UNSUPPORTED: return "splatter360: unsupported configuration";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "splatter360: unknown error";
  }
}

size_t s360_geom_bytes(int32_t P) { return geom_bytes(P > 0 ? P : 1); }
size_t s360_preprocess_scratch_bytes(int32_t P) { return pre_scratch_carve(nullptr, P, nullptr); }
size_t s360_binning_scratch_bytes(int32_t P, int64_t cap, int32_t H, int32_t W) {
  if (use_matrix(P, H, W)) return matrix_scratch_bytes(P, tiles_of(H, W));
  return bin_scratch_carve(nullptr, cap, H, W, nullptr);
}
size_t s360_image_bytes(int32_t H, int32_t W) { return image_bytes(H, W); }
size_t s360_backward_scratch_bytes(int32_t P) { return align_up((size_t)(P > 0 ? P : 1) * ACC_STRIDE * sizeof(float), 256); }

int s360_forward_project(const S360View* view, const float* means3D, const float* cov3D, const float* opacities,
                         const float* shs, const float* colors_precomp, void* geom, int32_t* radii,
                         S360Counters* counters, void* scratch, void* stream) {
  if (!view_ok(view) || !counters || !geom || !scratch) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (shs == nullptr) == (colors_precomp == nullptr)) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (!means3D || !cov3D || !opacities || !radii)) return S360_ERR_BAD_ARGUMENT;
  if (shs && view->M < (view->sh_degree < view->max_sh_degree ? (view->sh_degree + 1) * (view->sh_degree + 1)
                                                                : (view->max_sh_degree + 1) * (view->max_sh_degree + 1)))
    return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P;
  int rc = (int)cudaMemsetAsync(counters, 0, sizeof(S360Counters), st);
  if (rc) return rc;
  GeomState g = carve_geom(geom, P > 0 ? P : 1);
  PreScratch s;
  pre_scratch_carve(scratch, P, &s);
  StageTimer t(S360_STAGE_PREPROCESS, st);
  return launch_preprocess(*view, means3D, cov3D, opacities, shs, colors_precomp, g, radii, s.keys_a, s.ids_a, counters, st);
}

int s360_forward_order(const S360View* view, const void* geom, uint32_t* depth_order, uint32_t* inst_offsets,
                       S360Counters* counters, void* scratch, void* stream) {
  if (!view_ok(view) || !counters || !geom || !scratch) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (!depth_order || !inst_offsets)) return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P;
  GeomState g = carve_geom(const_cast<void*>(geom), P > 0 ? P : 1);
  PreScratch s;
  pre_scratch_carve(scratch, P, &s);
  return order_impl(P, nullptr, g, s, depth_order, inst_offsets, counters, use_matrix(P, view->image_height, view->image_width), st);
}

int s360_forward_preprocess(const S360View* view, const float* means3D, const float* cov3D, const float* opacities,
                            const float* shs, const float* colors_precomp, void* geom, int32_t* radii,
                            uint32_t* depth_order, uint32_t* inst_offsets, S360Counters* counters, void* scratch,
                            void* stream) {
  int rc = s360_forward_project(view, means3D, cov3D, opacities, shs, colors_precomp, geom, radii, counters, scratch, stream);
  if (rc) return rc;
  return s360_forward_order(view, geom, depth_order, inst_offsets, counters, scratch, stream);
}

int s360_forward_render(const S360View* view, const void* geom, const uint32_t* depth_order,
                        const uint32_t* inst_offsets, S360Counters* counters, int64_t instance_capacity,
                        uint32_t* point_list, void* image_state, float* out_color, float* out_depth,
                        int32_t depth_mode, float depth_near, float depth_far, void* scratch, void* stream) {
  if (!view_ok(view) || !geom || !counters || !image_state || !scratch || instance_capacity < 0) return S360_ERR_BAD_ARGUMENT;
  if (instance_capacity > 0 && !point_list) return S360_ERR_BAD_ARGUMENT;
  if ((size_t)view->image_height * view->image_width > 0 && !out_color) return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P;
  GeomState g = carve_geom(const_cast<void*>(geom), P > 0 ? P : 1);
  return render_impl(view, 1, P, nullptr, g, depth_order, inst_offsets, counters, instance_capacity, point_list, image_state,
                     out_color, out_depth, depth_mode, depth_near, depth_far, scratch, st);
}

int s360_backward(const S360View* view, const float* means3D, const float* cov3D, const float* opacities,
                  const float* shs, const float* colors_precomp, const void* geom, const int32_t* radii,
                  const uint32_t* point_list, const void* image_state, const float* dL_dcolor,
                  const float* dL_ddepth, int32_t depth_mode, float depth_near, float depth_far,
                  float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dcov3D, float* dL_dopacity, float* dL_dshs,
                  float* dL_dcolors, void* scratch, void* stream) {
  (void)colors_precomp;
  if (!view_ok(view) || !geom || !image_state || !scratch) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (!means3D || !cov3D || !opacities || !radii || !dL_dmeans3D || !dL_dmeans2D || !dL_dcov3D || !dL_dopacity)) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && shs && !dL_dshs) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && !shs && !dL_dcolors) return S360_ERR_BAD_ARGUMENT;
  if ((size_t)view->image_height * view->image_width > 0 && !dL_dcolor) return S360_ERR_BAD_ARGUMENT;
  if (dL_ddepth && (depth_mode < S360_DEPTH_DEPTH || depth_mode > S360_DEPTH_LOG)) return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P, H = view->image_height, W = view->image_width;
  GeomState g = carve_geom(const_cast<void*>(geom), P > 0 ? P : 1);
  ImageState img = carve_image(const_cast<void*>(image_state), H, W);
  float* acc = (float*)scratch;
  int rc = (int)cudaMemsetAsync(acc, 0, (size_t)(P > 0 ? P : 1) * ACC_STRIDE * sizeof(float), st);
  if (rc) return rc;
  if (P > 0) {
    StageTimer t(S360_STAGE_RENDER_BWD, st);
    rc = launch_tile_order(*view, 1, img.work, img.order_bwd, st);
    if (rc) return rc;
    rc = launch_render_backward(*view, 1, g, point_list, img, dL_dcolor, dL_ddepth, depth_mode, depth_near, depth_far, acc, st);
    if (rc) return rc;
  }
  StageTimer t(S360_STAGE_PREPROCESS_BWD, st);
  return launch_preprocess_backward(*view, means3D, cov3D, opacities, shs, g, radii, acc, dL_dmeans3D, dL_dmeans2D, dL_dcov3D,
                                    dL_dopacity, dL_dshs, dL_dcolors, dL_ddepth != nullptr, depth_mode, depth_near, depth_far, st);
}

// ---- batched multi-view path ---------------------------------------------------------------------------------
size_t s360_multi_geom_bytes(int32_t P, int64_t pair_capacity) { return geom_multi_bytes(P > 0 ? P : 1, pair_capacity > 0 ? pair_capacity : 1); }
size_t s360_multi_preprocess_scratch_bytes(int32_t P, int64_t pair_capacity) { return pre_scratch_carve(nullptr, pair_capacity, nullptr, P > 0 ? P : 1); }
size_t s360_multi_binning_scratch_bytes(int64_t pair_capacity, int64_t cap, int32_t V, int32_t H, int32_t W) {
  if (use_matrix(pair_capacity, H, W, V > 0 ? V : 1)) return matrix_scratch_bytes(pair_capacity, tiles_of(H, W, V > 0 ? V : 1));
  return bin_scratch_carve(nullptr, cap, H, W, nullptr, V > 0 ? V : 1);
}
size_t s360_multi_image_bytes(int32_t V, int32_t H, int32_t W) { return image_bytes(H, W, V > 0 ? V : 1); }
size_t s360_multi_backward_scratch_bytes(int64_t pair_capacity) { return align_up((size_t)(pair_capacity > 0 ? pair_capacity : 1) * ACC_STRIDE * sizeof(float), 256); }

int s360_multi_forward_project(const S360View* view, int32_t V, int64_t pair_capacity, const float* means3D,
                               const float* cov3D, const float* opacities, const float* shs,
                               const float* colors_precomp, void* geom, int32_t* radii, S360Counters* counters,
                               void* scratch, void* stream) {
  if (!batch_ok(view, V, pair_capacity) || !counters || !geom || !scratch) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (shs == nullptr) == (colors_precomp == nullptr)) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (!means3D || !cov3D || !opacities)) return S360_ERR_BAD_ARGUMENT;
  if (shs && view->M < (view->sh_degree < view->max_sh_degree ? (view->sh_degree + 1) * (view->sh_degree + 1)
                                                                : (view->max_sh_degree + 1) * (view->max_sh_degree + 1)))
    return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P;
  const int64_t cap = pair_capacity > 0 ? pair_capacity : 1;
  int rc = (int)cudaMemsetAsync(counters, 0, sizeof(S360Counters), st);
  if (rc) return rc;
  PairState ps;
  GeomState g = carve_geom_multi(geom, P > 0 ? P : 1, cap, &ps);
  PreScratch s;
  pre_scratch_carve(scratch, pair_capacity, &s, P > 0 ? P : 1);
  StageTimer t(S360_STAGE_PREPROCESS, st);
  return launch_preprocess_multi(*view, V, pair_capacity, means3D, cov3D, opacities, shs, colors_precomp, g, ps, radii,
                                 s.keys_a, s.ids_a, counters, s.k1_status, st);
}

int s360_multi_forward_order(const S360View* view, int32_t V, int64_t pair_capacity, const void* geom,
                             uint32_t* depth_order, uint32_t* inst_offsets, S360Counters* counters, void* scratch,
                             void* stream) {
  if (!batch_ok(view, V, pair_capacity) || !counters || !geom || !scratch) return S360_ERR_BAD_ARGUMENT;
  if (pair_capacity > 0 && (!depth_order || !inst_offsets)) return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P;
  PairState ps;
  GeomState g = carve_geom_multi(const_cast<void*>(geom), P > 0 ? P : 1, pair_capacity > 0 ? pair_capacity : 1, &ps);
  PreScratch s;
  pre_scratch_carve(scratch, pair_capacity, &s, P > 0 ? P : 1);
  return order_impl(pair_capacity, &counters->num_visible, g, s, depth_order, inst_offsets, counters,
                    use_matrix(pair_capacity, view->image_height, view->image_width, V), st);
}

int s360_multi_forward_render(const S360View* view, int32_t V, int64_t pair_capacity, const void* geom,
                              const uint32_t* depth_order, const uint32_t* inst_offsets, S360Counters* counters,
                              int64_t instance_capacity, uint32_t* point_list, void* image_state, float* out_color,
                              float* out_depth, int32_t depth_mode, float depth_near, float depth_far, void* scratch,
                              void* stream) {
  if (!batch_ok(view, V, pair_capacity) || !geom || !counters || !image_state || !scratch || instance_capacity < 0) return S360_ERR_BAD_ARGUMENT;
  if (instance_capacity > 0 && !point_list) return S360_ERR_BAD_ARGUMENT;
  if ((size_t)view->image_height * view->image_width > 0 && !out_color) return S360_ERR_BAD_ARGUMENT;
  const int P = view->P;
  PairState ps;
  GeomState g = carve_geom_multi(const_cast<void*>(geom), P > 0 ? P : 1, pair_capacity > 0 ? pair_capacity : 1, &ps);
  return render_impl(view, V, pair_capacity, &counters->num_visible, g, depth_order, inst_offsets, counters, instance_capacity,
                     point_list, image_state, out_color, out_depth, depth_mode, depth_near, depth_far, scratch, (cudaStream_t)stream);
}

int s360_multi_backward(const S360View* view, int32_t V, int64_t pair_capacity, const float* means3D,
                        const float* cov3D, const float* opacities, const float* shs, const float* colors_precomp,
                        const void* geom, const uint32_t* point_list, const void* image_state, const float* dL_dcolor,
                        const float* dL_ddepth, int32_t depth_mode, float depth_near, float depth_far,
                        float* dL_dmeans3D, float* dL_dcov3D, float* dL_dopacity, float* dL_dshs, float* dL_dcolors,
                        void* scratch, void* stream) {
  (void)colors_precomp;
  if (!batch_ok(view, V, pair_capacity) || !geom || !image_state || !scratch) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && (!means3D || !cov3D || !opacities || !dL_dmeans3D || !dL_dcov3D || !dL_dopacity)) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && shs && !dL_dshs) return S360_ERR_BAD_ARGUMENT;
  if (view->P > 0 && !shs && !dL_dcolors) return S360_ERR_BAD_ARGUMENT;
  if ((size_t)view->image_height * view->image_width > 0 && !dL_dcolor) return S360_ERR_BAD_ARGUMENT;
  if (dL_ddepth && (depth_mode < S360_DEPTH_DEPTH || depth_mode > S360_DEPTH_LOG)) return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = view->P, H = view->image_height, W = view->image_width;
  const int64_t cap = pair_capacity > 0 ? pair_capacity : 1;
  PairState ps;
  GeomState g = carve_geom_multi(const_cast<void*>(geom), P > 0 ? P : 1, cap, &ps);
  ImageState img = carve_image(const_cast<void*>(image_state), H, W, V);
  float* acc = (float*)scratch;
  int rc = 0;
  if (P > 0) {
    StageTimer t(S360_STAGE_RENDER_BWD, st);
    rc = launch_zero_acc(acc, ps.count, cap, st);
    if (rc) return rc;
    rc = launch_tile_order(*view, V, img.work, img.order_bwd, st);
    if (rc) return rc;
    rc = launch_render_backward(*view, V, g, point_list, img, dL_dcolor, dL_ddepth, depth_mode, depth_near, depth_far, acc, st);
    if (rc) return rc;
  }
  StageTimer t(S360_STAGE_PREPROCESS_BWD, st);
  return launch_preprocess_multi_backward(*view, V, means3D, cov3D, opacities, shs, g, ps, acc, dL_dmeans3D, dL_dcov3D,
                                          dL_dopacity, dL_dshs, dL_dcolors, dL_ddepth != nullptr, depth_mode, depth_near, depth_far, st);
}

int s360_mark_visible(const S360View* view, const float* means3D, uint8_t* present, void* stream) {
  if (!view_ok(view) || (view->P > 0 && (!means3D || !present))) return S360_ERR_BAD_ARGUMENT;
  return launch_mark_visible(*view, means3D, present, (cudaStream_t)stream);
}

// ---- debug unpackers ------------------------------------------------------------------------
__global__ void unpack_geom_kernel(int P, GeomState g, float* xy, float* depth, float* conop, float* rgb,
                                   uint32_t* tiles, uint8_t* clamped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint2 r = g.rect[i];
  const uint32_t t = (r.x >> 16) * (r.y >> 16);
  const float4 a = g.rec[3 * (size_t)i], b = g.rec[3 * (size_t)i + 1], c = g.rec[3 * (size_t)i + 2];
  if (tiles) tiles[i] = t;
  if (xy) { xy[2 * i] = a.x; xy[2 * i + 1] = a.y; }
  if (depth) depth[i] = c.w;
  float co[4], col[3];
  unpack_record(a, b, c, co, col);
  if (conop) { conop[4 * i] = co[0]; conop[4 * i + 1] = co[1]; conop[4 * i + 2] = co[2]; conop[4 * i + 3] = co[3]; }
  if (rgb) { rgb[3 * i] = col[0]; rgb[3 * i + 1] = col[1]; rgb[3 * i + 2] = col[2]; }
  if (clamped) { const uint8_t cl = g.clamped[i]; clamped[3 * i] = cl & 1; clamped[3 * i + 1] = (cl >> 1) & 1; clamped[3 * i + 2] = (cl >> 2) & 1; }
}

int s360_debug_unpack_geom(int32_t P, const void* geom, float* xy, float* depth, float* conic_opacity, float* rgb,
                           uint32_t* tiles_touched, uint8_t* clamped, void* stream) {
  if (P <= 0) return 0;
  GeomState g = carve_geom(const_cast<void*>(geom), P);
  unpack_geom_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, g, xy, depth, conic_opacity, rgb, tiles_touched, clamped);
  return (int)cudaGetLastError();
}

int s360_debug_unpack_pairs(int32_t P, int64_t pair_capacity, const void* geom, uint32_t* pair_base, uint32_t* view_mask,
                            uint32_t* num_pairs, void* stream) {
  if (P <= 0 || !geom) return 0;
  PairState ps;
  carve_geom_multi(const_cast<void*>(geom), P, pair_capacity > 0 ? pair_capacity : 1, &ps);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  if (pair_base) rc = (int)cudaMemcpyAsync(pair_base, ps.base, (size_t)P * 4, cudaMemcpyDeviceToDevice, st);
  if (!rc && view_mask) rc = (int)cudaMemcpyAsync(view_mask, ps.mask, (size_t)P * 4, cudaMemcpyDeviceToDevice, st);
  if (!rc && num_pairs) rc = (int)cudaMemcpyAsync(num_pairs, ps.count, 4, cudaMemcpyDeviceToDevice, st);
  return rc;
}

int s360_debug_unpack_image(int32_t H, int32_t W, const void* image_state, float* final_T, uint32_t* n_contrib,
                            uint32_t* tile_ranges, void* stream) {
  ImageState img = carve_image(const_cast<void*>(image_state), H, W);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)H * W;
  const size_t tiles = (size_t)((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
  int rc = 0;
  if (final_T && npix) rc = (int)cudaMemcpyAsync(final_T, img.final_T, npix * 4, cudaMemcpyDeviceToDevice, st);
  if (!rc && n_contrib && npix) rc = (int)cudaMemcpyAsync(n_contrib, img.n_contrib, npix * 4, cudaMemcpyDeviceToDevice, st);
  if (!rc && tile_ranges && tiles) rc = (int)cudaMemcpyAsync(tile_ranges, img.ranges, tiles * 8, cudaMemcpyDeviceToDevice, st);
  return rc;
}

}  // extern "C"
