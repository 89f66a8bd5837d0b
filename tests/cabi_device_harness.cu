// Torch-free consumer of the C-ABI (include/yolopost_b200.h) that runs the hot path with REAL device work: plain cudaMalloc'd
// buffers, seeded synthetic head tensors, and checks that need no oracle -
//   * ypb_decode_dense + ypb_nms_from_dense  ==  ypb_nms_from_head (LDG scan)  ==  ypb_nms_from_head (TMA scan)
//     == the same with plan-owned clean-on-exit counters and counts in mapped host memory: counts, kept rows and kept anchor
//     indices bit for bit (fp32 / bf16 / fp16, predict and multi-label val mode, OBB, an odd grid that takes the scalar path);
//   * ypb_nms_boxes against a scalar restatement of torchvision.ops.nms / TorchNMS.fast_nms semantics (utils/nms.py:187-296);
//   * ypb_pairwise_iou, ypb_dfl_expectation, ypb_dist2bbox, ypb_kpts_decode, ypb_scale_rows against scalar formulas;
//   * ypb_compact_results, ypb_process_mask (both forms), ypb_match_predictions, and the one-sided result ring
//     (ypb_nms_out.peer_* + ypb_peer_wait / ypb_peer_wait_copy) looped back onto one GPU.
// It is the program the round's compute-sanitizer runs (memcheck / racecheck / synccheck) execute: no Python, no torch, so the
// tools see only this library's kernels.  Test infrastructure; built by __graft_entry__.build() next to the library.
//
//   ypb_cabi_harness [--small] [--list]        exit code 0 = every case passed
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "yolopost_b200.h"

namespace {

int g_failed = 0, g_cases = 0;
bool g_small = false;

#define CUDA_OK(x)                                                                              \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      std::printf("CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, #x); \
      std::exit(3);                                                                             \
    }                                                                                           \
  } while (0)

#define YPB_OK_OR_FAIL(x)                                                                         \
  do {                                                                                            \
    int rc_ = (x);                                                                                \
    if (rc_ != YPB_OK) {                                                                          \
      std::printf("  %s -> %d: %s (%s:%d)\n", #x, rc_, ypb_last_error_string(), __FILE__, __LINE__); \
      return false;                                                                               \
    }                                                                                             \
  } while (0)

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) { next(); next(); }
  uint64_t next() {
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    return s * 0x2545F4914F6CDD1Dull;
  }
  float uni() { return static_cast<float>((next() >> 40) + 1) * (1.0f / 16777217.0f); }  // (0, 1)
  float gauss() {
    const float u = uni(), v = uni();
    return std::sqrt(-2.0f * std::log(u)) * std::cos(6.2831853f * v);
  }
};

std::vector<void*> g_allocs;
template <class T>
T* dalloc(size_t n) {
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  g_allocs.push_back(p);
  return static_cast<T*>(p);
}
void free_all() {
  for (void* p : g_allocs) cudaFree(p);
  g_allocs.clear();
}
template <class T>
T* to_device(const std::vector<T>& v) {
  T* p = dalloc<T>(v.size());
  if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}
template <class T>
std::vector<T> to_host(const T* p, size_t n) {
  std::vector<T> v(n);
  if (n) CUDA_OK(cudaMemcpy(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
  return v;
}

size_t dtype_size(int dt) { return dt == YPB_F32 ? 4 : 2; }
float round_to(int dt, float v) {
  if (dt == YPB_BF16) return __bfloat162float(__float2bfloat16(v));
  if (dt == YPB_F16) return __half2float(__float2half(v));
  return v;
}
// float values -> device buffer of the given dtype
void* upload_as(const std::vector<float>& v, int dt) {
  if (dt == YPB_F32) return to_device(v);
  std::vector<uint16_t> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    if (dt == YPB_BF16) { __nv_bfloat16 x = __float2bfloat16(v[i]); std::memcpy(&h[i], &x, 2); }
    else { __half x = __float2half(v[i]); std::memcpy(&h[i], &x, 2); }
  }
  return to_device(h);
}
std::vector<float> download_as(const void* p, size_t n, int dt) {
  if (dt == YPB_F32) return to_host(static_cast<const float*>(p), n);
  std::vector<uint16_t> h = to_host(static_cast<const uint16_t*>(p), n);
  std::vector<float> v(n);
  for (size_t i = 0; i < n; ++i) {
    if (dt == YPB_BF16) { __nv_bfloat16 x; std::memcpy(&x, &h[i], 2); v[i] = __bfloat162float(x); }
    else { __half x; std::memcpy(&x, &h[i], 2); v[i] = __half2float(x); }
  }
  return v;
}

struct Geom {
  int nl;
  int h[YPB_MAX_LEVELS], w[YPB_MAX_LEVELS];
  float stride[YPB_MAX_LEVELS];
  int nc;
  int anchors() const { int a = 0; for (int l = 0; l < nl; ++l) a += h[l] * w[l]; return a; }
};

struct Head {
  ypb_head_desc desc;
  void* angle = nullptr;  // (B, A) logits, head dtype (OBB) or NULL
  int anchors = 0;
};

// Seeded head batch: background class logits N(mu_bg, 1.5^2), `p_obj` of the anchors carry an object (one class logit N(1, 1.5^2),
// box logits peaked at 1..11 grid cells per side so that neighbouring objects overlap and suppression has work to do).
Head make_head(const Geom& g, int batch, int dt, uint64_t seed, float mu_bg, float p_obj, bool with_angle) {
  Head hd;
  std::memset(&hd.desc, 0, sizeof(hd.desc));
  const int no = 64 + g.nc;
  hd.desc.num_levels = g.nl; hd.desc.batch = batch; hd.desc.nc = g.nc; hd.desc.reg_max = 16; hd.desc.dtype = dt;
  Rng rng(seed);
  for (int l = 0; l < g.nl; ++l) {
    const int hw = g.h[l] * g.w[l];
    std::vector<float> v(static_cast<size_t>(batch) * no * hw);
    for (int b = 0; b < batch; ++b)
      for (int px = 0; px < hw; ++px) {
        const bool obj = rng.uni() < p_obj;
        const int oc = static_cast<int>(rng.uni() * g.nc) % g.nc;
        float* base = v.data() + static_cast<size_t>(b) * no * hw + px;
        for (int side = 0; side < 4; ++side) {
          const float d = 1.0f + 10.0f * rng.uni();
          for (int k = 0; k < 16; ++k)
            base[static_cast<size_t>(side * 16 + k) * hw] = obj ? -0.5f * (k - d) * (k - d) + 0.3f * rng.gauss() : rng.gauss();
        }
        for (int c = 0; c < g.nc; ++c)
          base[static_cast<size_t>(64 + c) * hw] = (obj && c == oc) ? 1.0f + 1.5f * rng.gauss() : mu_bg + 1.5f * rng.gauss();
      }
    hd.desc.level_ptr[l] = upload_as(v, dt);
    hd.desc.level_h[l] = g.h[l]; hd.desc.level_w[l] = g.w[l];
    hd.desc.level_batch_stride[l] = static_cast<int64_t>(no) * hw;
    hd.desc.level_channel_stride[l] = hw;
    hd.desc.level_stride[l] = g.stride[l];
  }
  hd.anchors = g.anchors();
  if (with_angle) {
    std::vector<float> a(static_cast<size_t>(batch) * hd.anchors);
    for (float& x : a) x = rng.gauss();
    hd.angle = upload_as(a, dt);
  }
  return hd;
}

struct NmsCfg {
  float conf = 0.25f, iou = 0.7f, max_wh = 7680.f;
  int max_det = 300, max_nms = 30000;
  bool multi_label = false, rotated = false;
};

struct NmsBuffers {
  ypb_nms_params p;
  ypb_nms_out o;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int batch = 0, cols = 0, max_det = 0;
  int32_t* counters = nullptr;
  int32_t* count_host = nullptr;  // mapped pinned
};

NmsBuffers make_buffers(int batch, int anchors, int nc, int extra, int dt, const NmsCfg& c, bool own_counters, bool host_counts) {
  NmsBuffers nb;
  std::memset(&nb.p, 0, sizeof(nb.p));
  std::memset(&nb.o, 0, sizeof(nb.o));
  const bool ml = c.multi_label && nc > 1;
  int rows_cap = std::max(1, ml ? anchors * nc : anchors);
  int max_nms = std::max(1, std::min(c.max_nms, rows_cap));
  int max_det = std::max(1, std::min(c.max_det, max_nms));
  nb.p.conf_thres = round_to(dt, c.conf);
  nb.p.iou_thres_eff = c.iou;  // 0.7f < 0.7: already the largest float not above the double threshold
  nb.p.nc = nc; nb.p.extra = extra; nb.p.max_det = max_det; nb.p.max_nms = max_nms; nb.p.max_wh = c.max_wh;
  nb.p.multi_label = ml ? 1 : 0;
  nb.p.rule = c.rotated ? YPB_NMS_FAST_PROBIOU : YPB_NMS_GREEDY;
  nb.p.rows_cap = rows_cap;
  nb.ws_bytes = ypb_nms_workspace_bytes(batch, anchors, rows_cap, max_det, max_nms, nb.p.rule);
  nb.ws = dalloc<uint8_t>(nb.ws_bytes);
  nb.batch = batch; nb.cols = 6 + extra; nb.max_det = max_det;
  nb.o.rows = dalloc<float>(static_cast<size_t>(batch) * max_det * nb.cols);
  nb.o.idx = reinterpret_cast<int64_t*>(dalloc<long long>(static_cast<size_t>(batch) * max_det));
  nb.o.count = dalloc<int32_t>(batch);
  nb.o.cand_count = dalloc<int32_t>(batch);
  CUDA_OK(cudaMemset(nb.o.rows, 0, static_cast<size_t>(batch) * max_det * nb.cols * sizeof(float)));
  CUDA_OK(cudaMemset(nb.o.idx, 0, static_cast<size_t>(batch) * max_det * sizeof(long long)));
  CUDA_OK(cudaMemset(nb.o.count, 0, batch * sizeof(int32_t)));
  CUDA_OK(cudaMemset(nb.o.cand_count, 0, batch * sizeof(int32_t)));
  if (own_counters) {
    nb.counters = dalloc<int32_t>(batch + 1);
    CUDA_OK(cudaMemset(nb.counters, 0, (batch + 1) * sizeof(int32_t)));
    nb.p.clean_counters = nb.counters;
  }
  if (host_counts) {
    CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&nb.count_host), std::max(batch, 1) * sizeof(int32_t), cudaHostAllocMapped));
    std::memset(nb.count_host, 0xFF, std::max(batch, 1) * sizeof(int32_t));
    nb.o.count_host = nb.count_host;  // unified addressing: the mapped allocation has the same address on the device
  }
  return nb;
}

struct Result {
  std::vector<int32_t> count, cand;
  std::vector<float> rows;
  std::vector<long long> idx;
  int max_det = 0, cols = 0;
};
Result fetch(const NmsBuffers& nb) {
  CUDA_OK(cudaDeviceSynchronize());
  Result r;
  r.max_det = nb.max_det; r.cols = nb.cols;
  r.count = to_host(nb.o.count, nb.batch);
  r.cand = to_host(nb.o.cand_count, nb.batch);
  r.rows = to_host(nb.o.rows, static_cast<size_t>(nb.batch) * nb.max_det * nb.cols);
  r.idx = to_host(reinterpret_cast<const long long*>(nb.o.idx), static_cast<size_t>(nb.batch) * nb.max_det);
  return r;
}
long long total_kept(const Result& r) { return std::accumulate(r.count.begin(), r.count.end(), 0LL); }

bool same_result(const Result& a, const Result& b, const char* what, bool with_cand = false) {
  if (a.count != b.count) {
    std::printf("  %s: kept counts differ (", what);
    for (size_t i = 0; i < a.count.size(); ++i) std::printf("%d/%d ", a.count[i], b.count[i]);
    std::printf(")\n");
    return false;
  }
  if (with_cand && a.cand != b.cand) { std::printf("  %s: candidate counts differ\n", what); return false; }
  for (size_t img = 0; img < a.count.size(); ++img) {
    const int n = std::min(a.count[img], a.max_det);
    const size_t ro = img * a.max_det * a.cols, io = img * a.max_det;
    if (std::memcmp(a.rows.data() + ro, b.rows.data() + ro, static_cast<size_t>(n) * a.cols * sizeof(float)) != 0) {
      std::printf("  %s: kept rows of image %zu differ\n", what, img);
      return false;
    }
    if (std::memcmp(a.idx.data() + io, b.idx.data() + io, static_cast<size_t>(n) * sizeof(long long)) != 0) {
      std::printf("  %s: kept anchor indices of image %zu differ\n", what, img);
      return false;
    }
  }
  return true;
}

// ---- case: the fused path against the two-call path, every scan form ------------------------------------------------
bool case_fused(const char* name, const Geom& g, int batch, int dt, const NmsCfg& c, float mu_bg, float p_obj, uint64_t seed) {
  Head hd = make_head(g, batch, dt, seed, mu_bg, p_obj, c.rotated);
  const int A = hd.anchors, extra = c.rotated ? 1 : 0, ch = 4 + g.nc + extra;
  // two calls: Detect._inference, then non_max_suppression on the dense tensor
  void* dense = dalloc<uint8_t>(static_cast<size_t>(batch) * ch * A * dtype_size(dt));
  YPB_OK_OR_FAIL(ypb_decode_dense(&hd.desc, hd.angle, 1, c.rotated ? 1 : 0, 0, dense, dt, static_cast<int64_t>(ch) * A, A, nullptr));
  NmsBuffers two = make_buffers(batch, A, g.nc, extra, dt, c, false, false);
  ypb_dense_desc dd;
  std::memset(&dd, 0, sizeof(dd));
  dd.ptr = dense; dd.dtype = dt; dd.batch = batch; dd.channels = ch; dd.anchors = A;
  dd.stride_b = static_cast<int64_t>(ch) * A; dd.stride_c = A; dd.stride_a = 1;
  YPB_OK_OR_FAIL(ypb_nms_from_dense(&dd, &two.p, &two.o, two.ws, two.ws_bytes, nullptr));
  Result r2 = fetch(two);
  // dense values must be finite
  {
    std::vector<float> y = download_as(dense, static_cast<size_t>(batch) * ch * A, dt);
    for (float v : y)
      if (!std::isfinite(v)) { std::printf("  non-finite value in the dense decode\n"); return false; }
  }
  bool ok = true;
  Result first;
  struct Form { const char* what; int scan; bool own_counters, host_counts; };
  const Form forms[] = {{"fused LDG", YPB_SCAN_LDG, false, false}, {"fused TMA", YPB_SCAN_TMA, false, false},
                        {"fused LDG + clean counters + host counts", YPB_SCAN_LDG, true, true},
                        {"fused TMA + clean counters", YPB_SCAN_TMA, true, false}};
  for (const Form& f : forms) {
    NmsBuffers nb = make_buffers(batch, A, g.nc, extra, dt, c, f.own_counters, f.host_counts);
    nb.p.scan_kernel = f.scan;
    for (int rep = 0; rep < (f.own_counters ? 2 : 1); ++rep)  // twice: the second call relies on the counters left clean by the first
      YPB_OK_OR_FAIL(ypb_nms_from_head(&hd.desc, hd.angle, 1, dt, &nb.p, &nb.o, nb.ws, nb.ws_bytes, nullptr));
    Result r1 = fetch(nb);
    ok = same_result(r2, r1, f.what) && ok;
    if (&f == &forms[0]) first = r1;
    else ok = same_result(first, r1, f.what, true) && ok;  // the candidate counts too: every form of the scan sees the same rows
    if (f.own_counters) {
      std::vector<int32_t> cn = to_host(nb.counters, batch + 1);
      for (int32_t v : cn)
        if (v != 0) { std::printf("  %s: counters not left clean\n", f.what); ok = false; break; }
    }
    if (f.host_counts) {
      for (int b = 0; b < batch; ++b)
        if (nb.count_host[b] != r1.count[b]) { std::printf("  %s: host count of image %d differs\n", f.what, b); ok = false; break; }
      CUDA_OK(cudaFreeHost(nb.count_host));
    }
  }
  long long cand = std::accumulate(r2.cand.begin(), r2.cand.end(), 0LL);
  std::printf("  %-28s B=%d A=%d: %lld candidates -> %lld kept\n", name, batch, A, cand, total_kept(r2));
  if (total_kept(r2) == 0) { std::printf("  nothing kept: the case does not exercise the path\n"); ok = false; }
  return ok;
}

// ---- case: TorchNMS.nms / fast_nms on one box set against a scalar restatement ----------------------------------------
std::vector<float> clustered_boxes(int n, Rng& rng) {
  std::vector<float> b(static_cast<size_t>(n) * 4);
  const int clusters = std::max(1, n / 12);
  std::vector<float> cx(clusters), cy(clusters), sz(clusters);
  for (int i = 0; i < clusters; ++i) { cx[i] = 40 + 560 * rng.uni(); cy[i] = 40 + 560 * rng.uni(); sz[i] = 20 + 120 * rng.uni(); }
  for (int i = 0; i < n; ++i) {
    const int k = static_cast<int>(rng.uni() * clusters) % clusters;
    const float x = cx[k] + 0.15f * sz[k] * rng.gauss(), y = cy[k] + 0.15f * sz[k] * rng.gauss();
    const float w = sz[k] * (0.7f + 0.6f * rng.uni()), h = sz[k] * (0.7f + 0.6f * rng.uni());
    b[i * 4 + 0] = x - w / 2; b[i * 4 + 1] = y - h / 2; b[i * 4 + 2] = x + w / 2; b[i * 4 + 3] = y + h / 2;
  }
  return b;
}
float iou_plain(const float* a, const float* b, float eps) {
  const float iw = std::max(0.0f, std::min(a[2], b[2]) - std::max(a[0], b[0]));
  const float ih = std::max(0.0f, std::min(a[3], b[3]) - std::max(a[1], b[1]));
  const float inter = iw * ih;
  const float a1 = (a[2] - a[0]) * (a[3] - a[1]), a2 = (b[2] - b[0]) * (b[3] - b[1]);
  return eps > 0 ? inter / (a1 + a2 - inter + eps) : inter / (a1 + a2 - inter);
}
bool case_nms_boxes(int n, uint64_t seed) {
  Rng rng(seed);
  std::vector<float> boxes = clustered_boxes(n, rng), scores(n);
  for (int i = 0; i < n; ++i) scores[i] = 0.05f + 0.9f * rng.uni() + 1e-6f * i;  // pairwise distinct with overwhelming probability
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return scores[a] > scores[b]; });
  const float thr = 0.5f;
  std::vector<long long> want_greedy, want_fast;
  {
    std::vector<char> dead(n, 0);
    for (int i = 0; i < n; ++i) {
      if (dead[i]) continue;
      want_greedy.push_back(order[i]);
      for (int j = i + 1; j < n; ++j)
        if (!dead[j] && iou_plain(&boxes[order[i] * 4], &boxes[order[j] * 4], 0.f) > thr) dead[j] = 1;
    }
    for (int j = 0; j < n; ++j) {
      bool keep = true;
      for (int i = 0; i < j && keep; ++i)
        if (iou_plain(&boxes[order[i] * 4], &boxes[order[j] * 4], 1e-7f) >= thr) keep = false;
      if (keep) want_fast.push_back(order[j]);
    }
  }
  float* dboxes = to_device(boxes);
  float* dscores = to_device(scores);
  long long* keep = dalloc<long long>(n);
  int32_t* kc = dalloc<int32_t>(1);
  const size_t wsb = ypb_nms_boxes_workspace_bytes(n);
  void* ws = dalloc<uint8_t>(wsb);
  bool ok = true;
  const int rules[2] = {YPB_NMS_GREEDY, YPB_NMS_FAST_BOXIOU};
  for (int r = 0; r < 2; ++r) {
    YPB_OK_OR_FAIL(ypb_nms_boxes(dboxes, dscores, n, 4, rules[r], thr, reinterpret_cast<int64_t*>(keep), kc, ws, wsb, nullptr));
    CUDA_OK(cudaDeviceSynchronize());
    const int got_n = to_host(kc, 1)[0];
    const std::vector<long long>& want = r == 0 ? want_greedy : want_fast;
    std::vector<long long> got = to_host(keep, std::max(0, std::min(got_n, n)));
    if (got != want) {
      std::printf("  nms_boxes n=%d rule %d: kept %d, scalar restatement kept %zu (or order differs)\n", n, rules[r], got_n, want.size());
      ok = false;
    }
  }
  std::printf("  nms_boxes n=%d: greedy keeps %zu, Fast-NMS keeps %zu\n", n, want_greedy.size(), want_fast.size());
  return ok;
}

bool case_pairwise(uint64_t seed) {
  Rng rng(seed);
  const int n = 77, m = 130;
  std::vector<float> a = clustered_boxes(n, rng), b = clustered_boxes(m, rng);
  float* da = to_device(a);
  float* db = to_device(b);
  float* out = dalloc<float>(static_cast<size_t>(n) * m);
  YPB_OK_OR_FAIL(ypb_pairwise_iou(da, n, db, m, 4, out, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  std::vector<float> got = to_host(out, static_cast<size_t>(n) * m);
  double worst = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) worst = std::max(worst, std::fabs(double(got[i * m + j]) - iou_plain(&a[i * 4], &b[j * 4], 1e-7f)));
  // rotated boxes: symmetric, 1 on the diagonal is not required (probiou of a box with itself is 1 - sqrt(eps-ish)); finite, in [0, 1]
  std::vector<float> r(static_cast<size_t>(n) * 5);
  for (int i = 0; i < n; ++i) {
    r[i * 5 + 0] = 100 + 400 * rng.uni(); r[i * 5 + 1] = 100 + 400 * rng.uni();
    r[i * 5 + 2] = 10 + 100 * rng.uni(); r[i * 5 + 3] = 10 + 100 * rng.uni(); r[i * 5 + 4] = -0.78f + 2.35f * rng.uni();
  }
  float* dr = to_device(r);
  float* out2 = dalloc<float>(static_cast<size_t>(n) * n);
  YPB_OK_OR_FAIL(ypb_pairwise_iou(dr, n, dr, n, 5, out2, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  std::vector<float> p = to_host(out2, static_cast<size_t>(n) * n);
  bool ok = worst <= 1e-6;
  for (int i = 0; i < n && ok; ++i)
    for (int j = 0; j < n; ++j) {
      const float v = p[i * n + j];
      if (!(v >= -1e-6f && v <= 1.0f + 1e-6f) || std::fabs(v - p[j * n + i]) > 1e-5f) { ok = false; std::printf("  probiou[%d,%d]=%g / %g\n", i, j, v, p[j * n + i]); break; }
    }
  std::printf("  pairwise_iou: box_iou max |diff| %.2e, probiou symmetric and in [0,1]: %s\n", worst, ok ? "yes" : "NO");
  return ok;
}

// ---- case: the free-standing decode pieces and the dense keypoint decode against scalar formulas ------------------------
bool case_decode_pieces(uint64_t seed) {
  Rng rng(seed);
  const int B = 2;
  Geom g{3, {20, 10, 5}, {20, 10, 5}, {8.f, 16.f, 32.f}, 1};
  const int A = g.anchors();
  std::vector<float> x(static_cast<size_t>(B) * 64 * A);
  for (float& v : x) v = 2.0f * rng.gauss();
  float* dx = to_device(x);
  float* dist = dalloc<float>(static_cast<size_t>(B) * 4 * A);
  YPB_OK_OR_FAIL(ypb_dfl_expectation(dx, YPB_F32, B, 16, A, 64LL * A, A, dist, 4LL * A, A, nullptr));
  std::vector<float> ap(2 * static_cast<size_t>(A));
  {
    int a = 0;
    for (int l = 0; l < g.nl; ++l)
      for (int yy = 0; yy < g.h[l]; ++yy)
        for (int xx = 0; xx < g.w[l]; ++xx, ++a) { ap[a] = xx + 0.5f; ap[A + a] = yy + 0.5f; }
  }
  float* dap = to_device(ap);
  float* box = dalloc<float>(static_cast<size_t>(B) * 4 * A);
  YPB_OK_OR_FAIL(ypb_dist2bbox(dist, 4LL * A, A, dap, 0, A, 1, nullptr, 0, YPB_F32, B, A, 1, box, 4LL * A, A, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  std::vector<float> gd = to_host(dist, static_cast<size_t>(B) * 4 * A), gb = to_host(box, static_cast<size_t>(B) * 4 * A);
  double wd = 0, wb = 0;
  for (int b = 0; b < B; ++b)
    for (int a = 0; a < A; ++a) {
      float d[4];
      for (int s = 0; s < 4; ++s) {
        const float* p = &x[(static_cast<size_t>(b) * 64 + s * 16) * A + a];
        float mx = p[0];
        for (int k = 1; k < 16; ++k) mx = std::max(mx, p[static_cast<size_t>(k) * A]);
        double num = 0, den = 0;
        for (int k = 0; k < 16; ++k) { const double e = std::exp(double(p[static_cast<size_t>(k) * A]) - mx); num += k * e; den += e; }
        d[s] = static_cast<float>(num / den);
        wd = std::max(wd, std::fabs(double(gd[(static_cast<size_t>(b) * 4 + s) * A + a]) - d[s]));
      }
      // the box from the DEVICE distances (isolates dist2bbox from the softmax tolerance)
      const float l = gd[(static_cast<size_t>(b) * 4 + 0) * A + a], t = gd[(static_cast<size_t>(b) * 4 + 1) * A + a];
      const float r = gd[(static_cast<size_t>(b) * 4 + 2) * A + a], bt = gd[(static_cast<size_t>(b) * 4 + 3) * A + a];
      const float x1 = ap[a] - l, y1 = ap[A + a] - t, x2 = ap[a] + r, y2 = ap[A + a] + bt;
      const float want[4] = {(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1};
      for (int s = 0; s < 4; ++s) wb = std::max(wb, std::fabs(double(gb[(static_cast<size_t>(b) * 4 + s) * A + a]) - want[s]));
    }
  // keypoints: (B, nk*3, A) -> x,y: (v*2 + (anchor - 0.5)) * stride; visibility: sigmoid (head.py:1254-1273)
  const int nk = 5, nd = 3, kc = nk * nd;
  std::vector<float> kp(static_cast<size_t>(B) * kc * A);
  for (float& v : kp) v = rng.gauss();
  float* dk = to_device(kp);
  float* ko = dalloc<float>(kp.size());
  ypb_head_desc hd;
  std::memset(&hd, 0, sizeof(hd));
  hd.num_levels = g.nl; hd.batch = B; hd.nc = 1; hd.reg_max = 16; hd.dtype = YPB_F32;
  for (int l = 0; l < g.nl; ++l) { hd.level_h[l] = g.h[l]; hd.level_w[l] = g.w[l]; hd.level_stride[l] = g.stride[l]; }
  YPB_OK_OR_FAIL(ypb_kpts_decode(&hd, dk, static_cast<int64_t>(kc) * A, A, kc, nd, ko, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  std::vector<float> gk = to_host(ko, kp.size());
  double wk = 0;
  for (int b = 0; b < B; ++b) {
    int a = 0;
    for (int l = 0; l < g.nl; ++l)
      for (int yy = 0; yy < g.h[l]; ++yy)
        for (int xx = 0; xx < g.w[l]; ++xx, ++a)
          for (int k = 0; k < nk; ++k) {
            const size_t o = (static_cast<size_t>(b) * kc + k * nd) * A + a;
            const double wx = (kp[o] * 2.0 + (xx + 0.5 - 0.5)) * g.stride[l], wy = (kp[o + A] * 2.0 + (yy + 0.5 - 0.5)) * g.stride[l];
            const double wv = 1.0 / (1.0 + std::exp(-double(kp[o + 2 * static_cast<size_t>(A)])));
            wk = std::max(wk, std::fabs(gk[o] - wx) / (1.0 + std::fabs(wx)));
            wk = std::max(wk, std::fabs(gk[o + A] - wy) / (1.0 + std::fabs(wy)));
            wk = std::max(wk, std::fabs(gk[o + 2 * static_cast<size_t>(A)] - wv));
          }
  }
  const bool ok = wd <= 1e-4 && wb <= 1e-5 && wk <= 1e-5;
  std::printf("  dfl max |diff| %.2e, dist2bbox %.2e, kpts_decode %.2e: %s\n", wd, wb, wk, ok ? "ok" : "TOO LARGE");
  return ok;
}

// ---- case: rows after NMS - compaction, rescale, masks, matching ----------------------------------------------------------
bool case_result_rows(uint64_t seed) {
  Rng rng(seed);
  const int B = 3, md = 40, C = 32, cols = 6 + C;
  const int counts[B] = {17, 0, 40};
  std::vector<float> rows(static_cast<size_t>(B) * md * cols, 0.f);
  std::vector<long long> idx(static_cast<size_t>(B) * md, -1);
  for (int b = 0; b < B; ++b) {
    std::vector<float> bx = clustered_boxes(md, rng);
    for (int r = 0; r < md; ++r) {
      float* row = &rows[(static_cast<size_t>(b) * md + r) * cols];
      for (int k = 0; k < 4; ++k) row[k] = std::min(639.f, std::max(0.f, bx[r * 4 + k]));
      row[4] = 0.9f - 0.01f * r; row[5] = static_cast<float>(r % 3);
      for (int k = 0; k < C; ++k) row[6 + k] = rng.gauss();
      idx[static_cast<size_t>(b) * md + r] = 1000 * b + r;
    }
  }
  std::vector<int32_t> cnt(counts, counts + B);
  float* drows = to_device(rows);
  long long* didx = to_device(idx);
  int32_t* dcnt = to_device(cnt);
  const int total = counts[0] + counts[1] + counts[2];
  bool ok = true;
  // compaction
  float* prow = dalloc<float>(static_cast<size_t>(B) * md * cols);
  long long* pidx = dalloc<long long>(static_cast<size_t>(B) * md);
  int32_t* poff = dalloc<int32_t>(B + 1);
  YPB_OK_OR_FAIL(ypb_compact_results(drows, reinterpret_cast<const int64_t*>(didx), dcnt, B, md, cols, prow,
                                     reinterpret_cast<int64_t*>(pidx), poff, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  {
    std::vector<int32_t> off = to_host(poff, B + 1);
    std::vector<float> pr = to_host(prow, static_cast<size_t>(total) * cols);
    std::vector<long long> pi = to_host(pidx, total);
    int run = 0;
    for (int b = 0; b < B; ++b) {
      if (off[b] != run) { std::printf("  compact_results: offsets[%d]=%d, expected %d\n", b, off[b], run); ok = false; }
      for (int r = 0; r < counts[b]; ++r, ++run) {
        if (std::memcmp(&pr[static_cast<size_t>(run) * cols], &rows[(static_cast<size_t>(b) * md + r) * cols], cols * sizeof(float)) != 0 ||
            pi[run] != idx[static_cast<size_t>(b) * md + r]) { std::printf("  compact_results: row %d differs\n", run); ok = false; b = B; break; }
      }
    }
    if (ok && off[B] != total) { std::printf("  compact_results: offsets[B]=%d, expected %d\n", off[B], total); ok = false; }
  }
  // masks of the batch: the one-kernel form and the work-list form must agree
  const int mh = 40, mw = 40, oh = 160, ow = 160;
  std::vector<float> protos(static_cast<size_t>(B) * C * mh * mw);
  for (float& v : protos) v = rng.gauss();
  float* dprotos = to_device(protos);
  ypb_protos_desc pd;
  std::memset(&pd, 0, sizeof(pd));
  pd.ptr = dprotos; pd.dtype = YPB_F32; pd.channels = C; pd.mh = mh; pd.mw = mw;
  pd.stride_b = static_cast<int64_t>(C) * mh * mw; pd.stride_c = static_cast<int64_t>(mh) * mw;
  std::vector<int32_t> off = {0, counts[0], counts[0] + counts[1], total};
  int32_t* doff = to_device(off);
  // boxes of the rows are in 640-pixel units: bring them into the 160-pixel output frame for this small case
  std::vector<float> rows160 = rows;
  for (size_t r = 0; r < rows160.size() / cols; ++r)
    for (int k = 0; k < 4; ++k) rows160[r * cols + k] *= 0.25f;
  float* drows160 = to_device(rows160);
  uint8_t* m1 = dalloc<uint8_t>(static_cast<size_t>(total) * oh * ow);
  uint8_t* m2 = dalloc<uint8_t>(static_cast<size_t>(total) * oh * ow);
  CUDA_OK(cudaMemset(m1, 0x55, static_cast<size_t>(total) * oh * ow));
  CUDA_OK(cudaMemset(m2, 0xAA, static_cast<size_t>(total) * oh * ow));
  const float rw = static_cast<float>(mw) / ow, rh = static_cast<float>(mh) / oh;
  YPB_OK_OR_FAIL(ypb_process_mask(&pd, drows160 + 6, static_cast<int64_t>(md) * cols, cols, drows160, static_cast<int64_t>(md) * cols, cols, doff,
                                  B, total, oh, ow, 0, 0, mh, mw, YPB_MASK_CROP_PROTO, rw, rh, m1, nullptr, 0, nullptr));
  const size_t mwsb = ypb_process_mask_workspace_bytes(total, oh, ow);
  void* mws = dalloc<uint8_t>(mwsb);
  YPB_OK_OR_FAIL(ypb_process_mask(&pd, drows160 + 6, static_cast<int64_t>(md) * cols, cols, drows160, static_cast<int64_t>(md) * cols, cols, doff,
                                  B, total, oh, ow, 0, 0, mh, mw, YPB_MASK_CROP_PROTO, rw, rh, m2, mws, mwsb, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  {
    std::vector<uint8_t> a = to_host(m1, static_cast<size_t>(total) * oh * ow), b = to_host(m2, static_cast<size_t>(total) * oh * ow);
    size_t on = 0, diff = 0, bad = 0;
    for (size_t i = 0; i < a.size(); ++i) { on += a[i] == 1; diff += a[i] != b[i]; bad += a[i] > 1 || b[i] > 1; }
    std::printf("  process_mask: %d masks of %dx%d, %zu pixels set, %zu differ between the two forms\n", total, oh, ow, on, diff);
    if (bad || on == 0 || diff * 1000 > a.size()) { std::printf("  process_mask: unexpected output\n"); ok = false; }
  }
  // rescale to an original image of 480x640 letterboxed into 640x640 (scale_boxes, ops.py:102-135)
  ypb_scale_xform xf = {1.0f, 0.f, 80.f, 640.f, 480.f, 0.f, 80.f, 0.f};
  float* scaled = to_device(rows);
  YPB_OK_OR_FAIL(ypb_scale_rows(scaled, static_cast<int64_t>(md) * cols, cols, B, md, dcnt, nullptr, &xf, YPB_BOXES_XYXY, YPB_SCALE_PADDING,
                                0, nullptr, 0, 0, 0, 0, nullptr));
  CUDA_OK(cudaDeviceSynchronize());
  {
    std::vector<float> s = to_host(scaled, rows.size());
    double worst = 0;
    for (int b = 0; b < B; ++b)
      for (int r = 0; r < md; ++r) {
        const float* in = &rows[(static_cast<size_t>(b) * md + r) * cols];
        const float* out = &s[(static_cast<size_t>(b) * md + r) * cols];
        for (int k = 0; k < 4; ++k) {
          if (r >= counts[b]) continue;  // rows past the kept count are not part of the result
          float want = (in[k] - ((k & 1) ? xf.pad_y : xf.pad_x)) / xf.gain;
          want = std::min(std::max(want, 0.f), (k & 1) ? xf.img_h : xf.img_w);
          worst = std::max(worst, std::fabs(double(out[k]) - want));
        }
        for (int k = 4; k < cols && r < counts[b]; ++k)
          if (out[k] != in[k]) worst = 1e9;
      }
    std::printf("  scale_rows: max |diff| %.2e\n", worst);
    if (worst > 1e-4) ok = false;
  }
  // validator matching, boxes mode: the labels are the first kept rows themselves -> every one of them is a true positive at 0.5
  {
    std::vector<float> labels;
    std::vector<int32_t> loff = {0};
    for (int b = 0; b < B; ++b) {
      const int m = std::min(counts[b], 6);
      for (int r = 0; r < m; ++r) {
        const float* row = &rows[(static_cast<size_t>(b) * md + r) * cols];
        labels.push_back(row[5]);
        for (int k = 0; k < 4; ++k) labels.push_back(row[k]);
      }
      loff.push_back(loff.back() + m);
    }
    float* dl = to_device(labels);
    int32_t* dlo = to_device(loff);
    const float thr[10] = {0.5f, 0.55f, 0.6f, 0.65f, 0.7f, 0.75f, 0.8f, 0.85f, 0.9f, 0.95f};
    uint8_t* corr = dalloc<uint8_t>(static_cast<size_t>(B) * md * 10);
    CUDA_OK(cudaMemset(corr, 0, static_cast<size_t>(B) * md * 10));
    YPB_OK_OR_FAIL(ypb_match_predictions(drows, static_cast<int64_t>(md) * cols, cols, 5, B, md, dcnt, dl, dlo, 0, 6, nullptr, 0, nullptr, thr, 10,
                                         corr, nullptr, 0, nullptr));
    CUDA_OK(cudaDeviceSynchronize());
    std::vector<uint8_t> c = to_host(corr, static_cast<size_t>(B) * md * 10);
    int tp = 0;
    for (uint8_t v : c) { if (v > 1) ok = false; tp += v; }
    // a prediction identical to its label matches at every threshold unless a better-scoring twin took the label
    int at50 = 0;
    for (int b = 0; b < B; ++b)
      for (int r = 0; r < md; ++r) at50 += c[(static_cast<size_t>(b) * md + r) * 10];
    std::printf("  match_predictions: %d true positives over 10 thresholds, %d at 0.5 (labels %d)\n", tp, at50, loff.back());
    if (at50 != loff.back()) { std::printf("  match_predictions: every label has an identical prediction, all must match at 0.5\n"); ok = false; }
  }
  return ok;
}

// ---- case: the one-sided result ring looped back onto this GPU (world = 1) ------------------------------------------------
bool case_peer_loopback(uint64_t seed) {
  Geom g{3, {20, 10, 5}, {20, 10, 5}, {8.f, 16.f, 32.f}, 80};
  const int B = 2, depth = 3;
  NmsCfg c;
  Head hd[2] = {make_head(g, B, YPB_F32, seed, -7.f, 0.05f, false), make_head(g, B, YPB_F32, seed + 1, -7.f, 0.05f, false)};
  const int A = hd[0].anchors;
  NmsBuffers nb = make_buffers(B, A, g.nc, 0, YPB_F32, c, true, false);
  const long long nrow = static_cast<long long>(B) * nb.max_det * nb.cols;
  const long long slot = (nrow + B + 3) / 4 * 4;
  const long long ring = depth * slot;
  float* buf = dalloc<float>(ring + 32);
  CUDA_OK(cudaMemset(buf, 0, (ring + 32) * sizeof(float)));
  int32_t* state = dalloc<int32_t>(4);
  CUDA_OK(cudaMemset(state, 0, 4 * sizeof(int32_t)));
  long long* slot_index = dalloc<long long>(1);
  float* copy = dalloc<float>(slot);
  int32_t* flags = reinterpret_cast<int32_t*>(buf + ring);
  int32_t* acks = reinterpret_cast<int32_t*>(buf + ring + 16);
  nb.o.num_peers = 1; nb.o.my_rank = 0; nb.o.peer_depth = depth;
  nb.o.peer_rows[0] = buf; nb.o.peer_count[0] = reinterpret_cast<int32_t*>(buf + nrow); nb.o.peer_flag[0] = flags;
  nb.o.peer_state = state; nb.o.peer_ack = acks; nb.o.peer_entry_stride = slot;
  int32_t* ack_ptrs[1] = {acks};
  bool ok = true;
  for (int step = 1; step <= 5; ++step) {  // 5 launches > 3 ring entries: entries are reused under the acknowledgement protocol
    YPB_OK_OR_FAIL(ypb_nms_from_head(&hd[step & 1].desc, nullptr, 0, YPB_F32, &nb.p, &nb.o, nb.ws, nb.ws_bytes, nullptr));
    if (step & 1) YPB_OK_OR_FAIL(ypb_peer_wait(flags, 1, state, 0, depth, ack_ptrs, 0, reinterpret_cast<int64_t*>(slot_index), nullptr));
    else YPB_OK_OR_FAIL(ypb_peer_wait_copy(flags, 1, state, 0, depth, ack_ptrs, 0, reinterpret_cast<int64_t*>(slot_index), buf, slot, copy, nullptr));
    Result r = fetch(nb);
    const long long si = to_host(slot_index, 1)[0];
    if (si != step % depth) { std::printf("  peer ring: step %d handed out entry %lld, expected %d\n", step, si, step % depth); ok = false; break; }
    std::vector<float> entry = (step & 1) ? to_host(buf + si * slot, slot) : to_host(copy, slot);
    for (int b = 0; b < B && ok; ++b) {
      int32_t n;
      std::memcpy(&n, &entry[nrow + b], 4);
      if (n != r.count[b]) { std::printf("  peer ring: step %d image %d count %d, expected %d\n", step, b, n, r.count[b]); ok = false; break; }
      const size_t o = static_cast<size_t>(b) * nb.max_det * nb.cols;
      if (std::memcmp(&entry[o], &r.rows[o], static_cast<size_t>(std::min(n, nb.max_det)) * nb.cols * sizeof(float)) != 0) {
        std::printf("  peer ring: step %d rows of image %d differ from the local result\n", step, b);
        ok = false;
      }
    }
    if (total_kept(r) == 0) { std::printf("  peer ring: nothing kept\n"); ok = false; }
  }
  std::vector<int32_t> st = to_host(state, 4);
  if (st[3] != 0) { std::printf("  peer ring: overrun marker %d\n", st[3]); ok = false; }
  std::printf("  peer ring loopback: 5 launches through %d entries, launches sent %d, handed out %d\n", depth, st[1], st[2]);
  return ok;
}

void run(const char* name, bool ok) {
  ++g_cases;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { std::printf("  CUDA error after the case: %s\n", cudaGetErrorString(e)); ok = false; }
  std::printf("[%s] %s\n", ok ? " ok " : "FAIL", name);
  if (!ok) ++g_failed;
  free_all();
}

}  // namespace

int main(int argc, char** argv) {
  bool list = false;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--small")) g_small = true;
    else if (!std::strcmp(argv[i], "--list")) list = true;
    else { std::printf("usage: %s [--small] [--list]\n", argv[0]); return 2; }
  }
  if (ypb_abi_version() != YPB_ABI_VERSION) { std::printf("ABI version %d != header %d\n", ypb_abi_version(), YPB_ABI_VERSION); return 2; }
  if (list) {
    std::printf("cases: fused-vs-two-call (fp32 / bf16 / fp16 / multi-label / OBB / odd grid / C1 geometry), nms_boxes, pairwise_iou, "
                "decode pieces, result rows, peer ring loopback; ABI version %d\n", ypb_abi_version());
    return 0;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { std::printf("no CUDA device: this path has no CPU fallback\n"); return 4; }
  CUDA_OK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, 0));
  std::printf("device 0: %s (sm_%d%d, %d SMs), %s sizes\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, g_small ? "small" : "full");

  const Geom small{3, {32, 16, 8}, {32, 16, 8}, {8.f, 16.f, 32.f}, 80};  // 256 px: every level a multiple of 8 anchors (vector paths)
  const Geom odd{3, {13, 7, 3}, {13, 7, 3}, {8.f, 16.f, 32.f}, 80};
  const Geom c1{3, {80, 40, 20}, {80, 40, 20}, {8.f, 16.f, 32.f}, 80};
  const Geom obb{3, {32, 16, 8}, {32, 16, 8}, {8.f, 16.f, 32.f}, 15};
  NmsCfg predict;
  NmsCfg val;
  val.conf = 0.001f; val.multi_label = true;
  NmsCfg rot;
  rot.rotated = true; rot.conf = 0.05f;
  NmsCfg agnostic;
  agnostic.max_wh = 0.f;

  run("fused == two-call, fp32 predict, 256 px", case_fused("fp32 predict", small, 3, YPB_F32, predict, -7.f, 0.06f, 11));
  run("fused == two-call, bf16 predict, 256 px", case_fused("bf16 predict", small, 3, YPB_BF16, predict, -7.f, 0.06f, 12));
  run("fused == two-call, fp16 predict, 256 px", case_fused("fp16 predict", small, 2, YPB_F16, predict, -7.f, 0.06f, 13));
  run("fused == two-call, fp32 agnostic, odd grids (scalar path)", case_fused("fp32 odd grid", odd, 3, YPB_F32, agnostic, -7.f, 0.08f, 14));
  run("fused == two-call, fp32 multi-label val mode (radix prefix, max_nms)", case_fused("fp32 val", small, 2, YPB_F32, val, -7.f, 0.06f, 15));
  run("fused == two-call, fp32 OBB (rotated ProbIoU Fast-NMS)", case_fused("fp32 obb", obb, 2, YPB_F32, rot, -6.f, 0.06f, 16));
  if (!g_small) {
    run("fused == two-call, fp32 predict, 640 px (C1 geometry), B=4", case_fused("fp32 C1", c1, 4, YPB_F32, predict, -7.f, 0.03f, 17));
    // 20 images: the scan grid exceeds two CTAs per SM, so the survivor decode runs as its own kernel and the persistent TMA scan is taken
    run("fused == two-call, fp32 predict, 640 px, B=20 (split decode, TMA scan)", case_fused("fp32 C1 B=20", c1, 20, YPB_F32, predict, -7.f, 0.03f, 20));
    run("fused == two-call, bf16 predict, 640 px, B=20 (split decode, TMA scan)", case_fused("bf16 C1 B=20", c1, 20, YPB_BF16, predict, -7.f, 0.03f, 27));
    run("fused == two-call, bf16 val mode, 640 px, B=2", case_fused("bf16 C1 val", c1, 2, YPB_BF16, val, -8.5f, 0.03f, 18));
    run("fused == two-call, bf16 OBB", case_fused("bf16 obb", obb, 3, YPB_BF16, rot, -6.f, 0.06f, 19));
  }
  run("TorchNMS.nms / fast_nms, 600 boxes", case_nms_boxes(600, 21));
  if (!g_small) run("TorchNMS.nms / fast_nms, 5000 boxes (radix-select path)", case_nms_boxes(5000, 22));
  run("box_iou / batch_probiou matrices", case_pairwise(23));
  run("DFL.forward, decode_bboxes, Pose.kpts_decode", case_decode_pieces(24));
  run("compact_results, process_mask, scale_rows, match_predictions", case_result_rows(25));
  run("one-sided result ring, loopback", case_peer_loopback(26));

  std::printf("%d of %d cases passed\n", g_cases - g_failed, g_cases);
  return g_failed ? 1 : 0;
}
