// extern "C" surface of libyolopost_b200 (declared in include/yolopost_b200.h): argument validation, geometry,
// workspace carve-up and kernel sequencing.  Nothing here allocates, frees or synchronises.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ypb_common.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  return fail(YPB_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
}

// Where the survivors' boxes are decoded: inside the class-scan kernel (one kernel boundary less: the latency form) or by the
// separate GPU-wide decode_tiles_kernel (CTAs holding survivors do not become the scan kernel's tail: the throughput form,
// 39 + 17 us overlappable vs 53 us at C2 B=64).  Auto: fused when the scan grid is at most two CTAs per SM - measured at
// C2 B=8 36.8 vs 47.9 us per call, B=1 equal GPU time and 3-5 us less host time.  YPB_FUSE_DECODE=1 / 0 forces either form.
bool split_decode_requested(const ypb::HeadGeom& g) {
  static const int forced = [] { const char* e = std::getenv("YPB_FUSE_DECODE"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
  if (forced >= 0) return forced == 0;
  const long long scan_ctas = static_cast<long long>((g.group_start[g.num_levels] + 127) / 128) * g.batch;
  return scan_ctas > 2 * 148;
}

// ypb_nms_params.scan_kernel == YPB_SCAN_AUTO: the one-wave LDG class scan unless YPB_SCAN_TMA=1 asks for the persistent
// TMA-fed kernel (ypb_scan_tma.cu) process-wide.
bool tma_scan_requested(int scan_kernel) {
  static const bool env = [] { const char* e = std::getenv("YPB_SCAN_TMA"); return e && e[0] == '1'; }();
  return scan_kernel == YPB_SCAN_TMA || (scan_kernel == YPB_SCAN_AUTO && env);
}

size_t dtype_size(int dt) { return dt == YPB_F32 ? 4 : 2; }
bool dtype_ok(int dt) { return dt == YPB_F32 || dt == YPB_F16 || dt == YPB_BF16; }

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Validates the head and fills the device-side geometry.  `vec` = anchors per thread (widest that every level,
// stride and pointer allows: 16-byte accesses, else scalar).
int build_geom(const ypb_head_desc* h, ypb::HeadGeom* g, int* vec_out, const void* extra_ptr_a, const void* extra_ptr_b,
               long long extra_stride_a, long long extra_stride_b) {
  if (!h) return fail(YPB_ERR_INVALID_ARGUMENT, "head descriptor is NULL");
  if (h->num_levels < 1 || h->num_levels > YPB_MAX_LEVELS)
    return fail(YPB_ERR_INVALID_ARGUMENT, "num_levels=%d outside [1,%d]", h->num_levels, YPB_MAX_LEVELS);
  if (h->batch < 0 || h->nc < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d nc=%d invalid", h->batch, h->nc);
  if (!dtype_ok(h->dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown dtype %d", h->dtype);
  if (h->reg_max != 16)
    return fail(YPB_ERR_UNSUPPORTED, "reg_max=%d: only 16 is built (every head in the reference bundle, head.py:89)", h->reg_max);
  const int wide = 16 / static_cast<int>(dtype_size(h->dtype));
  int vec = wide;
  long long anchors = 0;
  for (int l = 0; l < h->num_levels; ++l) {
    if (!h->level_ptr[l] && h->batch > 0) return fail(YPB_ERR_INVALID_ARGUMENT, "level %d pointer is NULL", l);
    if (h->level_h[l] < 1 || h->level_w[l] < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "level %d has empty grid", l);
    const long long hw = static_cast<long long>(h->level_h[l]) * h->level_w[l];
    if (h->level_channel_stride[l] < hw) return fail(YPB_ERR_INVALID_ARGUMENT, "level %d channel stride < H*W", l);
    anchors += hw;
    if (hw % wide || h->level_channel_stride[l] % wide || h->level_batch_stride[l] % wide || !aligned(h->level_ptr[l], 16))
      vec = 1;
  }
  if (ypb::bits_for(anchors) + ypb::bits_for(h->nc) > 31)
    return fail(YPB_ERR_UNSUPPORTED, "anchor and class index do not fit the 31-bit row id");
  if (extra_ptr_a && !aligned(extra_ptr_a, 16)) vec = 1;
  if (extra_ptr_b && !aligned(extra_ptr_b, 16)) vec = 1;
  if (extra_stride_a % wide || extra_stride_b % wide) vec = 1;
  std::memset(g, 0, sizeof(*g));
  g->num_levels = h->num_levels;
  g->batch = h->batch;
  g->nc = h->nc;
  g->reg_max = h->reg_max;
  int as = 0, gs = 0;
  for (int l = 0; l < h->num_levels; ++l) {
    g->ptr[l] = h->level_ptr[l];
    g->h[l] = h->level_h[l];
    g->w[l] = h->level_w[l];
    g->bstride[l] = h->level_batch_stride[l];
    g->cstride[l] = h->level_channel_stride[l];
    g->stride[l] = h->level_stride[l];
    g->anchor_start[l] = as;
    g->group_start[l] = gs;
    as += h->level_h[l] * h->level_w[l];
    gs += h->level_h[l] * h->level_w[l] / vec;
  }
  for (int l = h->num_levels; l <= YPB_MAX_LEVELS; ++l) {
    g->anchor_start[l] = as;
    g->group_start[l] = gs;
  }
  g->anchors = as;
  *vec_out = vec;
  return YPB_OK;
}

int check_params(const ypb_nms_params* p, const ypb_nms_out* out) {
  if (!p || !out) return fail(YPB_ERR_INVALID_ARGUMENT, "params/out is NULL");
  if (p->nc < 1 || p->extra < 0 || p->max_det < 1 || p->max_nms < 1 || p->rows_cap < 1)
    return fail(YPB_ERR_INVALID_ARGUMENT, "nc=%d extra=%d max_det=%d max_nms=%d rows_cap=%d invalid", p->nc, p->extra,
                p->max_det, p->max_nms, p->rows_cap);
  if (p->rule < YPB_NMS_GREEDY || p->rule > YPB_NMS_FAST_BOXIOU) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown rule %d", p->rule);
  if ((!(p->conf_thres >= 0.f) && !p->conf_per_image && !p->boxes_xyxy) || !(p->iou_thres_eff >= 0.f) || p->iou_thres_eff > 1.f)
    return fail(YPB_ERR_INVALID_ARGUMENT, "conf/iou threshold outside [0,1] (nms.py:59-60)");
  if (!out->rows || !out->count) return fail(YPB_ERR_INVALID_ARGUMENT, "out.rows / out.count is NULL");
  if (out->num_peers < 0 || out->num_peers > YPB_MAX_PEERS) return fail(YPB_ERR_INVALID_ARGUMENT, "num_peers=%d outside [0,%d]", out->num_peers, YPB_MAX_PEERS);
  if (out->num_peers > 0) {
    if (out->my_rank < 0 || out->my_rank >= out->num_peers || !out->peer_state) return fail(YPB_ERR_INVALID_ARGUMENT, "peer gather: my_rank / peer_state invalid");
    if (out->peer_depth < 1 || (out->peer_depth > 1 && out->peer_entry_stride < 1)) return fail(YPB_ERR_INVALID_ARGUMENT, "peer gather: peer_depth=%d / peer_entry_stride invalid", out->peer_depth);
    for (int i = 0; i < out->num_peers; ++i)
      if (!out->peer_rows[i] || !out->peer_count[i] || !out->peer_flag[i]) return fail(YPB_ERR_INVALID_ARGUMENT, "peer gather: pointer of peer %d is NULL", i);
  }
  return YPB_OK;
}

ypb::SuppressArgs suppress_args(const ypb_nms_params* p, const ypb_nms_out* out, const ypb::Workspace& w, int batch,
                                int anchors) {
  ypb::SuppressArgs s{};
  s.batch = batch; s.anchors = anchors; s.nc = p->nc; s.extra = p->extra; s.max_det = p->max_det; s.max_nms = p->max_nms;
  s.rule = p->rule; s.rows_cap = p->rows_cap; s.multi_label = p->multi_label;
  s.cls_bits = ypb::bits_for(p->nc); s.anchor_bits = ypb::bits_for(anchors);
  s.iou_thr = p->iou_thres_eff; s.max_wh = p->max_wh;
  s.box_div = p->nms_box_divisor; s.box_mult = p->nms_box_multiplier; s.pad_zero = p->pad_output;
  s.row_count = w.row_count; s.keys_a = w.keys_a; s.keys_b = w.keys_b; s.cand_box = w.cand_box;
  s.cand_ang = p->rule == YPB_NMS_FAST_PROBIOU ? w.cand_ang : nullptr;
  s.kept_box = w.kept_box; s.kept_area = w.kept_area; s.kept_key = w.kept_key; s.rec = w.rec;
  s.out_rows = out->rows; s.out_idx = reinterpret_cast<long long*>(out->idx); s.out_count = out->count; s.out_count_host = out->count_host; s.out_cand = out->cand_count;
  s.scale_xforms = out->scale_xforms; s.scale_padding = out->scale_padding;
  s.num_peers = out->num_peers; s.my_rank = out->my_rank; s.peer_state = out->peer_state;
  s.peer_ack = out->peer_ack; s.peer_depth = out->peer_depth > 0 ? out->peer_depth : 1; s.peer_entry_stride = out->peer_entry_stride;
  for (int i = 0; i < YPB_MAX_PEERS; ++i) {
    s.peer_rows[i] = out->peer_rows[i]; s.peer_count[i] = out->peer_count[i]; s.peer_flag[i] = out->peer_flag[i];
  }
  return s;
}

}  // namespace

extern "C" {

int ypb_abi_version(void) { return YPB_ABI_VERSION; }

const char* ypb_last_error_string(void) { return g_err; }

size_t ypb_nms_workspace_bytes(int32_t batch, int32_t anchors, int32_t rows_cap, int32_t max_det, int32_t max_nms,
                               int32_t rule) {
  if (batch < 0 || anchors < 0 || rows_cap < 0 || max_det < 0 || max_nms < 0) return 0;
  return ypb::carve_workspace(nullptr, batch, anchors, rows_cap, max_det, max_nms, rule).bytes + 256;
}

int ypb_decode_dense(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t append_angle,
                     int32_t xyxy, void* out, int32_t out_dtype, int64_t out_stride_b, int64_t out_stride_c,
                     void* stream) {
  ypb::HeadGeom g;
  int vec;
  int rc = build_geom(head, &g, &vec, angle, out, out_stride_b, out_stride_c);
  if (rc) return rc;
  if (!out) return fail(YPB_ERR_INVALID_ARGUMENT, "out is NULL");
  if (out_dtype != head->dtype)
    return fail(YPB_ERR_UNSUPPORTED, "out dtype %d != head dtype %d (Detect._inference keeps the dtype, head.py:169)", out_dtype, head->dtype);
  if (out_stride_c < g.anchors) return fail(YPB_ERR_INVALID_ARGUMENT, "out channel stride < anchors");
  if (head->batch == 0) return YPB_OK;
  // 16-bit heads: 4 anchors (64-bit accesses) per thread instead of 8.  The 8-wide form needs 168-192 registers (3 CTAs
  // per SM, 1.3 waves) and is latency-bound at half the bytes; YPB_DENSE16_VEC=8 restores it for comparison.
  // (a persistent TMA-fed form of this kernel was built and measured in round 2: 65 us vs 56 us for the bf16 C2 batch - the dense
  // decode is bound by instruction issue (31 M warp instructions, 224 MUFU per anchor), not by bytes in flight; removed)
  static const int dense16_vec = [] { const char* e = std::getenv("YPB_DENSE16_VEC"); return (e && e[0] == '8') ? 8 : 4; }();
  if (dtype_size(head->dtype) == 2 && vec == 8 && dense16_vec == 4) {
    vec = 4;
    int gs = 0;
    for (int l = 0; l < head->num_levels; ++l) { g.group_start[l] = gs; gs += head->level_h[l] * head->level_w[l] / vec; }
    for (int l = head->num_levels; l <= YPB_MAX_LEVELS; ++l) g.group_start[l] = gs;
  }
  cudaError_t e = ypb::launch_decode_dense(g, head->dtype, angle, angle_is_logit, append_angle, xyxy, out, out_dtype,
                                           out_stride_b, out_stride_c, vec, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_decode_dense");
  return YPB_OK;
}

int ypb_dfl_expectation(const void* x, int32_t dtype, int32_t batch, int32_t reg_max, int32_t anchors, int64_t stride_b,
                        int64_t stride_c, void* out, int64_t out_stride_b, int64_t out_stride_c, void* stream) {
  if (!dtype_ok(dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
  if (batch < 0 || anchors < 0) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d anchors=%d invalid", batch, anchors);
  if (reg_max != 16) return fail(YPB_ERR_UNSUPPORTED, "reg_max=%d: only 16 is built (block.py:232)", reg_max);
  if (batch == 0 || anchors == 0) return YPB_OK;
  if (!x || !out) return fail(YPB_ERR_INVALID_ARGUMENT, "x / out is NULL");
  if (batch > 65535) return fail(YPB_ERR_UNSUPPORTED, "batch %d > 65535", batch);
  cudaError_t e = ypb::launch_dfl(x, dtype, batch, anchors, stride_b, stride_c, out, out_stride_b, out_stride_c,
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_dfl_expectation");
  return YPB_OK;
}

int ypb_dist2bbox(const void* dist, int64_t dist_stride_b, int64_t dist_stride_c, const void* anchor_points,
                  int64_t anchor_stride_b, int64_t anchor_stride_c, int64_t anchor_stride_a, const void* angle, int64_t angle_stride_b, int32_t dtype,
                  int32_t batch, int32_t anchors, int32_t xywh, void* out, int64_t out_stride_b, int64_t out_stride_c,
                  void* stream) {
  if (!dtype_ok(dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
  if (batch < 0 || anchors < 0) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d anchors=%d invalid", batch, anchors);
  if (batch == 0 || anchors == 0) return YPB_OK;
  if (!dist || !anchor_points || !out) return fail(YPB_ERR_INVALID_ARGUMENT, "dist / anchor_points / out is NULL");
  if (batch > 65535) return fail(YPB_ERR_UNSUPPORTED, "batch %d > 65535", batch);
  ypb::Dist2BoxArgs d{};
  d.dist = dist; d.dsb = dist_stride_b; d.dsc = dist_stride_c;
  d.anchor_points = anchor_points; d.asb = anchor_stride_b; d.asc = anchor_stride_c; d.asa = anchor_stride_a;
  d.angle = angle; d.angle_sb = angle_stride_b;
  d.batch = batch; d.anchors = anchors; d.xywh = xywh;
  d.out = out; d.osb = out_stride_b; d.osc = out_stride_c;
  cudaError_t e = ypb::launch_dist2bbox(d, dtype, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_dist2bbox");
  return YPB_OK;
}

int ypb_nms_from_head(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t value_dtype,
                      const ypb_nms_params* p, const ypb_nms_out* out, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return ypb_nms_from_head_stage(head, angle, angle_is_logit, value_dtype, p, out, workspace, workspace_bytes, stream, 0);
}

static int nms_from_head_impl(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t value_dtype,
                              const ypb_riders_desc* riders, const ypb_nms_params* p, const ypb_nms_out* out,
                              void* workspace, size_t workspace_bytes, void* stream, int32_t stage);

int ypb_nms_from_head_stage(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t value_dtype,
                            const ypb_nms_params* p, const ypb_nms_out* out, void* workspace, size_t workspace_bytes,
                            void* stream, int32_t stage) {
  return nms_from_head_impl(head, angle, angle_is_logit, value_dtype, nullptr, p, out, workspace, workspace_bytes, stream, stage);
}

int ypb_nms_from_head_riders(const ypb_head_desc* head, const ypb_riders_desc* riders, int32_t value_dtype,
                             const ypb_nms_params* p, const ypb_nms_out* out, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (!riders) return fail(YPB_ERR_INVALID_ARGUMENT, "riders descriptor is NULL");
  return nms_from_head_impl(head, nullptr, 0, value_dtype, riders, p, out, workspace, workspace_bytes, stream, 0);
}

static int nms_from_head_impl(const ypb_head_desc* head, const void* angle, int32_t angle_is_logit, int32_t value_dtype,
                              const ypb_riders_desc* riders, const ypb_nms_params* p, const ypb_nms_out* out,
                              void* workspace, size_t workspace_bytes, void* stream, int32_t stage) {
  if (stage < 0 || stage > 15) return fail(YPB_ERR_INVALID_ARGUMENT, "stage mask=%d outside [0,15]", stage);
  const bool keep_counters = (stage & 8) != 0;
  stage &= 7;
  if (stage == 0) stage = 7;
  ypb::HeadGeom g;
  int vec;
  int rc = build_geom(head, &g, &vec, nullptr, nullptr, 0, 0);
  if (rc) return rc;
  rc = check_params(p, out);
  if (rc) return rc;
  if (p->nc != head->nc) return fail(YPB_ERR_INVALID_ARGUMENT, "params.nc=%d != head.nc=%d", p->nc, head->nc);
  if (value_dtype != head->dtype) return fail(YPB_ERR_UNSUPPORTED, "value dtype must equal the head dtype");
  const bool rotated = p->rule == YPB_NMS_FAST_PROBIOU;
  if (rotated && !angle) return fail(YPB_ERR_INVALID_ARGUMENT, "rotated rule needs the angle channel");
  if (!rotated && angle) return fail(YPB_ERR_INVALID_ARGUMENT, "angle given but rule is not FAST_PROBIOU");
  if (riders) {
    if (rotated) return fail(YPB_ERR_UNSUPPORTED, "riders with the rotated rule");
    if (!riders->ptr && head->batch > 0) return fail(YPB_ERR_INVALID_ARGUMENT, "riders pointer is NULL");
    if (riders->channels < 1 || riders->channels != p->extra)
      return fail(YPB_ERR_INVALID_ARGUMENT, "riders.channels=%d must equal params.extra=%d", riders->channels, p->extra);
    if (riders->kind != YPB_RIDER_RAW && riders->kind != YPB_RIDER_KEYPOINTS) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown rider kind %d", riders->kind);
    if (riders->kind == YPB_RIDER_KEYPOINTS && (riders->kpt_ndim < 2 || riders->channels % riders->kpt_ndim))
      return fail(YPB_ERR_INVALID_ARGUMENT, "keypoint riders: channels=%d not a multiple of ndim=%d", riders->channels, riders->kpt_ndim);
    if (riders->stride_c < g.anchors) return fail(YPB_ERR_INVALID_ARGUMENT, "riders channel stride < anchors");
  } else if (p->extra != (rotated ? 1 : 0)) {
    return fail(YPB_ERR_INVALID_ARGUMENT, "fused path carries extra=%d only (pass riders for more)", rotated ? 1 : 0);
  }
  if (p->rule == YPB_NMS_FAST_BOXIOU) return fail(YPB_ERR_UNSUPPORTED, "FAST_BOXIOU is only reachable through ypb_nms_boxes");
  if (p->nms_box_divisor != 0.f || p->boxes_xyxy || p->conf_per_image)
    return fail(YPB_ERR_UNSUPPORTED, "nms_box_divisor / boxes_xyxy / conf_per_image are served by ypb_nms_from_dense");
  ypb::Workspace w = ypb::carve_workspace(workspace, head->batch, g.anchors, p->rows_cap, p->max_det, p->max_nms, p->rule);
  if (!workspace || w.bytes > workspace_bytes || !aligned(workspace, 256))
    return fail(YPB_ERR_WORKSPACE_TOO_SMALL, "workspace %zu B (256-aligned) needed, %zu given", w.bytes, workspace_bytes);
  if (head->batch == 0) return YPB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  // counters: the caller's clean-on-exit array (no memset node on the full path) or the workspace's own
  int32_t* counters = p->clean_counters ? p->clean_counters : w.row_count;
  if (stage & 1) {
    if (!keep_counters && !(p->clean_counters && stage == 7)) {
      e = cudaMemsetAsync(counters, 0, sizeof(int32_t) * (head->batch + 1), st);
      if (e != cudaSuccess) return cuda_fail(e, "memset row_count");
    }
    ypb::FilterArgs f{};
    f.tile_count = counters + head->batch; f.tile_list = w.tile_list; f.tile_flags = w.tile_flags; f.tile_cap = w.tile_cap;
    f.fuse_decode = split_decode_requested(g) ? 0 : 1;
    f.conf = p->conf_thres; f.nc = p->nc; f.multi_label = p->multi_label; f.rotated = rotated; f.rows_cap = p->rows_cap;
    f.cls_bits = ypb::bits_for(p->nc); f.class_mask = p->class_mask; f.row_count = counters; f.keys = w.keys_a; f.cand_box = w.cand_box; f.cand_ang = w.cand_ang;
    // the persistent TMA-fed scan when asked for and the geometry fits it (16-byte vectorisable levels, nc <= 256), else the LDG kernel
    e = (tma_scan_requested(p->scan_kernel) && !f.fuse_decode) ? ypb::launch_scan_classes_tma(g, head->dtype, f, vec, st) : cudaErrorNotSupported;
    if (e == cudaErrorNotSupported) {
      (void)cudaGetLastError();
      e = ypb::launch_filter_from_head(g, head->dtype, value_dtype, angle, angle_is_logit, f, vec, 1, st);
    }
    if (e != cudaSuccess) return cuda_fail(e, "scan_classes");
  }
  if (stage & 2) {
    ypb::FilterArgs f{};
    f.conf = p->conf_thres; f.nc = p->nc; f.multi_label = p->multi_label; f.rotated = rotated; f.rows_cap = p->rows_cap;
    f.cls_bits = ypb::bits_for(p->nc); f.class_mask = p->class_mask; f.row_count = counters; f.keys = w.keys_a; f.cand_box = w.cand_box; f.cand_ang = w.cand_ang;
    f.tile_count = counters + head->batch; f.tile_list = w.tile_list; f.tile_flags = w.tile_flags; f.tile_cap = w.tile_cap;
    f.fuse_decode = split_decode_requested(g) ? 0 : 1;
    e = ypb::launch_filter_from_head(g, head->dtype, value_dtype, angle, angle_is_logit, f, vec, 2, st);
    if (e != cudaSuccess) return cuda_fail(e, "decode_candidates");
  }
  if (stage & 4) {
    ypb::SuppressArgs s = suppress_args(p, out, w, head->batch, g.anchors);
    s.row_count = counters;
    s.tile_counter = counters + head->batch;
    if (riders) {
      s.rider = riders->ptr; s.rider_dtype = head->dtype; s.rider_kind = riders->kind; s.rider_ndim = riders->kpt_ndim > 0 ? riders->kpt_ndim : 1;
      s.rider_sb = riders->stride_b; s.rider_sc = riders->stride_c;
      s.lv_n = g.num_levels;
      for (int l = 0; l < g.num_levels; ++l) { s.lv_start[l] = g.anchor_start[l]; s.lv_w[l] = g.w[l]; s.lv_stride[l] = g.stride[l]; }
      s.lv_start[g.num_levels] = g.anchors;
    }
    e = ypb::launch_sort_suppress(s, st);
    if (e != cudaSuccess) return cuda_fail(e, "sort_suppress");
  } else if (p->clean_counters && !keep_counters) {
    // a partial (diagnostic) call without the suppression kernel leaves rows counted: restore the clean-on-exit state
    e = cudaMemsetAsync(counters, 0, sizeof(int32_t) * (head->batch + 1), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset row_count");
  }
  return YPB_OK;
}

int ypb_kpts_decode(const ypb_head_desc* head, const void* kpts, int64_t stride_b, int64_t stride_c, int32_t channels,
                    int32_t kpt_ndim, void* out, void* stream) {
  if (!head) return fail(YPB_ERR_INVALID_ARGUMENT, "head descriptor is NULL");
  if (head->num_levels < 1 || head->num_levels > YPB_MAX_LEVELS) return fail(YPB_ERR_INVALID_ARGUMENT, "num_levels=%d invalid", head->num_levels);
  if (!dtype_ok(head->dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown dtype %d", head->dtype);
  if (head->batch < 0 || channels < 1 || kpt_ndim < 2 || channels % kpt_ndim)
    return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d channels=%d ndim=%d invalid", head->batch, channels, kpt_ndim);
  if (head->batch == 0) return YPB_OK;
  if (!kpts || !out) return fail(YPB_ERR_INVALID_ARGUMENT, "kpts / out is NULL");
  const int wide = 16 / static_cast<int>(dtype_size(head->dtype));
  int vec = wide;
  ypb::KptArgs a{};
  a.src = kpts; a.dst = out; a.sb = stride_b; a.sc = stride_c; a.batch = head->batch; a.channels = channels; a.ndim = kpt_ndim;
  a.num_levels = head->num_levels;
  int as = 0;
  for (int l = 0; l < head->num_levels; ++l) {
    if (head->level_h[l] < 1 || head->level_w[l] < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "level %d has empty grid", l);
    if ((head->level_h[l] * head->level_w[l]) % wide) vec = 1;
    a.w[l] = head->level_w[l]; a.stride[l] = head->level_stride[l]; a.anchor_start[l] = as;
    as += head->level_h[l] * head->level_w[l];
  }
  a.anchors = as;
  if (stride_c < as) return fail(YPB_ERR_INVALID_ARGUMENT, "kpts channel stride < anchors");
  if (stride_b % wide || stride_c % wide || !aligned(kpts, 16) || !aligned(out, 16) || as % wide) vec = 1;
  int gs = 0;
  for (int l = 0; l < head->num_levels; ++l) { a.group_start[l] = gs; gs += head->level_h[l] * head->level_w[l] / vec; }
  for (int l = head->num_levels; l <= YPB_MAX_LEVELS; ++l) { a.anchor_start[l] = as; a.group_start[l] = gs; }
  cudaError_t e = ypb::launch_kpts_decode(a, head->dtype, vec, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_kpts_decode");
  return YPB_OK;
}

int ypb_nms_from_dense(const ypb_dense_desc* pred, const ypb_nms_params* p, const ypb_nms_out* out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!pred) return fail(YPB_ERR_INVALID_ARGUMENT, "prediction descriptor is NULL");
  int rc = check_params(p, out);
  if (rc) return rc;
  if (!dtype_ok(pred->dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown dtype %d", pred->dtype);
  if (pred->batch < 0 || pred->anchors < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d anchors=%d invalid", pred->batch, pred->anchors);
  if (pred->channels != 4 + p->nc + p->extra)
    return fail(YPB_ERR_INVALID_ARGUMENT, "channels=%d != 4+nc+extra=%d (nms.py:73-75)", pred->channels, 4 + p->nc + p->extra);
  if (p->rule == YPB_NMS_FAST_BOXIOU) return fail(YPB_ERR_UNSUPPORTED, "FAST_BOXIOU is only reachable through ypb_nms_boxes");
  const bool rotated = p->rule == YPB_NMS_FAST_PROBIOU;
  if (rotated && p->extra < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "rotated rule needs the angle as last channel (nms.py:146)");
  if (ypb::bits_for(pred->anchors) + ypb::bits_for(p->nc) > 31)
    return fail(YPB_ERR_UNSUPPORTED, "anchor and class index do not fit the 31-bit row id");
  if (!pred->ptr && pred->batch > 0) return fail(YPB_ERR_INVALID_ARGUMENT, "prediction pointer is NULL");
  if (pred->anchor_subset && pred->subset_len < 0) return fail(YPB_ERR_INVALID_ARGUMENT, "subset_len=%d invalid", pred->subset_len);
  if (p->nms_box_divisor < 0.f || (p->nms_box_divisor > 0.f && rotated))
    return fail(YPB_ERR_UNSUPPORTED, "normalised NMS boxes (exporter flavour) are built for the axis-aligned rule only");
  ypb::Workspace w = ypb::carve_workspace(workspace, pred->batch, pred->anchors, p->rows_cap, p->max_det, p->max_nms, p->rule);
  if (!workspace || w.bytes > workspace_bytes || !aligned(workspace, 256))
    return fail(YPB_ERR_WORKSPACE_TOO_SMALL, "workspace %zu B (256-aligned) needed, %zu given", w.bytes, workspace_bytes);
  if (pred->batch == 0) return YPB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* counters = p->clean_counters ? p->clean_counters : w.row_count;
  cudaError_t e = cudaSuccess;
  if (!p->clean_counters) {
    e = cudaMemsetAsync(counters, 0, sizeof(int32_t) * pred->batch, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset row_count");
  }
  ypb::FilterArgs f{};
  f.conf = p->conf_thres; f.nc = p->nc; f.multi_label = p->multi_label; f.rotated = rotated; f.rows_cap = p->rows_cap;
  f.cls_bits = ypb::bits_for(p->nc); f.class_mask = p->class_mask; f.row_count = counters; f.keys = w.keys_a; f.cand_box = w.cand_box; f.cand_ang = w.cand_ang;
  f.boxes_xyxy = p->boxes_xyxy; f.conf_per_image = p->conf_per_image;
  e = ypb::launch_filter_from_dense(*pred, f, st);
  if (e != cudaSuccess) return cuda_fail(e, "filter_from_dense");
  ypb::SuppressArgs s = suppress_args(p, out, w, pred->batch, pred->anchors);
  s.row_count = counters;
  s.pred = pred->ptr; s.pred_dtype = pred->dtype; s.pred_sb = pred->stride_b; s.pred_sc = pred->stride_c; s.pred_sa = pred->stride_a;
  e = ypb::launch_sort_suppress(s, st);
  if (e != cudaSuccess) return cuda_fail(e, "sort_suppress");
  return YPB_OK;
}

size_t ypb_nms_boxes_workspace_bytes(int32_t n) {
  if (n < 0) return 0;
  const int m = n > 0 ? n : 1;
  return ypb::carve_workspace(nullptr, 1, m, m, m, m, YPB_NMS_FAST_PROBIOU).bytes + 256;
}

int ypb_nms_boxes(const float* boxes, const float* scores, int32_t n, int32_t box_dim, int32_t rule,
                  float iou_thres_eff, int64_t* keep, int32_t* keep_count, void* workspace, size_t workspace_bytes,
                  void* stream) {
  if (n < 0 || (box_dim != 4 && box_dim != 5)) return fail(YPB_ERR_INVALID_ARGUMENT, "n=%d box_dim=%d invalid", n, box_dim);
  if (rule < YPB_NMS_GREEDY || rule > YPB_NMS_FAST_BOXIOU) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown rule %d", rule);
  if ((rule == YPB_NMS_FAST_PROBIOU) != (box_dim == 5)) return fail(YPB_ERR_INVALID_ARGUMENT, "box_dim 5 <=> FAST_PROBIOU");
  if (!keep || !keep_count) return fail(YPB_ERR_INVALID_ARGUMENT, "keep / keep_count is NULL");
  if (n > 0 && (!boxes || !scores)) return fail(YPB_ERR_INVALID_ARGUMENT, "boxes / scores is NULL");
  const int m = n > 0 ? n : 1;
  ypb::Workspace w = ypb::carve_workspace(workspace, 1, m, m, m, m, YPB_NMS_FAST_PROBIOU);
  if (!workspace || w.bytes > workspace_bytes || !aligned(workspace, 256))
    return fail(YPB_ERR_WORKSPACE_TOO_SMALL, "workspace %zu B (256-aligned) needed, %zu given", w.bytes, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = ypb::launch_boxes_prep(boxes, scores, n, box_dim, w.keys_a, w.cand_box, w.cand_ang, w.row_count, st);
  if (e != cudaSuccess) return cuda_fail(e, "boxes_prep");
  ypb::SuppressArgs s{};
  s.batch = 1; s.anchors = m; s.nc = 1; s.extra = 0; s.max_det = m; s.max_nms = m; s.rule = rule; s.rows_cap = m;
  s.iou_thr = iou_thres_eff; s.max_wh = 0.f; s.cls_bits = 0; s.anchor_bits = ypb::bits_for(m);
  s.row_count = w.row_count; s.keys_a = w.keys_a; s.keys_b = w.keys_b; s.cand_box = w.cand_box;
  s.cand_ang = rule == YPB_NMS_FAST_PROBIOU ? w.cand_ang : nullptr;
  s.kept_box = w.kept_box; s.kept_area = w.kept_area; s.kept_key = w.kept_key; s.rec = w.rec;
  s.out_rows = nullptr; s.out_idx = reinterpret_cast<long long*>(keep); s.out_count = keep_count; s.idx_as_row = 1;
  e = ypb::launch_sort_suppress(s, st);
  if (e != cudaSuccess) return cuda_fail(e, "sort_suppress");
  return YPB_OK;
}

int ypb_compact_results(const float* rows, const int64_t* idx, const int32_t* count, int32_t batch, int32_t max_det,
                        int32_t cols, float* out_rows, int64_t* out_idx, int32_t* out_offsets, void* stream) {
  if (batch < 0 || max_det < 1 || cols < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d max_det=%d cols=%d invalid", batch, max_det, cols);
  if (batch == 0) return YPB_OK;
  if (!count || (out_rows && !rows) || (out_idx && !idx)) return fail(YPB_ERR_INVALID_ARGUMENT, "count / rows / idx is NULL");
  cudaError_t e = ypb::launch_compact_results(rows, reinterpret_cast<const long long*>(idx), count, batch, max_det, cols, out_rows,
                                              reinterpret_cast<long long*>(out_idx), out_offsets, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_compact_results");
  return YPB_OK;
}

int ypb_pairwise_iou(const float* boxes1, int32_t n, const float* boxes2, int32_t m, int32_t box_dim, float* out,
                     void* stream) {
  if (n < 0 || m < 0 || (box_dim != 4 && box_dim != 5)) return fail(YPB_ERR_INVALID_ARGUMENT, "n=%d m=%d box_dim=%d invalid", n, m, box_dim);
  if (n == 0 || m == 0) return YPB_OK;
  if (!boxes1 || !boxes2 || !out) return fail(YPB_ERR_INVALID_ARGUMENT, "boxes1 / boxes2 / out is NULL");
  if (static_cast<long long>(n) * m > (1LL << 38)) return fail(YPB_ERR_UNSUPPORTED, "%d x %d pairs: too large for one launch", n, m);
  cudaError_t e = ypb::launch_pairwise_iou(boxes1, n, boxes2, m, box_dim, out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_pairwise_iou");
  return YPB_OK;
}

int ypb_scale_rows(float* rows, int64_t image_stride, int64_t row_stride, int32_t batch, int32_t rows_per_image,
                   const int32_t* count, const ypb_scale_xform* xforms, const ypb_scale_xform* xform, int32_t box_mode,
                   int32_t flags, int32_t angle_col, float* coords, int64_t coord_image_stride, int64_t coord_row_stride,
                   int32_t nk, int32_t ndim, void* stream) {
  if (batch < 0 || rows_per_image < 0) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d rows_per_image=%d invalid", batch, rows_per_image);
  if (box_mode < YPB_BOXES_NONE || box_mode > YPB_BOXES_REGULARIZE_ONLY) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown box mode %d", box_mode);
  if (box_mode != YPB_BOXES_NONE && !rows && batch * rows_per_image > 0) return fail(YPB_ERR_INVALID_ARGUMENT, "rows is NULL");
  if (!xforms && !xform) return fail(YPB_ERR_INVALID_ARGUMENT, "neither a transform array nor a single transform given");
  if ((box_mode == YPB_BOXES_XYWHR || box_mode == YPB_BOXES_REGULARIZE_ONLY) && angle_col < 4)
    return fail(YPB_ERR_INVALID_ARGUMENT, "angle_col=%d must be >= 4", angle_col);
  if (nk < 0 || (nk > 0 && (ndim < 2 || !coords))) return fail(YPB_ERR_INVALID_ARGUMENT, "nk=%d ndim=%d coords=%p invalid", nk, ndim, (void*)coords);
  if (box_mode == YPB_BOXES_NONE && nk == 0) return YPB_OK;
  ypb::ScaleArgs s{};
  s.rows = rows; s.image_stride = image_stride; s.row_stride = row_stride; s.batch = batch; s.rows_per_image = rows_per_image;
  s.count = count; s.xforms = xforms;
  if (xform) s.xform = *xform;
  s.box_mode = box_mode; s.flags = flags; s.angle_col = angle_col;
  s.coords = coords; s.coord_image_stride = coord_image_stride; s.coord_row_stride = coord_row_stride; s.nk = nk; s.ndim = ndim;
  cudaError_t e = ypb::launch_scale_rows(s, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_scale_rows");
  return YPB_OK;
}

size_t ypb_process_mask_workspace_bytes(int32_t total, int32_t out_h, int32_t out_w) {
  if (total <= 0 || out_h <= 0 || out_w <= 0) return 0;
  const long long tiles = static_cast<long long>((out_w + 127) / 128) * ((out_h + 127) / 128) * total;
  // the work list (count + (tile, image) per tile), 16-byte rounded, then the fill cursor and one "can see its box" flag byte per
  // tile (overlapped form)
  return ((static_cast<size_t>(2 * tiles + 1) * sizeof(int32_t) + 15) & ~static_cast<size_t>(15)) + 16 + static_cast<size_t>(tiles);
}

int ypb_process_mask(const ypb_protos_desc* protos, const float* coeffs, int64_t coef_image_stride, int64_t coef_row_stride,
                     const float* boxes, int64_t box_image_stride, int64_t box_row_stride, const int32_t* offsets,
                     int32_t batch, int32_t total, int32_t out_h, int32_t out_w, int32_t win_top, int32_t win_left,
                     int32_t win_h, int32_t win_w, int32_t crop_mode, float ratio_w, float ratio_h, uint8_t* out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!protos) return fail(YPB_ERR_INVALID_ARGUMENT, "protos descriptor is NULL");
  if (!dtype_ok(protos->dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown dtype %d", protos->dtype);
  if (protos->channels < 1 || protos->mh < 1 || protos->mw < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "empty prototypes");
  if (batch < 1 || total < 0 || out_h < 1 || out_w < 1) return fail(YPB_ERR_INVALID_ARGUMENT, "batch=%d total=%d out=%dx%d invalid", batch, total, out_h, out_w);
  if (batch > 1 && !offsets) return fail(YPB_ERR_INVALID_ARGUMENT, "offsets is NULL for a batch of %d", batch);
  if (crop_mode != YPB_MASK_CROP_PROTO && crop_mode != YPB_MASK_CROP_OUTPUT) return fail(YPB_ERR_INVALID_ARGUMENT, "unknown crop mode %d", crop_mode);
  if (win_top < 0 || win_left < 0 || win_h < 1 || win_w < 1 || win_top + win_h > protos->mh || win_left + win_w > protos->mw)
    return fail(YPB_ERR_INVALID_ARGUMENT, "window (%d,%d,%d,%d) outside the %dx%d prototype grid", win_top, win_left, win_h, win_w, protos->mh, protos->mw);
  if (protos->stride_c < static_cast<int64_t>(protos->mh) * protos->mw) return fail(YPB_ERR_INVALID_ARGUMENT, "prototype channel stride < mh*mw");
  if (total == 0) return YPB_OK;
  if (!protos->ptr || !coeffs || !boxes || !out) return fail(YPB_ERR_INVALID_ARGUMENT, "protos / coeffs / boxes / out is NULL");
  if (total > 65535) return fail(YPB_ERR_UNSUPPORTED, "more than 65535 detections in one call (split the batch)");
  ypb::MaskArgs a{};
  a.protos = protos->ptr; a.proto_dtype = protos->dtype; a.proto_sb = protos->stride_b; a.proto_sc = protos->stride_c;
  a.C = protos->channels; a.mh = protos->mh; a.mw = protos->mw;
  a.coeffs = coeffs; a.coef_image_stride = coef_image_stride; a.coef_row_stride = coef_row_stride;
  a.boxes = boxes; a.box_image_stride = box_image_stride; a.box_row_stride = box_row_stride;
  a.offsets = batch > 1 ? offsets : nullptr; a.batch = batch; a.total = total;
  a.ih = out_h; a.iw = out_w; a.win_top = win_top; a.win_left = win_left; a.win_h = win_h; a.win_w = win_w;
  // ATen area_pixel_compute_scale<float>: static_cast<float>(input_size) / output_size
  a.scale_h = static_cast<float>(win_h) / static_cast<float>(out_h);
  a.scale_w = static_cast<float>(win_w) / static_cast<float>(out_w);
  a.ratio_w = ratio_w; a.ratio_h = ratio_h; a.crop_mode = crop_mode; a.out = out;
  cudaError_t e = ypb::launch_process_mask(a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidConfiguration) return fail(YPB_ERR_UNSUPPORTED, "resize ratio %dx%d -> %dx%d needs too large a shared-memory footprint", win_h, win_w, out_h, out_w);
  if (e != cudaSuccess) return cuda_fail(e, "ypb_process_mask");
  return YPB_OK;
}

int ypb_match_predictions(const float* preds, int64_t pred_image_stride, int64_t pred_row_stride, int32_t cls_col,
                          int32_t batch, int32_t rows_per_image, const int32_t* count, const float* labels,
                          const int32_t* label_offsets, int32_t m, int32_t max_labels, const float* iou, int64_t iou_stride,
                          const float* true_cls, const float* thresholds, int32_t nthr, uint8_t* correct, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (batch < 0 || rows_per_image < 0 || m < 0 || max_labels < 0) return fail(YPB_ERR_INVALID_ARGUMENT, "negative size");
  if (nthr < 1 || nthr > 16 || !thresholds) return fail(YPB_ERR_INVALID_ARGUMENT, "nthr=%d outside [1,16] or thresholds NULL", nthr);
  if (batch == 0 || rows_per_image == 0) return YPB_OK;
  if (!preds || !correct) return fail(YPB_ERR_INVALID_ARGUMENT, "preds / correct is NULL");
  if (iou) {
    if (batch != 1 || label_offsets) return fail(YPB_ERR_INVALID_ARGUMENT, "matrix mode handles one image per call");
    if (!true_cls && m > 0) return fail(YPB_ERR_INVALID_ARGUMENT, "matrix mode needs true_cls");
    if (iou_stride < rows_per_image) return fail(YPB_ERR_INVALID_ARGUMENT, "iou row stride < n");
  } else {
    if (!labels && (m > 0 || label_offsets)) return fail(YPB_ERR_INVALID_ARGUMENT, "labels is NULL");
    if (batch > 1 && !label_offsets) return fail(YPB_ERR_INVALID_ARGUMENT, "label_offsets is NULL for a batch of %d", batch);
    if (cls_col < 4) return fail(YPB_ERR_INVALID_ARGUMENT, "cls_col=%d must be >= 4 in boxes mode", cls_col);
  }
  const int ml = label_offsets ? max_labels : m;
  ypb::MatchArgs a{};
  a.preds = preds; a.pred_image_stride = pred_image_stride; a.pred_row_stride = pred_row_stride; a.cls_col = cls_col;
  a.batch = batch; a.rows_per_image = rows_per_image; a.count = count; a.labels = labels; a.label_offsets = label_offsets;
  a.m = m; a.iou = iou; a.iou_stride = iou_stride; a.true_cls = true_cls; a.nthr = nthr; a.correct = correct;
  for (int i = 0; i < nthr; ++i) a.thr[i] = thresholds[i];
  if (static_cast<size_t>(nthr) * ml * sizeof(int) > 200 * 1024) {
    // sum M is only known on the device in the batched form: the caller sizes the scratch for nthr * (total labels)
    if (!workspace || workspace_bytes < static_cast<size_t>(nthr) * ml * sizeof(int))
      return fail(YPB_ERR_WORKSPACE_TOO_SMALL, "%d labels need a workspace of nthr * total_labels * 4 bytes", ml);
    a.win_global = static_cast<int*>(workspace);
  }
  cudaError_t e = ypb::launch_match_predictions(a, ml, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_match_predictions");
  return YPB_OK;
}

int ypb_peer_wait(const int32_t* flags, int32_t world, int32_t* state, int32_t lag, int32_t depth, int32_t* const* peer_ack,
                  int32_t my_rank, int64_t* slot_index, void* stream) {
  if (!flags || !state || world < 1 || world > YPB_MAX_PEERS || lag < -1) return fail(YPB_ERR_INVALID_ARGUMENT, "flags / state NULL, world=%d or lag=%d invalid", world, lag);
  // in-order consumption (lag -1) beside the next step behaves like lag 1: the entry being read and the one being filled differ
  if (depth < 1 || (peer_ack && depth < (lag < 0 ? 1 : lag) + 2)) return fail(YPB_ERR_INVALID_ARGUMENT, "depth=%d: an acknowledged ring needs depth >= lag + 2 = %d", depth, (lag < 0 ? 1 : lag) + 2);
  if (my_rank < 0 || my_rank >= world) return fail(YPB_ERR_INVALID_ARGUMENT, "my_rank=%d outside [0,%d)", my_rank, world);
  if (peer_ack)
    for (int i = 0; i < world; ++i)
      if (!peer_ack[i]) return fail(YPB_ERR_INVALID_ARGUMENT, "peer_ack[%d] is NULL", i);
  cudaError_t e = ypb::launch_peer_wait(flags, world, state, lag, depth, peer_ack, my_rank, reinterpret_cast<long long*>(slot_index),
                                        static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_peer_wait");
  return YPB_OK;
}

int ypb_peer_wait_copy(const int32_t* flags, int32_t world, int32_t* state, int32_t lag, int32_t depth, int32_t* const* peer_ack,
                       int32_t my_rank, int64_t* slot_index, const float* ring, int64_t entry_floats, float* out, void* stream) {
  if (!flags || !state || world < 1 || world > YPB_MAX_PEERS || lag < -1) return fail(YPB_ERR_INVALID_ARGUMENT, "flags / state NULL, world=%d or lag=%d invalid", world, lag);
  // in-order consumption (lag -1) beside the next step behaves like lag 1: the entry being read and the one being filled differ
  if (depth < 1 || (peer_ack && depth < (lag < 0 ? 1 : lag) + 2)) return fail(YPB_ERR_INVALID_ARGUMENT, "depth=%d: an acknowledged ring needs depth >= lag + 2 = %d", depth, (lag < 0 ? 1 : lag) + 2);
  if (my_rank < 0 || my_rank >= world) return fail(YPB_ERR_INVALID_ARGUMENT, "my_rank=%d outside [0,%d)", my_rank, world);
  if (!ring || !out || entry_floats < 4 || entry_floats % 4 || !aligned(ring, 16) || !aligned(out, 16))
    return fail(YPB_ERR_INVALID_ARGUMENT, "ring / out NULL or not 16-byte aligned, or entry_floats=%lld not a positive multiple of 4", (long long)entry_floats);
  if (peer_ack)
    for (int i = 0; i < world; ++i)
      if (!peer_ack[i]) return fail(YPB_ERR_INVALID_ARGUMENT, "peer_ack[%d] is NULL", i);
  cudaError_t e = ypb::launch_peer_wait_copy(flags, world, state, lag, depth, peer_ack, my_rank, reinterpret_cast<long long*>(slot_index),
                                             ring, entry_floats, out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ypb_peer_wait_copy");
  return YPB_OK;
}

void ypb_debug_set_phase_buffer(void* device_buffer) { ypb::set_phase_buffer(static_cast<long long*>(device_buffer)); }

int ypb_selftest_sigmoid_monotone(int32_t dtype, unsigned long long* violations, void* stream) {
  if (!violations || !dtype_ok(dtype)) return fail(YPB_ERR_INVALID_ARGUMENT, "bad arguments");
  cudaError_t e = ypb::launch_sigmoid_selftest(dtype, violations, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "sigmoid selftest");
  return YPB_OK;
}

}  // extern "C"
