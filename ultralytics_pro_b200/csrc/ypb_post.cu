// Result-side kernels of libyolopost_b200 (sm_100a): what the reference runs on the kept rows right after NMS
// (SURVEY.md section 8f rank 1):
//   scale_rows_kernel    utils/ops.py:102-135 scale_boxes, :152-177 clip_boxes, :621-636 regularize_rboxes,
//                        :562-595 scale_coords, :598-618 clip_coords - applied in place to the (B, max_det, 6+extra) rows
//                        of a whole batch (per-image transform, per-image kept count) or to one box / point set
// These are a few kB of data per batch: the point is ONE launch instead of the reference's 9-13 tiny ATen launches per
// image, with the reference's rounding (every step a separately rounded fp32 operation, true IEEE division).
#include "ypb_common.cuh"

#include <cstdlib>

namespace ypb {

namespace {

__global__ void __launch_bounds__(256)
scale_rows_kernel(const ScaleArgs s) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(s.batch) * s.rows_per_image) return;
  const int b = static_cast<int>(t / s.rows_per_image);
  const int k = static_cast<int>(t - static_cast<long long>(b) * s.rows_per_image);
  if (s.count && k >= s.count[b]) return;
  const ypb_scale_xform xf = s.xforms ? s.xforms[b] : s.xform;
  const bool padding = s.flags & YPB_SCALE_PADDING;

  if (s.box_mode != YPB_BOXES_NONE) {
    float* r = s.rows + static_cast<long long>(b) * s.image_stride + static_cast<long long>(k) * s.row_stride;
    float x0 = r[0], y0 = r[1], x1 = r[2], y1 = r[3];
    scale_box(x0, y0, x1, y1, r + s.angle_col, xf, s.box_mode, padding);
    r[0] = x0; r[1] = y0; r[2] = x1; r[3] = y1;
  }

  if (s.nk > 0) {  // ops.py:562-595 scale_coords on the (nk, ndim) points of the row
    float* p = s.coords + static_cast<long long>(b) * s.coord_image_stride + static_cast<long long>(k) * s.coord_row_stride;
    for (int j = 0; j < s.nk; ++j, p += s.ndim) {
      float x = p[0], y = p[1];
      if (!(s.flags & YPB_SCALE_COORDS_CLIP_ONLY)) {
        if (padding) { x = __fsub_rn(x, xf.cpad_x); y = __fsub_rn(y, xf.cpad_y); }
        x = __fdiv_rn(x, xf.gain); y = __fdiv_rn(y, xf.gain);
      }
      x = torch_clamp(x, 0.f, xf.img_w); y = torch_clamp(y, 0.f, xf.img_h);  // ops.py:598-618
      if (s.flags & YPB_SCALE_NORMALIZE) { x = __fdiv_rn(x, xf.img_w); y = __fdiv_rn(y, xf.img_h); }
      p[0] = x; p[1] = y;
    }
  }
}

// Pose.kpts_decode (head.py:1254-1273) over the whole (B, nk*ndim, A) tensor: thread = VEC consecutive anchors of
// KPT_CH channels (KPT_CH independent 128-bit streaming loads in flight); HBM-bound (2 * B * C * A * s bytes).
constexpr int KPT_CH = 8;

template <int DT, int VEC>
__global__ void __launch_bounds__(256)
kpts_decode_kernel(const __grid_constant__ KptArgs a) {
  using T = typename DType<DT>::type;
  const int grp = blockIdx.x * blockDim.x + threadIdx.x;
  if (grp >= a.group_start[a.num_levels]) return;
  const int c0 = blockIdx.y * KPT_CH, b = blockIdx.z;
  int l = 0;
#pragma unroll
  for (int i = 1; i < YPB_MAX_LEVELS; ++i)
    if (i < a.num_levels && grp >= a.group_start[i]) l = i;
  const int a_local = (grp - a.group_start[l]) * VEC;
  const int a_glob = a.anchor_start[l] + a_local;
  const T* src = static_cast<const T*>(a.src) + static_cast<long long>(b) * a.sb + a_glob;
  T* dst = static_cast<T*>(a.dst) + static_cast<long long>(b) * a.channels * a.anchors + a_glob;
  const int W = a.w[l];
  const float stride = a.stride[l];
  const int gy0 = a_local / W, gx0 = a_local - gy0 * W;
  Pack<T, VEC> p[KPT_CH];
#pragma unroll
  for (int j = 0; j < KPT_CH; ++j)
    if (c0 + j < a.channels) p[j] = load_pack<T, VEC>(src + static_cast<long long>(c0 + j) * a.sc);
#pragma unroll
  for (int j = 0; j < KPT_CH; ++j) {
    const int c = c0 + j;
    if (c < a.channels) {
      const int d = c % a.ndim;
      int gy = gy0, gx = gx0;
      Pack<T, VEC> q;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float ax = DType<DT>::rnd(static_cast<float>(gx) + 0.5f), ay = DType<DT>::rnd(static_cast<float>(gy) + 0.5f);
        q.v[i] = DType<DT>::from_f(kpt_value<DT>(DType<DT>::to_f(p[j].v[i]), d, ax, ay, stride));
        if (++gx == W) { gx = 0; ++gy; }
      }
      store_pack<T, VEC>(dst + static_cast<long long>(c) * a.anchors, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// process_mask / process_mask_native (utils/ops.py:489-541): per kept detection, mask = coeffs . protos (32-term dot
// product per prototype pixel), cropped to the box, bilinearly resized (F.interpolate, align_corners=False) and
// thresholded at 0 -> uint8.  One CTA = one detection x one output tile of MT_W x MT_H pixels:
//   1. the prototype-resolution values the tile's bilinear taps touch (<= (MT_H/scale+2) x (MT_W/scale+2)) are computed into
//      shared memory straight from the (L2-resident) prototypes - the (n, mh, mw) fp32 intermediate of the reference is
//      never written;
//   2. every thread resizes + thresholds 16 consecutive pixels of one row and stores them with one 128-bit store.
// Tiles that cannot see the box are zero-filled without touching the prototypes.  The kernel is bound by the HBM write
// of the (n, H, W) uint8 result: algorithmic bytes = n*H*W.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MT_W = 128, MT_H = 128, MT_THREADS = 256;  // 16 KB of output per CTA: the grid is not CTA-launch bound

// ATen UpSampleKernel / UpSample.h area_pixel_compute_source_index (align_corners=False) + guard_index_and_lambda
__device__ __forceinline__ void bilinear_tap(float scale, int dst, int in_size, int& i0, int& i1, float& w0, float& w1) {
  float src = __fsub_rn(__fmul_rn(scale, static_cast<float>(dst) + 0.5f), 0.5f);
  if (src < 0.f) src = 0.f;
  i0 = min(static_cast<int>(floorf(src)), in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  w1 = fminf(fmaxf(__fsub_rn(src, static_cast<float>(i0)), 0.f), 1.f);
  w0 = __fsub_rn(1.f, w1);
}

// detection d -> (image, row) through the offsets prefix (binary search; offsets[lo] <= d < offsets[hi])
__device__ __forceinline__ void mask_locate(const MaskArgs& a, int d, int& b, int& k) {
  b = 0; k = d;
  if (a.offsets) {
    int lo = 0, hi = a.batch;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.offsets[mid] <= d) lo = mid; else hi = mid; }
    b = lo; k = d - a.offsets[lo];
  }
}

// Can the output tile [X0, X1] x [Y0, Y1] of detection (b, k) see the box at all?  (crop_mask keeps column r iff x1 <= r < x2,
// row c iff y1 <= c < y2, ops.py:464-486).  Exactly the test the tile kernel makes before it computes anything.
__device__ __forceinline__ bool mask_tile_empty(const MaskArgs& a, int b, int k, int X0, int Y0, int X1, int Y1) {
  const float* bp = a.boxes + static_cast<long long>(b) * a.box_image_stride + static_cast<long long>(k) * a.box_row_stride;
  float bx1 = bp[0], by1 = bp[1], bx2 = bp[2], by2 = bp[3];
  if (a.crop_mode == YPB_MASK_CROP_PROTO) {
    bx1 = __fmul_rn(bx1, a.ratio_w); by1 = __fmul_rn(by1, a.ratio_h); bx2 = __fmul_rn(bx2, a.ratio_w); by2 = __fmul_rn(by2, a.ratio_h);
    int ry0, ry1, rx0, rx1, t0, t1;
    float f0, f1;
    bilinear_tap(a.scale_h, Y0, a.win_h, ry0, t1, f0, f1);
    bilinear_tap(a.scale_h, Y1, a.win_h, t0, ry1, f0, f1);
    bilinear_tap(a.scale_w, X0, a.win_w, rx0, t1, f0, f1);
    bilinear_tap(a.scale_w, X1, a.win_w, t0, rx1, f0, f1);
    return !(static_cast<float>(rx1 + a.win_left) >= bx1 && static_cast<float>(rx0 + a.win_left) < bx2 &&
             static_cast<float>(ry1 + a.win_top) >= by1 && static_cast<float>(ry0 + a.win_top) < by2);
  }
  return !(static_cast<float>(X1) >= bx1 && static_cast<float>(X0) < bx2 && static_cast<float>(Y1) >= by1 && static_cast<float>(Y0) < by2);
}

// Work list of the two-step form: one thread per (detection, tile); tiles that can see their box are appended (unordered) to
// list[1..] as (tile id, image) pairs, list[0] = their number.  Everything else of the (total, H, W) result is zero: a plain memset writes it at the
// write-only ceiling of the memory system and no CTA is spent on it.
__global__ void __launch_bounds__(256)
mask_tile_list_kernel(const __grid_constant__ MaskArgs a, int tiles_x, int tiles_y, int32_t* __restrict__ list,
                      uint8_t* __restrict__ tile_live) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per = tiles_x * tiles_y;
  bool live = false;
  int live_b = 0;
  const bool in_range = t < static_cast<long long>(a.total) * per;
  if (in_range) {
    const int d = static_cast<int>(t / per), r = static_cast<int>(t - static_cast<long long>(d) * per);
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    int b, k;
    mask_locate(a, d, b, k);
    const int X0 = tx * MT_W, Y0 = ty * MT_H;
    live = !mask_tile_empty(a, b, k, X0, Y0, min(X0 + MT_W, a.iw) - 1, min(Y0 + MT_H, a.ih) - 1);
    live_b = b;
  }
  if (tile_live && in_range) tile_live[t] = live ? 1 : 0;  // the overlapped form's fill warps skip the listed tiles
  if (tile_live && t == 0) *reinterpret_cast<int32_t*>(tile_live - 16) = 0;  // cursor of the fill chunks, 16 bytes in front of the flags
  const unsigned m = __ballot_sync(0xffffffffu, live);
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(&list[0], __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (live) {  // entry = (tile id, image): the tile kernel then needs no search through the offsets
    const int e = base + __popc(m & ((1u << lane) - 1u));
    list[1 + 2 * e] = static_cast<int32_t>(t);
    list[2 + 2 * e] = live_b;
  }
}

// barrier of the MT_THREADS tile-computing threads only (named barrier 1): the overlapped form runs one more warp per CTA
// that never joins them
__device__ __forceinline__ void mt_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MT_THREADS) : "memory"); }

template <int DT>
__device__ __forceinline__ void mask_tile(const MaskArgs& a, float* sm_f, int* s_img, int d, int X0, int Y0, bool fill_empty,
                                          int b_known = -1) {
  using T = typename DType<DT>::type;
  float* coef = sm_f;                 // [C]
  float4* xtap = reinterpret_cast<float4*>(sm_f + ((a.C + 3) & ~3));  // [MT_W] (x0, x1 as int bits, w0, w1) of every tile column
  float4* ytap = xtap + MT_W;                                          // [MT_H] the same for every tile row
  float* reg = reinterpret_cast<float*>(ytap + MT_H);  // [rh][rw] prototype-resolution values of this tile's footprint
  const int tid = threadIdx.x;
  const int X1 = min(X0 + MT_W, a.iw) - 1, Y1 = min(Y0 + MT_H, a.ih) - 1;  // inclusive

  // detection -> (image, row)
  int b = 0, k = d;
  if (a.offsets && b_known >= 0) {
    b = b_known; k = d - a.offsets[b];
  } else if (a.offsets) {
    if (tid == 0) {
      int lo = 0, hi = a.batch;  // offsets[lo] <= d < offsets[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.offsets[mid] <= d) lo = mid; else hi = mid; }
      s_img[0] = lo; s_img[1] = d - a.offsets[lo];
    }
    mt_sync();
    b = s_img[0]; k = s_img[1];
  }
  const float* bp = a.boxes + static_cast<long long>(b) * a.box_image_stride + static_cast<long long>(k) * a.box_row_stride;
  float bx1 = bp[0], by1 = bp[1], bx2 = bp[2], by2 = bp[3];
  if (a.crop_mode == YPB_MASK_CROP_PROTO) {  // ops.py:505-510: boxes * (mw/W, mh/H, mw/W, mh/H)
    bx1 = __fmul_rn(bx1, a.ratio_w); by1 = __fmul_rn(by1, a.ratio_h); bx2 = __fmul_rn(bx2, a.ratio_w); by2 = __fmul_rn(by2, a.ratio_h);
  }

  // footprint of the tile in the (windowed) prototype grid
  int ry0, ry1, rx0, rx1, t0, t1;
  float f0, f1;
  bilinear_tap(a.scale_h, Y0, a.win_h, ry0, t1, f0, f1);
  bilinear_tap(a.scale_h, Y1, a.win_h, t0, ry1, f0, f1);
  bilinear_tap(a.scale_w, X0, a.win_w, rx0, t1, f0, f1);
  bilinear_tap(a.scale_w, X1, a.win_w, t0, rx1, f0, f1);
  const int rh = ry1 - ry0 + 1, rw = rx1 - rx0 + 1;

  // can the tile see the box at all?  (crop_mask keeps column r iff x1 <= r < x2, row c iff y1 <= c < y2, ops.py:464-486)
  bool empty;
  if (a.crop_mode == YPB_MASK_CROP_PROTO)
    empty = !(static_cast<float>(rx1 + a.win_left) >= bx1 && static_cast<float>(rx0 + a.win_left) < bx2 &&
              static_cast<float>(ry1 + a.win_top) >= by1 && static_cast<float>(ry0 + a.win_top) < by2);
  else
    empty = !(static_cast<float>(X1) >= bx1 && static_cast<float>(X0) < bx2 && static_cast<float>(Y1) >= by1 && static_cast<float>(Y0) < by2);

  uint8_t* out = a.out + static_cast<long long>(d) * a.ih * a.iw;
  const int row = tid >> 3, seg = tid & 7;  // 32 rows x 8 segments of 16 pixels per pass, MT_H / 32 passes
  const int XS = X0 + seg * 16;
  const bool vec_ok = (a.iw & 15) == 0;

  if (!empty) {
    const float* cp = a.coeffs + static_cast<long long>(b) * a.coef_image_stride + static_cast<long long>(k) * a.coef_row_stride;
    for (int c = tid; c < a.C; c += MT_THREADS) coef[c] = cp[c];
    if (tid < MT_W) {  // bilinear taps of every column / row of the tile, computed once
      int i0, i1; float w0, w1;
      bilinear_tap(a.scale_w, min(X0 + tid, a.iw - 1), a.win_w, i0, i1, w0, w1);
      xtap[tid] = make_float4(__int_as_float(i0), __int_as_float(i1), w0, w1);
    } else if (tid < MT_W + MT_H) {
      int i0, i1; float w0, w1;
      bilinear_tap(a.scale_h, min(Y0 + tid - MT_W, a.ih - 1), a.win_h, i0, i1, w0, w1);
      ytap[tid - MT_W] = make_float4(__int_as_float(i0), __int_as_float(i1), w0, w1);
    }
    mt_sync();
    const T* pr = static_cast<const T*>(a.protos) + static_cast<long long>(b) * a.proto_sb;
    // the part of the footprint whose values are needed: all of it, or (PROTO crop, ops.py:486 masks * bool: outside the box
    // the product is zero whatever the mask value) its intersection with the box; the rest is zero.  The needed pixels are
    // dealt evenly over the threads, and all channel loads of a pixel are in flight together (32 when C is a multiple of 32):
    // this phase is bound by the latency of its L2 loads, i.e. by round trips per thread.
    int fx0 = 0, fy0 = 0, fw = rw, fh = rh;
    if (a.crop_mode == YPB_MASK_CROP_PROTO) {
      // pixel (px, py) is inside iff bx1 <= px < bx2 and by1 <= py < by2 (floats compared with integers-as-floats)
      // (bounds clamped before the conversion: infinities must not overflow the integer arithmetic; a NaN bound yields an
      // empty range, as every comparison with it is false)
      auto icl = [](float v) { return static_cast<int>(ceilf(fminf(fmaxf(v, -1.0e6f), 1.0e6f))); };
      const bool box_ok = bx1 == bx1 && bx2 == bx2 && by1 == by1 && by2 == by2;
      int lx = icl(bx1) - (rx0 + a.win_left), hx = box_ok ? icl(bx2) - (rx0 + a.win_left) : lx;  // [lx, hx)
      int ly = icl(by1) - (ry0 + a.win_top), hy = box_ok ? icl(by2) - (ry0 + a.win_top) : ly;
      lx = max(lx, 0); ly = max(ly, 0); hx = min(hx, rw); hy = min(hy, rh);
      fx0 = lx; fy0 = ly; fw = max(hx - lx, 0); fh = max(hy - ly, 0);
      for (int i = tid; i < rh * rw; i += MT_THREADS) {
        const int y = i / rw, x = i - y * rw;
        if (x < fx0 || x >= fx0 + fw || y < fy0 || y >= fy0 + fh) reg[i] = 0.f;
      }
    }
    for (int i = tid; i < fw * fh; i += MT_THREADS) {
      const int yy = i / fw, y = fy0 + yy, x = fx0 + (i - yy * fw);
      const int py = ry0 + y + a.win_top, px = rx0 + x + a.win_left;
      const T* p = pr + static_cast<long long>(py) * a.mw + px;
      float acc = 0.f;
      int c = 0;
      for (; c + 32 <= a.C; c += 32) {  // 32 independent loads in flight, then the sum in channel order
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = DType<DT>::to_f(p[static_cast<long long>(c + j) * a.proto_sc]);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(coef[c + j], v[j], acc);
      }
      for (; c + 16 <= a.C; c += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = DType<DT>::to_f(p[static_cast<long long>(c + j) * a.proto_sc]);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fmaf(coef[c + j], v[j], acc);
      }
      for (; c < a.C; ++c) acc = fmaf(coef[c], DType<DT>::to_f(p[static_cast<long long>(c) * a.proto_sc]), acc);
      reg[y * rw + x] = acc;
    }
    mt_sync();
    if (a.h_cap > 0) {
      // horizontally interpolated footprint rows: H[y][c] = w0(c) * reg[y][x0(c)] + w1(c) * reg[y][x1(c)] - the inner sums of
      // ATen's h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11), which depend on (footprint row, output column) only.
      // Every footprint row serves 1/scale output rows as their upper and as their lower row: computing the sums once per
      // (row, column) instead of once per output pixel leaves 2 loads + 3 operations per pixel (the same operations on
      // the same operands in the same order: bit-identical).
      // (restricting the table to the rows / segments that can see the box was measured: no gain on the C2-like workload, 11 %
      // slower for whole-image boxes - the test costs what it saves)
      float* H = reg + a.reg_cap;
      for (int i = tid; i < rh * MT_W; i += MT_THREADS) {
        const int y = i / MT_W, c = i - y * MT_W;
        const float4 tx = xtap[c];
        const float* r = reg + y * rw - rx0;
        H[i] = __fadd_rn(__fmul_rn(tx.z, r[__float_as_int(tx.x)]), __fmul_rn(tx.w, r[__float_as_int(tx.y)]));
      }
      mt_sync();
    }
  }

  if (XS > X1) return;
  if (empty && !fill_empty) return;  // two-step form: the memset already wrote the zeros
  const int XE = min(XS + 15, X1);
  // a segment / row whose taps all fall outside the crop box is zero (PROTO mode: the footprint values are zero there)
  bool seg_live = !empty;
  if (seg_live && a.crop_mode == YPB_MASK_CROP_PROTO) {
    const float4 tl = xtap[XS - X0], tr = xtap[XE - X0];
    seg_live = static_cast<float>(__float_as_int(tr.y) + a.win_left) >= bx1 && static_cast<float>(__float_as_int(tl.x) + a.win_left) < bx2;
  } else if (seg_live) {
    seg_live = static_cast<float>(XE) >= bx1 && static_cast<float>(XS) < bx2;
  }
  for (int Y = Y0 + row; Y <= Y1; Y += MT_THREADS / 8) {
    uint32_t w4[4] = {0u, 0u, 0u, 0u};
    bool live = seg_live;
    float4 ty = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      ty = ytap[Y - Y0];
      if (a.crop_mode == YPB_MASK_CROP_PROTO)
        live = static_cast<float>(__float_as_int(ty.y) + a.win_top) >= by1 && static_cast<float>(__float_as_int(ty.x) + a.win_top) < by2;
      else
        live = static_cast<float>(Y) >= by1 && static_cast<float>(Y) < by2;
    }
    if (live && a.h_cap > 0) {
      const float wy0 = ty.z, wy1 = ty.w;
      const float4* h0 = reinterpret_cast<const float4*>(reg + a.reg_cap + (__float_as_int(ty.x) - ry0) * MT_W + seg * 16);
      const float4* h1 = reinterpret_cast<const float4*>(reg + a.reg_cap + (__float_as_int(ty.y) - ry0) * MT_W + seg * 16);
      // the 8 threads of a row read 64-byte apart: thread `seg` starts at 16-byte chunk (seg / 2) of its segment and wraps, so
      // that a quarter-warp's 128-bit loads cover all 32 banks
#pragma unroll
      for (int q0 = 0; q0 < 4; ++q0) {
        const int q = (q0 + (seg >> 1)) & 3;
        const float4 t4 = h0[q], b4 = h1[q];
        const float tv[4] = {t4.x, t4.y, t4.z, t4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        uint32_t bits = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int X = XS + 4 * q + e;
          float v = __fadd_rn(__fmul_rn(wy0, tv[e]), __fmul_rn(wy1, bv[e]));
          if (a.crop_mode == YPB_MASK_CROP_OUTPUT) {
            const bool in = static_cast<float>(X) >= bx1 && static_cast<float>(X) < bx2;
            v = __fmul_rn(v, in ? 1.f : 0.f);
          }
          if (v > 0.f && X <= X1) bits |= 1u << (8 * e);  // ops.py:513 masks.gt_(0.0).byte()
        }
#pragma unroll
        for (int z = 0; z < 4; ++z)
          if (z == q) w4[z] = bits;
      }
    } else if (live) {
      const float wy0 = ty.z, wy1 = ty.w;
      const float* r0 = reg + (__float_as_int(ty.x) - ry0) * rw - rx0;
      const float* r1 = reg + (__float_as_int(ty.y) - ry0) * rw - rx0;
      // The 8 threads of a row sit 16 pixels apart: walking their segments in step would hit the tap table at a stride of
      // 16 entries (256 B: an 8-way shared-memory bank conflict).  Thread `seg` therefore starts at pixel `seg` of its
      // segment and wraps - a stride of 17 entries, conflict-free; measured 3.1 ms -> see profiles for whole-image boxes.
      unsigned long long lo = 0ull, hi = 0ull;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = (i + seg) & 15;
        const int X = XS + j;
        if (X > X1) continue;
        const float4 tx = xtap[X - X0];
        const int x0 = __float_as_int(tx.x), x1 = __float_as_int(tx.y);
        // ATen upsample_bilinear2d: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
        const float top = __fadd_rn(__fmul_rn(tx.z, r0[x0]), __fmul_rn(tx.w, r0[x1]));
        const float bot = __fadd_rn(__fmul_rn(tx.z, r1[x0]), __fmul_rn(tx.w, r1[x1]));
        float v = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(wy1, bot));
        if (a.crop_mode == YPB_MASK_CROP_OUTPUT) {
          const bool in = static_cast<float>(X) >= bx1 && static_cast<float>(X) < bx2;
          v = __fmul_rn(v, in ? 1.f : 0.f);
        }
        if (v > 0.f) {  // ops.py:513 masks.gt_(0.0).byte()
          if (j < 8) lo |= 1ull << (8 * j); else hi |= 1ull << (8 * (j - 8));
        }
      }
      w4[0] = static_cast<uint32_t>(lo); w4[1] = static_cast<uint32_t>(lo >> 32);
      w4[2] = static_cast<uint32_t>(hi); w4[3] = static_cast<uint32_t>(hi >> 32);
    }
    uint8_t* o = out + static_cast<long long>(Y) * a.iw + XS;
    if (vec_ok && XS + 15 <= X1) {
      __stcs(reinterpret_cast<uint4*>(o), make_uint4(w4[0], w4[1], w4[2], w4[3]));
    } else {
      for (int i = 0; i < 16 && XS + i <= X1; ++i) o[i] = static_cast<uint8_t>((w4[i >> 2] >> ((i & 3) * 8)) & 0xffu);
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(MT_THREADS, 3)
process_mask_kernel(const __grid_constant__ MaskArgs a) {
  extern __shared__ __align__(16) float sm_f[];
  __shared__ int s_img[4];
  mask_tile<DT>(a, sm_f, s_img, blockIdx.z, blockIdx.x * MT_W, blockIdx.y * MT_H, true);
}

// Two-step form: grid-stride over the work list of mask_tile_list_kernel (only tiles that can see their box).
template <int DT>
__global__ void __launch_bounds__(MT_THREADS, 3)
process_mask_list_kernel(const __grid_constant__ MaskArgs a, int tiles_x, int tiles_y, const int32_t* __restrict__ list) {
  extern __shared__ __align__(16) float sm_f[];
  __shared__ int s_img[4];
  const int n = list[0], per = tiles_x * tiles_y;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int t = list[1 + 2 * e], tb = list[2 + 2 * e];
    const int d = t / per, r = t - d * per;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    mask_tile<DT>(a, sm_f, s_img, d, tx * MT_W, ty * MT_H, false, tb);
    mt_sync();  // shared memory is reused by the next tile
  }
}

// Overlapped form: the zero fill and the tile computation in ONE kernel, on different warps.  Every CTA carries one extra
// warp (threads MT_THREADS..MT_THREADS+31) that zero-fills the tiles that cannot see their box with 128-bit streaming stores - a
// posted-store stream that needs next to no issue slots - while the MT_THREADS other threads work through the list of tiles
// that can (prototype dot products from L2, bilinear taps: bound by instruction issue, L2 latency and shared memory).  The two
// halves of the two-step form (memset at the write ceiling, then compute) ran one after the other; here the write stream hides
// under the arithmetic.  The fill is handed out DYNAMICALLY in chunks of 32 consecutive tiles (one atomicAdd, one coalesced
// read of 32 flag bytes per chunk), so a fill warp whose SM is busy computing simply takes fewer chunks.
// Needs out_w % 16 == 0 (vector stores); else the two-step form.
__device__ __forceinline__ void mask_fill_chunks(const MaskArgs& a, int tiles_x, int per, long long ntiles,
                                                 const uint8_t* __restrict__ tile_live, int32_t* cursor, int lane) {
  const int rsub = lane >> 3, seg = lane & 7;  // 4 rows x 8 segments of 16 bytes per store instruction
  for (;;) {
    int c0 = 0;
    if (lane == 0) c0 = atomicAdd(cursor, 32);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    if (c0 >= ntiles) break;
    const long long mine = static_cast<long long>(c0) + lane;
    unsigned todo = __ballot_sync(0xffffffffu, mine < ntiles && tile_live[mine] == 0);
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const long long t = static_cast<long long>(c0) + j;
      const int d = static_cast<int>(t / per), r = static_cast<int>(t - static_cast<long long>(d) * per);
      const int ty = r / tiles_x, tx = r - ty * tiles_x;
      const int X0 = tx * MT_W, Y0 = ty * MT_H;
      const int rows = min(MT_H, a.ih - Y0), cols = min(MT_W, a.iw - X0);
      if (seg * 16 >= cols) continue;  // out_w % 16 == 0: a segment is inside or outside as a whole
      uint8_t* o = a.out + (static_cast<long long>(d) * a.ih + Y0 + rsub) * a.iw + X0 + seg * 16;
      const long long step = 4ll * a.iw;
      for (int y = rsub; y < rows; y += 4, o += step) __stcs(reinterpret_cast<uint4*>(o), make_uint4(0u, 0u, 0u, 0u));
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(MT_THREADS + 32, 3)
process_mask_overlap_kernel(const __grid_constant__ MaskArgs a, int tiles_x, int tiles_y, const int32_t* __restrict__ list,
                            const uint8_t* __restrict__ tile_live, int dbg) {
  extern __shared__ __align__(16) float sm_f[];
  __shared__ int s_img[4];
  const int per = tiles_x * tiles_y;
  const long long ntiles = static_cast<long long>(a.total) * per;
  int32_t* cursor = reinterpret_cast<int32_t*>(const_cast<uint8_t*>(tile_live) - 16);
  if (threadIdx.x >= MT_THREADS) {
    if (dbg != 1) mask_fill_chunks(a, tiles_x, per, ntiles, tile_live, cursor, threadIdx.x - MT_THREADS);
    return;
  }
  if (dbg == 2) return;
  const int n = list[0];
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int t = list[1 + 2 * e], tb = list[2 + 2 * e];
    const int d = t / per, r = t - d * per;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    mask_tile<DT>(a, sm_f, s_img, d, tx * MT_W, ty * MT_H, false, tb);
    mt_sync();  // shared memory is reused by the next tile
  }
  // (letting the computing warps join the fill once their tiles are done was measured: equal on the C2-like workload, SLOWER
  // when nothing is visible - 0.23 vs 0.16 ms: 27 instead of 3 store streams per SM scatter the DRAM writes)
}

// ---------------------------------------------------------------------------------------------------------------
// Validator matching (engine/validator.py:267-307 match_predictions, non-scipy branch, fed by detect/val.py:274-288
// _process_batch with metrics.py:54 box_iou).  Restated: with iou'[l, d] = iou[l, d] * (cls_l == cls_d),
//   best(d) = the label with the largest iou'[., d]            (first np.unique: one label per detection)
//   for a threshold t, detection d is a true positive iff iou'[best(d), d] >= t and d is the LOWEST-index detection among
//   those whose best label is best(d) and that reach t         (second np.unique: one detection per label, in d order)
// One CTA per image; thread = detection; the per-(threshold, label) winners are an atomicMin table in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MATCH_THREADS = 256;

__device__ __forceinline__ float box_iou_eps(const float4 a, const float4 b) {  // metrics.py:54-75, eps = 1e-7
  const float iw = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-7f));
}

__global__ void __launch_bounds__(MATCH_THREADS)
match_predictions_kernel(const __grid_constant__ MatchArgs a) {
  extern __shared__ int s_win[];  // [nthr][m]: lowest detection index claiming label l at threshold i
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = a.count ? min(a.count[b], a.rows_per_image) : a.rows_per_image;
  const int l0 = a.label_offsets ? a.label_offsets[b] : 0;
  const int m = a.label_offsets ? a.label_offsets[b + 1] - l0 : a.m;
  uint8_t* out = a.correct + static_cast<long long>(b) * a.rows_per_image * a.nthr;
  int* win = a.win_global ? a.win_global + static_cast<long long>(l0) * a.nthr : s_win;
  for (int i = tid; i < a.nthr * m; i += MATCH_THREADS) win[i] = 0x7fffffff;
  __syncthreads();
  const float* labels = a.labels + static_cast<long long>(l0) * 5;  // cls, x1, y1, x2, y2
  for (int d0 = 0; d0 < n; d0 += MATCH_THREADS) {
    const int d = d0 + tid;
    float best = 0.f;
    int bl = -1;
    if (d < n) {
      const float* pr = a.preds + static_cast<long long>(b) * a.pred_image_stride + static_cast<long long>(d) * a.pred_row_stride;
      const float4 pb = a.iou ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(pr[0], pr[1], pr[2], pr[3]);
      const float pc = pr[a.cls_col];
      for (int l = 0; l < m; ++l) {
        float v;
        if (a.iou) {
          v = a.iou[static_cast<long long>(l) * a.iou_stride + d];
        } else {
          const float* g = labels + l * 5;
          v = box_iou_eps(make_float4(g[1], g[2], g[3], g[4]), pb);
        }
        const float tc = a.true_cls ? a.true_cls[l] : labels[l * 5];
        v = __fmul_rn(v, tc == pc ? 1.f : 0.f);  // validator.py:285
        if (v > best) { best = v; bl = l; }
      }
      if (bl >= 0)
        for (int i = 0; i < a.nthr; ++i)
          if (best >= a.thr[i]) atomicMin(&win[i * m + bl], d);
    }
    __syncthreads();
    if (d < n)
      for (int i = 0; i < a.nthr; ++i) out[d * a.nthr + i] = (bl >= 0 && best >= a.thr[i] && win[i * m + bl] == d) ? 1 : 0;
  }
}

}  // namespace

cudaError_t launch_match_predictions(const MatchArgs& a, int max_labels, cudaStream_t st) {
  if (a.batch <= 0 || a.rows_per_image <= 0) return cudaSuccess;
  const size_t smem = a.win_global ? 0 : static_cast<size_t>(a.nthr) * (max_labels > 0 ? max_labels : 1) * sizeof(int);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(match_predictions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  match_predictions_kernel<<<a.batch, MATCH_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

static cudaError_t launch_process_mask_sized(const MaskArgs& a, size_t smem, void* workspace, size_t workspace_bytes, cudaStream_t st);

cudaError_t launch_process_mask(const MaskArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (a.total <= 0 || a.ih <= 0 || a.iw <= 0) return cudaSuccess;
  // worst-case footprint of a tile in the prototype grid
  auto span = [](float scale, int n_out, int tile, int in) {
    long long s = static_cast<long long>(scale * tile) + 3;
    (void)n_out;
    return static_cast<int>(s < in ? s : in);
  };
  const int rh = span(a.scale_h, a.ih, MT_H, a.win_h), rw = span(a.scale_w, a.iw, MT_W, a.win_w);
  size_t smem = (static_cast<size_t>((a.C + 3) & ~3) + static_cast<size_t>(rh) * rw) * sizeof(float) + (MT_W + MT_H) * sizeof(float4);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  // room for the horizontally interpolated rows (rh x MT_W floats) while three CTAs still fit an SM?  (always, for upsampling by
  // >= 2; a near-1:1 resize keeps the per-pixel form)
  MaskArgs a2 = a;
  a2.reg_cap = (rh * rw + 3) & ~3;
  a2.h_cap = 0;
  {
    const size_t with_h = (static_cast<size_t>((a.C + 3) & ~3) + a2.reg_cap + static_cast<size_t>(rh) * MT_W) * sizeof(float) + (MT_W + MT_H) * sizeof(float4);
    static const bool no_h = [] { const char* e = std::getenv("YPB_MASK_NO_HROWS"); return e && e[0] == '1'; }();
    if (with_h <= 72 * 1024 && !no_h) { a2.h_cap = rh; smem = with_h; }
  }
  return launch_process_mask_sized(a2, smem, workspace, workspace_bytes, st);
}

static cudaError_t launch_process_mask_sized(const MaskArgs& a, size_t smem, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int tiles_x = (a.iw + MT_W - 1) / MT_W, tiles_y = (a.ih + MT_H - 1) / MT_H;
  const long long ntiles = static_cast<long long>(a.total) * tiles_x * tiles_y;
  // two-step form when the caller lends a work list: memset the result (the zeros of every tile that cannot see its box are
  // written at the write-only ceiling, no CTA spent on them), list the tiles that can, compute only those.
  // Overlapped form when the workspace also holds one flag byte per tile and rows are 16-byte multiples: no memset - an
  // extra warp per CTA writes the zeros of the unlisted tiles WHILE the others compute the listed ones (YPB_MASK_TWO_STEP=1 keeps
  // the two-step form for comparison).
  const bool two_step = workspace && workspace_bytes >= (2 * static_cast<size_t>(ntiles) + 1) * sizeof(int32_t) && ntiles < (1ll << 30);
  static const bool force_two_step = [] { const char* e = std::getenv("YPB_MASK_TWO_STEP"); return e && e[0] == '1'; }();
  const size_t list_bytes = ((2 * static_cast<size_t>(ntiles) + 1) * sizeof(int32_t) + 15) & ~static_cast<size_t>(15);
  const bool overlap = two_step && !force_two_step && workspace_bytes >= list_bytes + 16 + static_cast<size_t>(ntiles) && (a.iw & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
  dim3 grid(tiles_x, tiles_y, a.total);
  if (!two_step && (grid.z > 65535u || grid.y > 65535u)) return cudaErrorInvalidConfiguration;
  int32_t* list = static_cast<int32_t*>(workspace);
  uint8_t* tile_live = overlap ? static_cast<uint8_t*>(workspace) + list_bytes + 16 : nullptr;  // 16 bytes in front: the fill cursor
  if (two_step) {
    cudaError_t e = cudaSuccess;
    if (!overlap) e = cudaMemsetAsync(a.out, 0, static_cast<size_t>(a.total) * a.ih * a.iw, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(list, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    mask_tile_list_kernel<<<static_cast<unsigned>((ntiles + 255) / 256), 256, 0, st>>>(a, tiles_x, tiles_y, list, tile_live);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const unsigned list_grid = static_cast<unsigned>(ntiles < 148 * 6 ? ntiles : 148 * 6);
  static const int mask_dbg = [] { const char* e = std::getenv("YPB_MASK_DBG"); return e ? std::atoi(e) : 0; }();  // timing diagnostics only
  const unsigned overlap_grid = static_cast<unsigned>(ntiles < 148 * 3 ? ntiles : 148 * 3);  // all CTAs resident: fill warps on every SM from the start
#define YPB_PM(DT)                                                                                                  \
  do {                                                                                                              \
    if (smem > 48 * 1024) {                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(process_mask_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                               \
      e = cudaFuncSetAttribute(process_mask_list_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                               \
      e = cudaFuncSetAttribute(process_mask_overlap_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                               \
    }                                                                                                               \
    if (overlap) process_mask_overlap_kernel<DT><<<overlap_grid, MT_THREADS + 32, smem, st>>>(a, tiles_x, tiles_y, list, tile_live, mask_dbg); \
    else if (two_step) process_mask_list_kernel<DT><<<list_grid, MT_THREADS, smem, st>>>(a, tiles_x, tiles_y, list); \
    else process_mask_kernel<DT><<<grid, MT_THREADS, smem, st>>>(a);                                                \
  } while (0)
  if (a.proto_dtype == YPB_F32) YPB_PM(YPB_F32);
  else if (a.proto_dtype == YPB_F16) YPB_PM(YPB_F16);
  else YPB_PM(YPB_BF16);
#undef YPB_PM
  return cudaGetLastError();
}

cudaError_t launch_kpts_decode(const KptArgs& a, int dtype, int vec, cudaStream_t st) {
  const int groups = a.group_start[a.num_levels];
  if (groups <= 0 || a.batch <= 0 || a.channels <= 0) return cudaSuccess;
  dim3 grid((groups + 255) / 256, (a.channels + KPT_CH - 1) / KPT_CH, a.batch);
#define YPB_KPT(DT, V) kpts_decode_kernel<DT, V><<<grid, 256, 0, st>>>(a)
  if (dtype == YPB_F32) { if (vec == 4) YPB_KPT(YPB_F32, 4); else YPB_KPT(YPB_F32, 1); }
  else if (dtype == YPB_F16) { if (vec == 8) YPB_KPT(YPB_F16, 8); else YPB_KPT(YPB_F16, 1); }
  else { if (vec == 8) YPB_KPT(YPB_BF16, 8); else YPB_KPT(YPB_BF16, 1); }
#undef YPB_KPT
  return cudaGetLastError();
}

cudaError_t launch_scale_rows(const ScaleArgs& s, cudaStream_t st) {
  const long long n = static_cast<long long>(s.batch) * s.rows_per_image;
  if (n <= 0) return cudaSuccess;
  const int blocks = static_cast<int>((n + 255) / 256);
  scale_rows_kernel<<<blocks, 256, 0, st>>>(s);
  return cudaGetLastError();
}

}  // namespace ypb
