// Shared device helpers and the internal launch interface of libyolopost_b200.
// sm_100a only; no CPU path.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/yolopost_b200.h"

#ifndef YPB_LOAD_MODE
#define YPB_LOAD_MODE 3  // ld.global.nc.L1::no_allocate: read-only streaming data, free to overtake the kernel's own stores
#endif

namespace ypb {

// ---------------------------------------------------------------------------------------------------------------
// dtype traits: storage type, widening to fp32 and "round an fp32 value through the storage type" (what a torch op
// on a T tensor does to its fp32 opmath result).
// ---------------------------------------------------------------------------------------------------------------
template <int DT> struct DType;
template <> struct DType<YPB_F32> {
  using type = float;
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <> struct DType<YPB_F16> {
  using type = __half;
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct DType<YPB_BF16> {
  using type = __nv_bfloat16;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// packed x2 view of the 16-bit types (class scan of bf16/fp16 heads)
template <int DT> struct Packed2;
template <> struct Packed2<YPB_F16> {
  using type = __half2;
  static __device__ __forceinline__ __half2 neg_inf() { return __half2half2(__ushort_as_half(0xFC00)); }
  static __device__ __forceinline__ float lo(__half2 v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(__half2 v) { return __high2float(v); }
};
template <> struct Packed2<YPB_BF16> {
  using type = __nv_bfloat162;
  static __device__ __forceinline__ __nv_bfloat162 neg_inf() { return __bfloat162bfloat162(__ushort_as_bfloat16(0xFF80)); }
  static __device__ __forceinline__ float lo(__nv_bfloat162 v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(__nv_bfloat162 v) { return __high2float(v); }
};
template <> struct Packed2<YPB_F32> { using type = float2; };

// A VEC-wide, naturally aligned group of T moved with one LDG/STG (128-bit for fp32 x4 / 16-bit x8).
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> load_pack(const T* p) {
  // streaming read-once data: read-only (non-coherent) path, no L1 allocation.  Measured on the dense decode: the
  // coherent ld.global.cs form cannot overtake the kernel's own stores and reached 53% of the copy peak; .nc reaches 92%.
  if constexpr (sizeof(T) * VEC == 16) {
    Pack<T, VEC> r;
    int4 raw;
#if YPB_LOAD_MODE == 1
    raw = *reinterpret_cast<const int4*>(p);
#elif YPB_LOAD_MODE == 2
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p));
#elif YPB_LOAD_MODE == 3
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p));
#elif YPB_LOAD_MODE == 4
    asm volatile("ld.global.cs.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p));
#else
    raw = __ldcs(reinterpret_cast<const int4*>(p));
#endif
    *reinterpret_cast<int4*>(&r) = raw;
    return r;
  } else if constexpr (sizeof(T) * VEC == 8) {
    Pack<T, VEC> r;
    int2 raw;
#if defined(YPB_LOAD8_L2_256)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.s32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "l"(p));
#elif defined(YPB_LOAD8_L2_128)
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.s32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "l"(p));
#endif
    *reinterpret_cast<int2*>(&r) = raw;
    return r;
  } else if constexpr (sizeof(T) * VEC == 4) {
    Pack<T, VEC> r;
    int raw;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(raw) : "l"(p));
    *reinterpret_cast<int*>(&r) = raw;
    return r;
  } else {
    Pack<T, VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = p[i];
    return r;
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void store_pack(T* p, const Pack<T, VEC>& r) {
  if constexpr (sizeof(T) * VEC == 16) {
    __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&r));
  } else if constexpr (sizeof(T) * VEC == 8) {
#if defined(YPB_STORE8_PLAIN)
    *reinterpret_cast<int2*>(p) = *reinterpret_cast<const int2*>(&r);
#elif defined(YPB_STORE8_WT)
    __stwt(reinterpret_cast<int2*>(p), *reinterpret_cast<const int2*>(&r));
#else
    __stcs(reinterpret_cast<int2*>(p), *reinterpret_cast<const int2*>(&r));
#endif
  } else if constexpr (sizeof(T) * VEC == 4) {
    __stcs(reinterpret_cast<int*>(p), *reinterpret_cast<const int*>(&r));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
  }
}

// Element i of a pack, widened to fp32 with ONE instruction: the compiler turns a 16-bit field read followed by
// __bfloat162float into PRMT + SHL (two issue slots per element in kernels bound by instruction issue); reading the
// containing 32-bit word and shifting / masking it is the same value in one.
template <int DT, int VEC>
__device__ __forceinline__ float pack_elem(const Pack<typename DType<DT>::type, VEC>& p, int i) {
  if constexpr (DT == YPB_BF16 && VEC % 2 == 0) {
    const uint32_t w = reinterpret_cast<const uint32_t*>(&p)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  } else if constexpr (DT == YPB_F16 && VEC % 2 == 0) {
    const __half2 h = reinterpret_cast<const __half2*>(&p)[i >> 1];
    return (i & 1) ? __high2float(h) : __low2float(h);
  } else {
    return DType<DT>::to_f(p.v[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// math of the decode (head.py:151-169, block.py:250-253, tal.py:367-403), fp32 opmath.
// The SAME functions are used by the dense kernel and by the fused filter so the two agree bit for bit.
// ---------------------------------------------------------------------------------------------------------------
// exp / reciprocal as ONE special-function instruction each: `ex2.approx.ftz` / `rcp.approx.ftz`.  The non-ftz forms that
// __expf / __fdividef compile to (no -ftz flag: the exact fp32 arithmetic of the IoU tests must keep denormals) wrap every
// MUFU in a range check and two scalings - 3 extra instructions per exp, 144 exps + 84 reciprocals per anchor in the dense
// decode, which is bound by instruction issue.  The values are the same: e^(v-m) <= 1 is only ever added to a sum >= 1 and
// 1 + e^-x absorbs a denormal, so flushing changes no result bit; 1 / (1 + e^-x) for the overflowing x < -87 is 0 in
// both forms (div.approx returns 0 for 2^126 < y).
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_exp(float x) { return ex2_ftz(__fmul_rn(x, 1.4426950408889634f)); }  // == __expf
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(__fadd_rn(1.0f, fast_exp(-x))); }

// torch max semantics: NaN is sticky.
__device__ __forceinline__ float nanmax(float m, float v) { return (v > m || v != v) ? v : m; }

// Expectation of the softmax over 16 bins (DFL, block.py:250-253).  v[k] = logit of bin k.
// The sums use a FIXED association order of explicitly rounded operations, so every kernel that inlines them (dense decode,
// survivor decode of the fused path, stand-alone DFL) produces the same bits.
__device__ __forceinline__ float tree16(const float (&v)[16]) {
  float a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = __fadd_rn(v[i], v[i + 8]);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = __fadd_rn(a[i], a[i + 4]);
  return __fadd_rn(__fadd_rn(b[0], b[2]), __fadd_rn(b[1], b[3]));
}

// YPB_DFL_FORM 1 (default): the shift by the maximum is folded into the exponent's scaling, t_k = fma(v_k, log2 e, -rn(m log2 e)),
// and the weighted sum runs as four FFMA chains: 91 instead of 134 instructions per (anchor, side) in kernels that are
// bound by instruction issue.  The constant subtracted is the SAME for the 16 bins, so its rounding cancels in the ratio;
// every operation is an explicit _rn intrinsic, so all inlining contexts (dense, fused, stand-alone DFL) round alike.
// YPB_DFL_FORM 0: round 1's form (separate subtract / scale, balanced trees for both sums).
#ifndef YPB_DFL_FORM
#define YPB_DFL_FORM 1
#endif
template <int REG>
__device__ __forceinline__ float dfl_expect(const float (&v)[REG]) {
  static_assert(REG == 16, "only reg_max = 16 is built");
  float m = v[0];
#pragma unroll
  for (int k = 1; k < REG; ++k) m = fmaxf(m, v[k]);
#if YPB_DFL_FORM == 0
  float e[REG], p[REG];
#pragma unroll
  for (int k = 0; k < REG; ++k) {
    e[k] = fast_exp(__fsub_rn(v[k], m));
    p[k] = __fmul_rn(static_cast<float>(k), e[k]);
  }
  return __fmul_rn(tree16(p), rcp_ftz(tree16(e)));  // == __fdividef: x * rcp(y), sum of e in [1, 16]
#else
  const float L2E = 1.4426950408889634f;
  const float nm = __fmul_rn(-m, L2E);
  float e[REG];
#pragma unroll
  for (int k = 0; k < REG; ++k) e[k] = ex2_ftz(__fmaf_rn(v[k], L2E, nm));
  float q[4];
  q[0] = __fmaf_rn(3.f, e[3], __fmaf_rn(2.f, e[2], e[1]));
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    const int k = 4 * j;
    q[j] = __fmaf_rn(static_cast<float>(k + 3), e[k + 3],
                     __fmaf_rn(static_cast<float>(k + 2), e[k + 2],
                               __fmaf_rn(static_cast<float>(k + 1), e[k + 1], __fmul_rn(static_cast<float>(k), e[k]))));
  }
  const float sp = __fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3]));
  return __fmul_rn(sp, rcp_ftz(tree16(e)));  // sum of e in (0.99, 16]
#endif
}

struct BoxXYWH { float cx, cy, w, h; };

// tal.py:367-376 (dist2bbox, xywh) then head.py:168 (x stride).  ax, ay: anchor centre in grid units.
__device__ __forceinline__ BoxXYWH decode_axis_aligned(float dl, float dt, float dr, float db, float ax, float ay,
                                                       float stride, bool xyxy) {
  float x1 = ax - dl, y1 = ay - dt, x2 = ax + dr, y2 = ay + db;
  BoxXYWH o;
  if (xyxy) {
    o.cx = x1 * stride; o.cy = y1 * stride; o.w = x2 * stride; o.h = y2 * stride;
  } else {
    o.cx = (x1 + x2) * 0.5f * stride; o.cy = (y1 + y2) * 0.5f * stride;
    o.w = (x2 - x1) * stride; o.h = (y2 - y1) * stride;
  }
  return o;
}

// tal.py:385-403 (dist2rbox) then head.py:168.
__device__ __forceinline__ BoxXYWH decode_rotated(float dl, float dt, float dr, float db, float theta, float ax,
                                                  float ay, float stride) {
  float s, c;
  sincosf(theta, &s, &c);
  float xf = (dr - dl) * 0.5f, yf = (db - dt) * 0.5f;
  BoxXYWH o;
  // explicit _rn intrinsics: no FMA contraction, so every inlining context rounds identically
  o.cx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(xf, c), __fmul_rn(yf, s)), ax), stride);
  o.cy = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(xf, s), __fmul_rn(yf, c)), ay), stride);
  o.w = (dl + dr) * stride;
  o.h = (dt + db) * stride;
  return o;
}

// head.py:1031: (sigmoid(t) - 0.25) * pi
__device__ __forceinline__ float activate_angle(float t) { return (sigmoid_f(t) - 0.25f) * 3.14159265358979323846f; }

// ---------------------------------------------------------------------------------------------------------------
// result-side rescale (utils/ops.py:102-135 scale_boxes, :152-177 clip_boxes, :621-636 regularize_rboxes): shared by
// scale_rows_kernel and the fused gather of sort_suppress_kernel.  Every step is one separately rounded fp32 operation.
// ---------------------------------------------------------------------------------------------------------------
// torch.remainder on fp32 (ATen cpu/BinaryOpsKernel.cpp remainder_kernel): fmod, then shifted into the sign of b.
__device__ __forceinline__ float torch_remainder(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && ((b < 0.f) != (m < 0.f))) m = __fadd_rn(m, b);
  return m;
}

// torch.clamp_(lo, hi) on fp32: NaN propagates.
__device__ __forceinline__ float torch_clamp(float x, float lo, float hi) {
  return x != x ? x : fminf(fmaxf(x, lo), hi);
}

__device__ __forceinline__ void scale_box(float& x0, float& y0, float& x1, float& y1, float* angle,
                                          const ypb_scale_xform& xf, int mode, bool padding) {
  if (mode == YPB_BOXES_XYWHR || mode == YPB_BOXES_REGULARIZE_ONLY) {
    // ops.py:621-636: swap w/h when (t mod pi) >= pi/2, then t mod pi/2
    const float PI_F = 3.14159274101257324f, HALF_PI_F = 1.57079637050628662f;  // float32(math.pi), float32(math.pi / 2)
    const float th = *angle;
    const bool swap = torch_remainder(th, PI_F) >= HALF_PI_F;
    const float w = swap ? y1 : x1, h = swap ? x1 : y1;
    x1 = w; y1 = h;
    *angle = torch_remainder(th, HALF_PI_F);
  }
  if (mode != YPB_BOXES_CLIP_ONLY && mode != YPB_BOXES_REGULARIZE_ONLY) {
    if (padding) {  // ops.py:128-133
      x0 = __fsub_rn(x0, xf.pad_x); y0 = __fsub_rn(y0, xf.pad_y);
      if (mode == YPB_BOXES_XYXY) { x1 = __fsub_rn(x1, xf.pad_x); y1 = __fsub_rn(y1, xf.pad_y); }
    }
    x0 = __fdiv_rn(x0, xf.gain); y0 = __fdiv_rn(y0, xf.gain);  // ops.py:134
    x1 = __fdiv_rn(x1, xf.gain); y1 = __fdiv_rn(y1, xf.gain);
  }
  if (mode == YPB_BOXES_XYXY || mode == YPB_BOXES_CLIP_ONLY) {  // ops.py:135 -> :163-177
    x0 = torch_clamp(x0, 0.f, xf.img_w); y0 = torch_clamp(y0, 0.f, xf.img_h);
    x1 = torch_clamp(x1, 0.f, xf.img_w); y1 = torch_clamp(y1, 0.f, xf.img_h);
  }
}

// Pose.kpts_decode (head.py:1254-1273), one value: channel e of the (nk*ndim) keypoint block, d = e % ndim.
// x/y: (v*2 + (anchor - 0.5)) * stride, visibility (ndim == 3): sigmoid.  Every torch op rounds to the tensor dtype T.
// ax, ay = the anchor centre as cached in dtype T (head.py:163: arange(w, dtype=T) + 0.5).
template <int DT>
__device__ __forceinline__ float kpt_value(float v, int d, float ax, float ay, float stride) {
  using D = DType<DT>;
  if (d == 2) return D::rnd(sigmoid_f(v));
  if (d > 2) return v;
  const float a = D::rnd(__fsub_rn(d == 0 ? ax : ay, 0.5f));
  float t = D::rnd(__fmul_rn(v, 2.0f));
  t = D::rnd(__fadd_rn(t, a));
  return D::rnd(__fmul_rn(t, stride));
}

// scale_coords (ops.py:562-595) on one coordinate of a keypoint: d = 0 -> x, 1 -> y, else untouched.
__device__ __forceinline__ float scale_coord(float v, int d, const ypb_scale_xform& xf, bool padding) {
  if (d > 1) return v;
  if (padding) v = __fsub_rn(v, d == 0 ? xf.cpad_x : xf.cpad_y);
  v = __fdiv_rn(v, xf.gain);
  return torch_clamp(v, 0.f, d == 0 ? xf.img_w : xf.img_h);
}

// ---------------------------------------------------------------------------------------------------------------
// sort keys: 64-bit, unique per image.  hi = ~orderable(score), lo = row id ((anchor << cls_bits) | cls), so an ascending
// sort is "score descending, then lower row first" == torchvision's stable descending sort (Ties).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t orderable_bits(float f) {
  if (f == 0.0f) f = 0.0f;  // -0 -> +0
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable_bits(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(~orderable_bits(score)) << 32) | row;
}
__device__ __forceinline__ float key_score(uint64_t k) { return from_orderable_bits(~static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return static_cast<uint32_t>(k); }
constexpr uint64_t KEY_SENTINEL = ~0ull;

// ---------------------------------------------------------------------------------------------------------------
// workspace carve-up (all offsets 256-byte aligned)
// ---------------------------------------------------------------------------------------------------------------
struct Workspace {
  int32_t* row_count;  // [B+1]          [0,B): rows emitted by the filter (atomic); [B]: number of tiles to decode
  int32_t* tile_list;  // [tile_cap]     octets (8 consecutive anchors) holding a survivor
  uint8_t* tile_flags; // [B*(A+128)]    per kernel-1 lane (anchor group): which of its VEC anchors survived
  int tile_cap;
  uint64_t* keys_a;    // [B][rows_cap]
  uint64_t* keys_b;    // [B][rows_cap]  radix ping-pong
  float4* cand_box;    // [B][A]         box of each candidate anchor (xyxy, or xywh when rotated), un-offset
  float* cand_ang;     // [B][A]         angle of each candidate anchor (rotated only)
  float4* kept_box;    // [B][max_det]   class-offset boxes of rows already kept (greedy rule)
  float* kept_area;    // [B][max_det]
  uint64_t* kept_key;  // [B][max_det]
  float* rec;          // [B][min(rows_cap,max_nms)][8]  per-rank records (fast rules only)
  size_t bytes;
};

__host__ inline int bits_for(long long n) {  // smallest b with n <= 2^b
  int b = 0;
  while ((1LL << b) < n) ++b;
  return b;
}

__host__ inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

__host__ inline Workspace carve_workspace(void* base, int batch, int anchors, int rows_cap, int max_det, int max_nms,
                                          int rule) {
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align256(bytes);
    return p;
  };
  size_t B = static_cast<size_t>(batch);
  w.row_count = reinterpret_cast<int32_t*>(take((B + 1) * sizeof(int32_t)));
  // kernel 1 runs G <= A/VEC + 127 lanes per image, one flag byte each; an octet is 8 anchors: <= A/8 + 128 per image
  w.tile_cap = static_cast<int>(B * (static_cast<size_t>(anchors) / 8 + 128));
  const size_t flag_bytes = B * (static_cast<size_t>(anchors) + 128);
  w.tile_list = reinterpret_cast<int32_t*>(take(static_cast<size_t>(w.tile_cap) * sizeof(int32_t)));
  w.tile_flags = reinterpret_cast<uint8_t*>(take(flag_bytes));
  w.keys_a = reinterpret_cast<uint64_t*>(take(B * rows_cap * sizeof(uint64_t)));
  w.keys_b = reinterpret_cast<uint64_t*>(take(B * rows_cap * sizeof(uint64_t)));
  w.cand_box = reinterpret_cast<float4*>(take(B * anchors * sizeof(float4)));
  w.cand_ang = reinterpret_cast<float*>(take(B * anchors * sizeof(float)));
  w.kept_box = reinterpret_cast<float4*>(take(B * max_det * sizeof(float4)));
  w.kept_area = reinterpret_cast<float*>(take(B * max_det * sizeof(float)));
  w.kept_key = reinterpret_cast<uint64_t*>(take(B * max_det * sizeof(uint64_t)));
  size_t m = static_cast<size_t>(rows_cap < max_nms ? rows_cap : max_nms);
  w.rec = reinterpret_cast<float*>(take(rule == YPB_NMS_GREEDY ? 0 : B * m * 8 * sizeof(float)));
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// internal launch interface (defined in ypb_decode.cu / ypb_nms.cu, called from ypb_abi.cu)
// ---------------------------------------------------------------------------------------------------------------
struct HeadGeom {  // device-side view of ypb_head_desc, passed by value
  int num_levels, batch, nc, reg_max;
  const void* ptr[YPB_MAX_LEVELS];
  int h[YPB_MAX_LEVELS], w[YPB_MAX_LEVELS];
  long long bstride[YPB_MAX_LEVELS], cstride[YPB_MAX_LEVELS];
  float stride[YPB_MAX_LEVELS];
  int anchor_start[YPB_MAX_LEVELS + 1];  // prefix of H*W
  int group_start[YPB_MAX_LEVELS + 1];   // prefix of H*W/VEC
  int anchors;
};

struct FilterArgs {
  float conf;
  int nc, multi_label, rotated, rows_cap;
  int cls_bits;  // row id = (anchor << cls_bits) | cls
  const uint32_t* class_mask;
  int32_t* row_count;
  int32_t* tile_count;
  int32_t* tile_list;
  uint8_t* tile_flags;
  int tile_cap;
  int boxes_xyxy;   // filter_from_dense: columns 0..3 are corners already (no nms.py:86 conversion)
  const float* conf_per_image;  // filter_from_dense: per-image threshold replacing `conf`, or null
  int fuse_decode;  // 1: the class-scan kernel decodes the boxes of its own survivors; 0: separate decode_tiles kernel
  uint64_t* keys;
  float4* cand_box;
  float* cand_ang;
};

struct SuppressArgs {
  int batch, anchors, nc, extra, max_det, max_nms, rule, rows_cap, multi_label;
  int cls_bits, anchor_bits;  // row id = (anchor << cls_bits) | cls; anchors < 2^anchor_bits
  float iou_thr, max_wh;
  int32_t* row_count;     // per-image rows emitted by the filter; zeroed again by this kernel (clean on exit)
  int32_t* tile_counter;  // octet counter of the fused path (or null); zeroed by the CTA of image 0
  uint64_t* keys_a;
  uint64_t* keys_b;
  const float4* cand_box;
  const float* cand_ang;
  float4* kept_box;
  float* kept_area;
  uint64_t* kept_key;
  float* rec;
  // extras gather source (dense path) - may be null
  const void* pred;
  int pred_dtype;
  long long pred_sb, pred_sc, pred_sa;
  // outputs
  float* out_rows;
  long long* out_idx;
  int32_t* out_count;
  int32_t* out_count_host;  // optional second home of the counts in mapped pinned host memory (ypb_nms_out.count_host)
  int32_t* out_cand;
  int idx_as_row;  // ypb_nms_boxes: write the row id itself
  // riders of the fused path (ypb_riders_desc): per-anchor channels that follow the kept rows as extras
  const void* rider;
  int rider_dtype, rider_kind, rider_ndim;
  long long rider_sb, rider_sc;
  int lv_n;                                 // anchor index -> level / grid cell (keypoint riders)
  int lv_start[YPB_MAX_LEVELS + 1], lv_w[YPB_MAX_LEVELS];
  float lv_stride[YPB_MAX_LEVELS];
  // optional fused construct_result rescale of the kept rows (ypb_nms_out.scale_*)
  const ypb_scale_xform* scale_xforms;
  int scale_padding;
  // exporter NMSModel flavour (ypb_nms_params.nms_box_*): suppression on box_mult * (box / box_div); zero padding
  float box_div, box_mult;
  int pad_zero;
  // one-sided gather over peer memory (ypb_nms_out.peer_*)
  int num_peers, my_rank;
  float* peer_rows[YPB_MAX_PEERS];
  int32_t* peer_count[YPB_MAX_PEERS];
  int32_t* peer_flag[YPB_MAX_PEERS];
  int32_t* peer_state;
  const int32_t* peer_ack;        // local: acknowledgements written by the peers' ypb_peer_wait (or null: no back-pressure)
  int peer_depth;                 // ring entries per rank
  long long peer_entry_stride;    // floats between ring entries
  int peer_debug;                 // diagnostic bit mask (YPB_PEER_DEBUG): 1 local ring only, 2 no fence, 4 no acknowledgement wait
  long long* dbg;  // diagnostic phase timestamps or null
};
void set_phase_buffer(long long* p);

struct ScaleArgs {  // ypb_scale_rows: see include/yolopost_b200.h
  float* rows;
  long long image_stride, row_stride;  // elements
  int batch, rows_per_image;
  const int32_t* count;                // device (B) or null = every row
  const ypb_scale_xform* xforms;       // device (B) or null = `xform` for every image
  ypb_scale_xform xform;
  int box_mode, flags, angle_col;
  float* coords;                       // first point of row 0 of image 0, or null
  long long coord_image_stride, coord_row_stride;
  int nk, ndim;
};
cudaError_t launch_scale_rows(const ScaleArgs& s, cudaStream_t st);

struct KptArgs {  // ypb_kpts_decode
  const void* src;
  void* dst;
  long long sb, sc;
  int batch, channels, ndim, anchors, num_levels;
  int w[YPB_MAX_LEVELS];
  float stride[YPB_MAX_LEVELS];
  int anchor_start[YPB_MAX_LEVELS + 1], group_start[YPB_MAX_LEVELS + 1];
};
cudaError_t launch_kpts_decode(const KptArgs& a, int dtype, int vec, cudaStream_t st);

struct MaskArgs {  // ypb_process_mask
  const void* protos;
  int proto_dtype;
  long long proto_sb, proto_sc;  // elements; rows of mw contiguous pixels
  int C, mh, mw;
  const float* coeffs;
  long long coef_image_stride, coef_row_stride;
  const float* boxes;
  long long box_image_stride, box_row_stride;
  const int32_t* offsets;  // device (B+1) exclusive prefix of the per-image detection counts, or null (one image)
  int batch, total;
  int ih, iw;
  int win_top, win_left, win_h, win_w;
  float scale_h, scale_w, ratio_w, ratio_h;
  int crop_mode;
  uint8_t* out;
  int reg_cap, h_cap;  // set by the launcher: floats reserved for the footprint; rows of the pre-interpolated table (0 = off)
};
cudaError_t launch_process_mask(const MaskArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st);

struct MatchArgs {  // ypb_match_predictions
  const float* preds;      // rows: x1,y1,x2,y2,...,cls at cls_col
  long long pred_image_stride, pred_row_stride;
  int cls_col, batch, rows_per_image;
  const int32_t* count;    // device (B) or null
  const float* labels;     // (sum M, 5) cls,x1,y1,x2,y2 - or null when iou + true_cls are given
  const int32_t* label_offsets;  // device (B+1) or null (single image: m labels)
  int m;
  const float* iou;        // optional precomputed (M, N) matrix (single image)
  long long iou_stride;
  const float* true_cls;   // with iou
  float thr[16];
  int nthr;
  uint8_t* correct;        // (B, rows_per_image, nthr)
  int* win_global;         // optional scratch [nthr * sum M] when the labels do not fit shared memory
};
cudaError_t launch_match_predictions(const MatchArgs& a, int max_labels, cudaStream_t st);

struct Dist2BoxArgs {  // ypb_dist2bbox
  const void* dist;            // (B, 4, A) l,t,r,b
  long long dsb, dsc;
  const void* anchor_points;   // (1|B, 2, A)
  long long asb, asc, asa;
  const void* angle;           // (B, 1, A) or null
  long long angle_sb;
  int batch, anchors, xywh;
  void* out;                   // (B, 4, A)
  long long osb, osc;
};
cudaError_t launch_dfl(const void* x, int dtype, int batch, int anchors, long long sb, long long sc, void* out, long long osb,
                       long long osc, cudaStream_t st);
cudaError_t launch_dist2bbox(const Dist2BoxArgs& d, int dtype, cudaStream_t st);

cudaError_t launch_decode_dense(const HeadGeom& g, int in_dtype, const void* angle, int angle_is_logit,
                                int append_angle, int xyxy, void* out, int out_dtype, long long osb, long long osc,
                                int vec, cudaStream_t st);
// which: 1 = class scan + compaction kernel, 2 = survivor box-decode kernel
cudaError_t launch_filter_from_head(const HeadGeom& g, int in_dtype, int value_dtype, const void* angle,
                                    int angle_is_logit, const FilterArgs& f, int vec, int which, cudaStream_t st);
cudaError_t launch_filter_from_dense(const ypb_dense_desc& d, const FilterArgs& f, cudaStream_t st);
// persistent TMA-fed form of the class scan (ypb_scan_tma.cu); cudaErrorNotSupported = geometry outside its envelope
cudaError_t launch_scan_classes_tma(const HeadGeom& g, int in_dtype, const FilterArgs& f, int vec, cudaStream_t st);
cudaError_t launch_sort_suppress(const SuppressArgs& a, cudaStream_t st);
cudaError_t launch_boxes_prep(const float* boxes, const float* scores, int n, int box_dim, uint64_t* keys,
                              float4* cand_box, float* cand_ang, int32_t* row_count, cudaStream_t st);
cudaError_t launch_compact_results(const float* rows, const long long* idx, const int32_t* count, int batch, int max_det,
                                   int cols, float* out_rows, long long* out_idx, int32_t* out_offsets, cudaStream_t st);
cudaError_t launch_pairwise_iou(const float* a, int n, const float* b, int m, int box_dim, float* out, cudaStream_t st);
cudaError_t launch_peer_wait(const int32_t* flags, int world, int32_t* state, int lag, int depth, int32_t* const* peer_ack_host,
                             int my_rank, long long* slot_index, cudaStream_t st);
cudaError_t launch_peer_wait_copy(const int32_t* flags, int world, int32_t* state, int lag, int depth, int32_t* const* peer_ack_host,
                                  int my_rank, long long* slot_index, const float* ring, long long entry_floats, float* out,
                                  cudaStream_t st);
cudaError_t launch_sigmoid_selftest(int dtype, unsigned long long* violations, cudaStream_t st);

}  // namespace ypb
