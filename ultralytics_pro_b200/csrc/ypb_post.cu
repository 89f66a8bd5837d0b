// Result-side kernels of libyolopost_b200 (sm_100a): what the reference runs on the kept rows right after NMS
// (SURVEY.md section 8f rank 1):
//   scale_rows_kernel    utils/ops.py:102-135 scale_boxes, :152-177 clip_boxes, :621-636 regularize_rboxes,
//                        :562-595 scale_coords, :598-618 clip_coords - applied in place to the (B, max_det, 6+extra) rows
//                        of a whole batch (per-image transform, per-image kept count) or to one box / point set
// These are a few kB of data per batch: the point is ONE launch instead of the reference's 9-13 tiny ATen launches per
// image, with the reference's rounding (every step a separately rounded fp32 operation, true IEEE division).
#include "ypb_common.cuh"

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

// Pose.kpts_decode (head.py:1254-1273) over the whole (B, nk*ndim, A) tensor: thread = VEC consecutive anchors of one
// channel, 128-bit streaming loads/stores; HBM-bound (2 * B * C * A * s bytes).
template <int DT, int VEC>
__global__ void __launch_bounds__(256)
kpts_decode_kernel(const __grid_constant__ KptArgs a) {
  using T = typename DType<DT>::type;
  const int grp = blockIdx.x * blockDim.x + threadIdx.x;
  if (grp >= a.group_start[a.num_levels]) return;
  const int c = blockIdx.y, b = blockIdx.z;
  int l = 0;
#pragma unroll
  for (int i = 1; i < YPB_MAX_LEVELS; ++i)
    if (i < a.num_levels && grp >= a.group_start[i]) l = i;
  const int a_local = (grp - a.group_start[l]) * VEC;
  const int a_glob = a.anchor_start[l] + a_local;
  const T* src = static_cast<const T*>(a.src) + static_cast<long long>(b) * a.sb + static_cast<long long>(c) * a.sc + a_glob;
  T* dst = static_cast<T*>(a.dst) + (static_cast<long long>(b) * a.channels + c) * a.anchors + a_glob;
  const int d = c % a.ndim;
  const int W = a.w[l];
  const float stride = a.stride[l];
  int gy = a_local / W, gx = a_local - gy * W;
  Pack<T, VEC> p = load_pack<T, VEC>(src), q;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float ax = DType<DT>::rnd(static_cast<float>(gx) + 0.5f), ay = DType<DT>::rnd(static_cast<float>(gy) + 0.5f);
    q.v[i] = DType<DT>::from_f(kpt_value<DT>(DType<DT>::to_f(p.v[i]), d, ax, ay, stride));
    if (++gx == W) { gx = 0; ++gy; }
  }
  store_pack<T, VEC>(dst, q);
}

}  // namespace

cudaError_t launch_kpts_decode(const KptArgs& a, int dtype, int vec, cudaStream_t st) {
  const int groups = a.group_start[a.num_levels];
  if (groups <= 0 || a.batch <= 0 || a.channels <= 0) return cudaSuccess;
  dim3 grid((groups + 255) / 256, a.channels, a.batch);
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
