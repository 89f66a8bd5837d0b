// Decode-side kernels of libyolopost_b200 (sm_100a):
//   decode_dense_kernel       Detect._inference drop-in (head.py:151-169, OBB head.py:1026-1042)
//   scan_classes_kernel       fused path 1/2: streaming class scan + confidence filter + row compaction (nms.py:76-131)
//   decode_tiles_kernel       fused path 2/2: DFL box decode of the octets (8 anchors) that hold a survivor (head.py:167-168)
//   filter_from_dense_kernel  confidence filter + compaction of an already decoded tensor (nms.py:76-131)
//
// Layout facts the mapping is built on: every head level is (B, 4*reg_max+nc, H, W) with the H*W anchors contiguous,
// so lanes map to ANCHORS (coalesced, 128-bit per lane) and the 16 DFL bins / nc classes of an anchor are walked by
// the owning thread down the channel stride.  (A lane-per-channel mapping for the sparse survivors was measured and
// rejected: every lane then touches its own 128-byte line and the kernel is bound by L1TEX wavefronts, profiles/.)
#include "ypb_common.cuh"

namespace ypb {

#ifndef YPB_DEC_THREADS
#define YPB_DEC_THREADS 128
#endif
#ifndef YPB_SCAN_DEPTH
#define YPB_SCAN_DEPTH 8
#endif
#ifndef YPB_SCAN_DEPTH16
#define YPB_SCAN_DEPTH16 16
#endif
#ifndef YPB_SCAN_PIPE16
#define YPB_SCAN_PIPE16 8   // loads per software-pipeline batch, 16-bit heads (two batches live in registers)
#endif
#ifndef YPB_SCAN_PIPE32
#define YPB_SCAN_PIPE32 8
#endif
constexpr int DEC_THREADS = YPB_DEC_THREADS;

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_level(const HeadGeom& g, int grp) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < YPB_MAX_LEVELS; ++i)
    if (i < g.num_levels && grp >= g.group_start[i]) l = i;
  return l;
}

__device__ __forceinline__ bool class_allowed(const uint32_t* mask, int c) {
  return mask == nullptr || ((mask[c >> 5] >> (c & 31)) & 1u);
}

// Exclusive scan of one int per thread over a DEC_THREADS block; `total` is returned to every thread.
__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
  __shared__ int warp_sum[DEC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < DEC_THREADS / 32; ++w) {
    int s = warp_sum[w];
    if (w < warp) before += s;
    tot += s;
  }
  total = tot;
  return before + inc - v;
}

// Reserve `total` rows of image b for this block; every thread gets the block base.
__device__ __forceinline__ int reserve_rows(int32_t* row_count, int b, int total) {
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = total > 0 ? atomicAdd(&row_count[b], total) : 0;
  __syncthreads();
  return base_s;
}

// xywh (T-rounded) -> corners in T arithmetic, exactly ops.py:281-283 evaluated on a T tensor.
template <int DT>
__device__ __forceinline__ float4 corners_in_dtype(float cx, float cy, float w, float h) {
  using D = DType<DT>;
  float hw = D::rnd(w * 0.5f), hh = D::rnd(h * 0.5f);
  return make_float4(D::rnd(__fsub_rn(cx, hw)), D::rnd(__fsub_rn(cy, hh)), D::rnd(__fadd_rn(cx, hw)),
                     D::rnd(__fadd_rn(cy, hh)));
}

// Walk `n` channel rows of one anchor group with DEPTH independent 128-bit loads in flight per thread: the loads of a
// batch are issued back to back before the first value is consumed (Little's law: ~30 warps/SM x 32 lanes x DEPTH x 16 B).
template <typename TI, int VEC, int DEPTH = YPB_SCAN_DEPTH, typename F>
__device__ __forceinline__ void stream_rows(const TI* base, long long cs, int n, F&& visit) {
  int c = 0;
  for (; c + DEPTH <= n; c += DEPTH) {
    Pack<TI, VEC> p[DEPTH];
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) p[j] = load_pack<TI, VEC>(base + static_cast<long long>(c + j) * cs);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) visit(p[j], c + j);
  }
  for (; c < n; ++c) {
    Pack<TI, VEC> p = load_pack<TI, VEC>(base + static_cast<long long>(c) * cs);
    visit(p, c);
  }
}

// The same walk, software-pipelined: the loads of batch k+1 are issued BEFORE batch k is consumed, so DEPTH independent
// loads per thread are in flight during the arithmetic as well, not only between batches.  (Measured with the read-only
// probe of this access shape, tools/read_bw_bench.cu: 16-bit rows 16.3 us pipelined vs 18.4-18.7 us batch-by-batch.)
template <typename TI, int VEC, int DEPTH, typename F>
__device__ __forceinline__ void stream_rows_pipelined(const TI* base, long long cs, int n, F&& visit) {
  if (n < 2 * DEPTH) {
    stream_rows<TI, VEC, DEPTH>(base, cs, n, visit);
    return;
  }
  Pack<TI, VEC> cur[DEPTH], nxt[DEPTH];
#pragma unroll
  for (int j = 0; j < DEPTH; ++j) cur[j] = load_pack<TI, VEC>(base + static_cast<long long>(j) * cs);
  int c = DEPTH;
  for (; c + DEPTH <= n; c += DEPTH) {
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) nxt[j] = load_pack<TI, VEC>(base + static_cast<long long>(c + j) * cs);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) visit(cur[j], c - DEPTH + j);
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) cur[j] = nxt[j];
  }
#pragma unroll
  for (int j = 0; j < DEPTH; ++j) visit(cur[j], c - DEPTH + j);
  for (; c < n; ++c) {
    Pack<TI, VEC> p = load_pack<TI, VEC>(base + static_cast<long long>(c) * cs);
    visit(p, c);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dense decode
// ---------------------------------------------------------------------------------------------------------------
enum { MODE_XYWH = 0, MODE_XYXY = 1, MODE_ROT = 2 };

template <int DT_IN, int VEC, int REG>
__device__ __forceinline__ void load_ltrb(const typename DType<DT_IN>::type* src, long long cs, float (&d)[4][VEC]) {
  using TI = typename DType<DT_IN>::type;
#pragma unroll
  for (int side = 0; side < 4; ++side) {
    Pack<TI, VEC> raw[REG];
#pragma unroll
    for (int k = 0; k < REG; ++k) raw[k] = load_pack<TI, VEC>(src + static_cast<long long>(side * REG + k) * cs);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float v[REG];
#pragma unroll
      for (int k = 0; k < REG; ++k) v[k] = DType<DT_IN>::to_f(raw[k].v[i]);
      d[side][i] = dfl_expect<REG>(v);
    }
  }
}

template <int DT_IN, int DT_OUT, int VEC, int REG, int MODE>
__global__ void __launch_bounds__(DEC_THREADS)
decode_dense_kernel(const __grid_constant__ HeadGeom g, const void* __restrict__ angle_v, int angle_is_logit,
                    int append_angle, void* __restrict__ out_v, long long osb, long long osc) {
  using TI = typename DType<DT_IN>::type;
  using TO = typename DType<DT_OUT>::type;
  const int grp = blockIdx.x * DEC_THREADS + threadIdx.x;
  const int b = blockIdx.y;
  if (grp >= g.group_start[g.num_levels]) return;
  const int l = find_level(g, grp);
  const int a_local = (grp - g.group_start[l]) * VEC;
  const int a_glob = g.anchor_start[l] + a_local;
  const long long cs = g.cstride[l];
  const TI* src = static_cast<const TI*>(g.ptr[l]) + static_cast<long long>(b) * g.bstride[l] + a_local;
  TO* out = static_cast<TO*>(out_v) + static_cast<long long>(b) * osb + a_glob;
  const int nc = g.nc;

  // ---- boxes: DFL expectation per side, then dist2bbox / dist2rbox, x stride -------------------------------------
  float d[4][VEC];
  load_ltrb<DT_IN, VEC, REG>(src, cs, d);

  float theta[VEC];
  if constexpr (MODE == MODE_ROT) {
    const TI* ang = static_cast<const TI*>(angle_v) + static_cast<long long>(b) * g.anchors + a_glob;
    Pack<TI, VEC> pa = load_pack<TI, VEC>(ang);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float t = DType<DT_IN>::to_f(pa.v[i]);
      theta[i] = angle_is_logit ? DType<DT_OUT>::rnd(activate_angle(t)) : t;
    }
  }

  const int W = g.w[l];
  const float stride = g.stride[l];
  int gy = a_local / W, gx = a_local - gy * W;
  Pack<TO, VEC> o0, o1, o2, o3;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float ax = static_cast<float>(gx) + 0.5f, ay = static_cast<float>(gy) + 0.5f;
    BoxXYWH bx;
    if constexpr (MODE == MODE_ROT)
      bx = decode_rotated(d[0][i], d[1][i], d[2][i], d[3][i], theta[i], ax, ay, stride);
    else
      bx = decode_axis_aligned(d[0][i], d[1][i], d[2][i], d[3][i], ax, ay, stride, MODE == MODE_XYXY);
    o0.v[i] = DType<DT_OUT>::from_f(bx.cx);
    o1.v[i] = DType<DT_OUT>::from_f(bx.cy);
    o2.v[i] = DType<DT_OUT>::from_f(bx.w);
    o3.v[i] = DType<DT_OUT>::from_f(bx.h);
    if (++gx == W) { gx = 0; ++gy; }
  }
  store_pack<TO, VEC>(out, o0);
  store_pack<TO, VEC>(out + osc, o1);
  store_pack<TO, VEC>(out + 2 * osc, o2);
  store_pack<TO, VEC>(out + 3 * osc, o3);

  // ---- class scores: sigmoid, streamed ---------------------------------------------------------------------------
  const TI* csrc = src + static_cast<long long>(4 * REG) * cs;
  TO* cdst = out + 4 * osc;
  {
    // keep the loads of the next 8 rows in flight while the current 8 are activated and stored (software pipeline), instead
    // of load-batch / compute-batch phases: this loop carries no per-anchor state, so the second register buffer is free
    auto act = [&](const Pack<TI, VEC>& p, int c) {
      Pack<TO, VEC> q;
#pragma unroll
      for (int i = 0; i < VEC; ++i) q.v[i] = DType<DT_OUT>::from_f(sigmoid_f(DType<DT_IN>::to_f(p.v[i])));
      store_pack<TO, VEC>(cdst + static_cast<long long>(c) * osc, q);
    };
    stream_rows_pipelined<TI, VEC, 8>(csrc, cs, nc, act);
  }
  if constexpr (MODE == MODE_ROT) {
    if (append_angle) {
      Pack<TO, VEC> q;
#pragma unroll
      for (int i = 0; i < VEC; ++i) q.v[i] = DType<DT_OUT>::from_f(theta[i]);
      store_pack<TO, VEC>(out + static_cast<long long>(4 + nc) * osc, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dense decode of 16-bit heads as ONE uniform row pipeline (round 2; fp32 heads keep decode_dense_kernel, 0.95 of the copy peak)
//
//   thread = 4 anchors (64-bit accesses); the 64 + nc channel rows of its anchors are walked in batches of 16 rows: a batch
//   is exactly one DFL side (16 bins) or 16 class rows.  Two register buffers alternate: the 16 loads of batch k+1 are
//   issued before batch k is consumed, so every thread keeps 16 independent 8-byte loads in flight through the arithmetic
//   of BOTH phases (16 warps x 32 lanes x 16 x 8 B = 64 KB per SM; round 1's form had 8 in flight in the class phase and
//   none across the four sides).  The batch loop is rolled (two bodies, one per buffer) instead of four unrolled sides plus an
//   unrolled class loop: 1/3 of the code, which matters because the old kernel lost 0.84 warp-stalls per issue to
//   instruction fetch (ncu, profiles/r02_dense_bf16_raw.csv).  Element unpacking is one instruction (pack_elem).
// ---------------------------------------------------------------------------------------------------------------
#ifndef YPB_DD16_MIN_BLOCKS
#define YPB_DD16_MIN_BLOCKS 4
#endif
template <int DT, int MODE>
__global__ void __launch_bounds__(DEC_THREADS, YPB_DD16_MIN_BLOCKS)
decode_dense16_kernel(const __grid_constant__ HeadGeom g, const void* __restrict__ angle_v, int angle_is_logit,
                      int append_angle, void* __restrict__ out_v, long long osb, long long osc) {
  using T = typename DType<DT>::type;
  using D = DType<DT>;
  constexpr int VEC = 4, R = 16;
  using P = Pack<T, VEC>;
  const int grp = blockIdx.x * DEC_THREADS + threadIdx.x;
  const int b = blockIdx.y;
  if (grp >= g.group_start[g.num_levels]) return;
  const int l = find_level(g, grp);
  const int a_local = (grp - g.group_start[l]) * VEC;
  const int a_glob = g.anchor_start[l] + a_local;
  const long long cs = g.cstride[l];
  const T* src = static_cast<const T*>(g.ptr[l]) + static_cast<long long>(b) * g.bstride[l] + a_local;
  T* out = static_cast<T*>(out_v) + static_cast<long long>(b) * osb + a_glob;
  const int nc = g.nc;
  const int nb = 4 + nc / R;  // whole batches: 4 sides, then the class rows 16 at a time

  float dl[VEC], dt[VEC], dr[VEC];  // sides 0..2; side 3 is consumed where it is produced

  auto load_batch = [&](P (&buf)[R], int k) {
    const T* p = src + static_cast<long long>(k * R) * cs;
#pragma unroll
    for (int j = 0; j < R; ++j) { buf[j] = load_pack<T, VEC>(p); p += cs; }
  };
  auto process = [&](const P (&buf)[R], int k) {
    if (k < 4) {
      float d[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float v[R];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) v[kk] = pack_elem<DT, VEC>(buf[kk], i);
        d[i] = dfl_expect<R>(v);
      }
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) dl[i] = d[i];
      } else if (k == 1) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) dt[i] = d[i];
      } else if (k == 2) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) dr[i] = d[i];
      } else {
        // ---- all four sides known: dist2bbox / dist2rbox, x stride (tal.py:367-403, head.py:168) ----
        float theta[VEC];
        if constexpr (MODE == MODE_ROT) {
          const T* ang = static_cast<const T*>(angle_v) + static_cast<long long>(b) * g.anchors + a_glob;
          const P pa = load_pack<T, VEC>(ang);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float t = pack_elem<DT, VEC>(pa, i);
            theta[i] = angle_is_logit ? D::rnd(activate_angle(t)) : t;
          }
        }
        const int W = g.w[l];
        const float stride = g.stride[l];
        int gy = a_local / W, gx = a_local - gy * W;
        P o0, o1, o2, o3;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float ax = static_cast<float>(gx) + 0.5f, ay = static_cast<float>(gy) + 0.5f;
          BoxXYWH bx;
          if constexpr (MODE == MODE_ROT)
            bx = decode_rotated(dl[i], dt[i], dr[i], d[i], theta[i], ax, ay, stride);
          else
            bx = decode_axis_aligned(dl[i], dt[i], dr[i], d[i], ax, ay, stride, MODE == MODE_XYXY);
          o0.v[i] = D::from_f(bx.cx);
          o1.v[i] = D::from_f(bx.cy);
          o2.v[i] = D::from_f(bx.w);
          o3.v[i] = D::from_f(bx.h);
          if (++gx == W) { gx = 0; ++gy; }
        }
        store_pack<T, VEC>(out, o0);
        store_pack<T, VEC>(out + osc, o1);
        store_pack<T, VEC>(out + 2 * osc, o2);
        store_pack<T, VEC>(out + 3 * osc, o3);
        if constexpr (MODE == MODE_ROT) {
          if (append_angle) {
            P q;
#pragma unroll
            for (int i = 0; i < VEC; ++i) q.v[i] = D::from_f(theta[i]);
            store_pack<T, VEC>(out + static_cast<long long>(4 + nc) * osc, q);
          }
        }
      }
    } else {
      // ---- 16 class rows: sigmoid (head.py:169), stored where the cat of head.py:169 puts them ----
      T* q_row = out + static_cast<long long>(4 + (k - 4) * R) * osc;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        P q;
#pragma unroll
        for (int i = 0; i < VEC; ++i) q.v[i] = D::from_f(sigmoid_f(pack_elem<DT, VEC>(buf[j], i)));
        store_pack<T, VEC>(q_row, q);
        q_row += osc;
      }
    }
  };

  P bufa[R], bufb[R];
#ifdef YPB_DD16_STAGGER
  // odd warps walk the class batches first and the four sides last: the two phases differ in their special-function density
  // (15 % vs 24 % of the instructions), and warps that start together would otherwise load the XU pipe in lockstep
  const int shift = ((threadIdx.x >> 5) & 1) ? 4 : 0;
  auto batch_of = [&](int k) { int kk = k + shift; return kk >= nb ? kk - nb : kk; };
#else
  auto batch_of = [&](int k) { return k; };
#endif
  load_batch(bufa, batch_of(0));
#pragma unroll 1
  for (int k = 0; k < nb; k += 2) {
    if (k + 1 < nb) load_batch(bufb, batch_of(k + 1));
    process(bufa, batch_of(k));
    if (k + 1 >= nb) break;
    if (k + 2 < nb) load_batch(bufa, batch_of(k + 2));
    process(bufb, batch_of(k + 1));
  }
  // class rows beyond the last whole batch (nc % 16)
  const int c0 = (nc / R) * R;
  if (c0 < nc) {
    const T* csrc = src + static_cast<long long>(4 * R + c0) * cs;
    T* cdst = out + static_cast<long long>(4 + c0) * osc;
    auto act = [&](const P& p, int c) {
      P q;
#pragma unroll
      for (int i = 0; i < VEC; ++i) q.v[i] = D::from_f(sigmoid_f(pack_elem<DT, VEC>(p, i)));
      store_pack<T, VEC>(cdst + static_cast<long long>(c) * osc, q);
    };
    stream_rows<T, VEC, 8>(csrc, cs, nc - c0, act);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused path, kernel 1: class scan + confidence filter + compaction (nms.py:76-131 evaluated on head.py:169's scores
// without ever materialising them)
//
//   thread = VEC consecutive anchors; the nc class rows are streamed once with 128-bit loads, DEPTH loads in flight.
//   single-label: per anchor the max logit m (NaN-propagating), its first index and the runner-up m2: 5 ALU ops per
//     element, no transcendental.  sigmoid_f is monotone after rounding (ypb_selftest_sigmoid_monotone), so
//     max_c score == round_T(sigmoid(m)) - ONE sigmoid per anchor - and unless the runner-up rounds to the same score
//     the first argmax of the scores (nms.py:120) is the first argmax of the logits; the rare tie re-reads the anchor.
//   multi-label: every (anchor, class) with score > conf is a row (nms.py:115): counted in the streaming pass, the
//     survivors' classes are re-read to write the keys.
//   Output: unique 64-bit sort keys (row order irrelevant: one atomicAdd per block reserves the slots) and the list of
//   octets (8 anchors) that contain a survivor, with a per-lane flag byte, for kernel 2.
// ---------------------------------------------------------------------------------------------------------------
template <int DT_IN, int DT_VAL, bool ROT>
__device__ __forceinline__ void decode_survivor(const HeadGeom& g, const FilterArgs& f, const void* angle_v,
                                                int angle_is_logit, int b, int l, int a_local);

__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

template <int DT_IN, int DT_VAL, int VEC, bool MULTI, bool ROT>
__global__ void __launch_bounds__(DEC_THREADS)
scan_classes_kernel(const __grid_constant__ HeadGeom g, const void* __restrict__ angle_v, int angle_is_logit,
                    const __grid_constant__ FilterArgs f) {
  using TI = typename DType<DT_IN>::type;
  using DV = DType<DT_VAL>;
  __shared__ int s_base[2];
  __shared__ int s_active[DEC_THREADS / 32];
  __shared__ uint8_t s_flags[DEC_THREADS];             // fused decode: survivor flags of every lane of the CTA
  __shared__ uint8_t s_oct[DEC_THREADS * VEC / 8 + 1]; // fused decode: octets of the CTA that hold a survivor
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = blockIdx.x * DEC_THREADS + tid;
  const int b = blockIdx.y;
  const int nc = g.nc;
  const float conf = f.conf;

  int rows[VEC];
  float score[VEC];
  int cls[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { rows[i] = 0; score[i] = 0.f; cls[i] = 0; }
  int a_glob = 0;
  long long cs = 0;
  const TI* csrc = nullptr;

  if (grp < g.group_start[g.num_levels]) {
    const int l = find_level(g, grp);
    const int a_local = (grp - g.group_start[l]) * VEC;
    a_glob = g.anchor_start[l] + a_local;
    cs = g.cstride[l];
    csrc = static_cast<const TI*>(g.ptr[l]) + static_cast<long long>(b) * g.bstride[l] + a_local + 64 * cs;
    if constexpr (MULTI) {
      float nanacc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) nanacc[i] = -INFINITY;
      auto visit = [&](const Pack<TI, VEC>& p, int c) {
        const bool ok = class_allowed(f.class_mask, c);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float v = DType<DT_IN>::to_f(p.v[i]);
          nanacc[i] = max_nan(nanacc[i], v);
          rows[i] += (DV::rnd(sigmoid_f(v)) > conf && ok) ? 1 : 0;
        }
      };
      stream_rows<TI, VEC>(csrc, cs, nc, visit);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (nanacc[i] != nanacc[i]) rows[i] = 0;  // amax -> NaN -> the anchor is not a candidate (nms.py:76)
    } else {
      float m[VEC], m2[VEC];
      if constexpr (DT_IN != YPB_F32 && VEC == 8) {
        // 16-bit inputs: the whole scan stays in packed x2 arithmetic (comparisons and max/min of 16-bit floats are
        // exact): per PAIR of elements one compare-mask, min, max, NaN-propagating max and a bit-select of the packed
        // 16-bit class indices - 2.5 instructions per element instead of ~7 through fp32.
        using T2 = typename Packed2<DT_IN>::type;
        T2 pm[4], pm2[4];
        uint32_t pidx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { pm[j] = Packed2<DT_IN>::neg_inf(); pm2[j] = pm[j]; pidx[j] = 0u; }
        auto visit = [&](const Pack<TI, VEC>& p, int c) {
          const T2* v2 = reinterpret_cast<const T2*>(&p);
          const uint32_t cc = static_cast<uint32_t>(c) * 0x00010001u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t gt = __hgt2_mask(v2[j], pm[j]);
            pm2[j] = __hmax2(pm2[j], __hmin2(pm[j], v2[j]));
            pm[j] = __hmax2_nan(pm[j], v2[j]);
            pidx[j] = (cc & gt) | (pidx[j] & ~gt);
          }
        };
        stream_rows_pipelined<TI, VEC, YPB_SCAN_PIPE16>(csrc, cs, nc, visit);  // 16-bit rows: loads of the next batch in flight during the math
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m[2 * j] = Packed2<DT_IN>::lo(pm[j]);   m[2 * j + 1] = Packed2<DT_IN>::hi(pm[j]);
          m2[2 * j] = Packed2<DT_IN>::lo(pm2[j]); m2[2 * j + 1] = Packed2<DT_IN>::hi(pm2[j]);
          cls[2 * j] = static_cast<int>(pidx[j] & 0xffffu); cls[2 * j + 1] = static_cast<int>(pidx[j] >> 16);
        }
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { m[i] = -INFINITY; m2[i] = -INFINITY; }
        auto visit = [&](const Pack<TI, VEC>& p, int c) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float v = DType<DT_IN>::to_f(p.v[i]);
            const bool gt = v > m[i];
            m2[i] = fmaxf(m2[i], fminf(m[i], v));
            m[i] = max_nan(m[i], v);
            cls[i] = gt ? c : cls[i];
          }
        };
        // fp32 rows: batch-by-batch (the pipelined form costs 36 more registers, i.e. 5 instead of 9 resident CTAs per SM, and
        // measured 2 % slower: 32.4 vs 31.8 us for C2)
        stream_rows<TI, VEC>(csrc, cs, nc, visit);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float s = DV::rnd(sigmoid_f(m[i]));
        if (s > conf) {  // false for a NaN max
          if (DV::rnd(sigmoid_f(m2[i])) == s) {
            // the runner-up rounds to the same score: take the FIRST class that reaches it (nms.py:120)
            for (int c = 0; c < cls[i]; ++c) {
              if (DV::rnd(sigmoid_f(DType<DT_IN>::to_f(csrc[static_cast<long long>(c) * cs + i]))) == s) { cls[i] = c; break; }
            }
          }
          score[i] = s;
          rows[i] = class_allowed(f.class_mask, cls[i]) ? 1 : 0;  // nms.py:127-131
        }
      }
    }
  }

  int my_rows = 0;
  uint32_t flags = 0;
#pragma unroll
  for (int i = 0; i < VEC; ++i) { my_rows += rows[i]; flags |= rows[i] > 0 ? 1u << i : 0u; }
  // octets (8 consecutive anchors = one 32-byte fp32 sector per row) of this warp's span that hold a survivor:
  // kernel 2's work list.  LPO lanes of this kernel cover one octet.
  constexpr int LPO = VEC >= 8 ? 1 : 8 / VEC;
  constexpr int NOCT = 32 / LPO;
  const unsigned bal = __ballot_sync(0xffffffffu, flags != 0);
  uint32_t oct_mask = 0;
#pragma unroll
  for (int o = 0; o < NOCT; ++o)
    if ((bal >> (o * LPO)) & ((1u << LPO) - 1u)) oct_mask |= 1u << o;
  if (lane == 0) s_active[warp] = __popc(oct_mask);
  int total_rows;
  int roff = block_exclusive_scan(my_rows, total_rows);  // contains a __syncthreads
  if (total_rows == 0) return;  // uniform
  if (tid == 0) {
    int act = 0;
#pragma unroll
    for (int w = 0; w < DEC_THREADS / 32; ++w) act += s_active[w];
    s_base[0] = atomicAdd(&f.row_count[b], total_rows);
    s_base[1] = f.fuse_decode ? 0 : atomicAdd(f.tile_count, act);
  }
  __syncthreads();
  int oct_rank = 0;
  for (int w = 0; w < warp; ++w) oct_rank += s_active[w];
  int oct_total = oct_rank;
  for (int w = warp; w < DEC_THREADS / 32; ++w) oct_total += s_active[w];
  if (f.fuse_decode) {
    s_flags[tid] = static_cast<uint8_t>(flags);
    if (lane < NOCT && ((oct_mask >> lane) & 1u))
      s_oct[oct_rank + __popc(oct_mask & ((1u << lane) - 1u))] = static_cast<uint8_t>(warp * NOCT + lane);
  } else if (oct_mask) {  // warp-uniform
    // octet ids index a dense per-lane space: index = b * G + lane-in-image, G = lanes kernel 1 runs per image
    const int G = static_cast<int>(gridDim.x) * DEC_THREADS;
    const int lane_idx = b * G + blockIdx.x * DEC_THREADS + tid;
    if (lane < NOCT && ((oct_mask >> lane) & 1u))
      f.tile_list[s_base[1] + oct_rank + __popc(oct_mask & ((1u << lane) - 1u))] = (lane_idx - lane) / LPO + lane;
    f.tile_flags[lane_idx] = static_cast<uint8_t>(flags);
  }
  if (my_rows > 0) {
    uint64_t* keys = f.keys + static_cast<long long>(b) * f.rows_cap;
    int rpos = s_base[0] + roff;
    if constexpr (MULTI) {
      // second streaming pass over the same class rows (now L2 hits), with the same coalesced 128-bit loads: anchor i of
      // this thread owns the slots [cur[i], cur[i] + rows[i]) and fills them class by class.  In val mode (conf 0.001) nearly
      // every anchor has a row, so a per-anchor scalar re-read would serialise 4 x nc uncoalesced loads per thread.
      int cur[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) { cur[i] = rpos; rpos += rows[i]; }
      auto emit = [&](const Pack<TI, VEC>& p, int c) {
        if (!class_allowed(f.class_mask, c)) return;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float s = DV::rnd(sigmoid_f(DType<DT_IN>::to_f(p.v[i])));
          if (rows[i] > 0 && s > conf) {  // rows[i] == 0: NaN anchor (nms.py:76) or nothing above conf
            if (cur[i] < f.rows_cap) keys[cur[i]] = make_key(s, (static_cast<uint32_t>(a_glob + i) << f.cls_bits) + c);
            ++cur[i];
          }
        }
      };
      stream_rows<TI, VEC>(csrc, cs, nc, emit);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (rows[i] == 0) continue;
        const uint32_t row0 = static_cast<uint32_t>(a_glob + i) << f.cls_bits;
        if (rpos < f.rows_cap) keys[rpos] = make_key(score[i], row0 + cls[i]);
        ++rpos;
      }
    }
  }
  if (!f.fuse_decode) return;

  // ---- fused survivor decode (head.py:167-168 restricted to survivors): quarter-warp = one octet, lane = one anchor ----
  __syncthreads();
  for (int t = warp * 4 + (lane >> 3); t < oct_total; t += (DEC_THREADS / 32) * 4) {
    const int oct = s_oct[t];
    const int k = lane & 7;
    const int src_lane = oct * LPO + (VEC >= 8 ? 0 : k / VEC);  // lane of this CTA that scanned the anchor
    const int i = VEC >= 8 ? k : k % VEC;
    if (!((s_flags[src_lane] >> i) & 1u)) continue;
    const int grp2 = blockIdx.x * DEC_THREADS + src_lane;
    const int l2 = find_level(g, grp2);
    const int a_loc = (grp2 - g.group_start[l2]) * VEC + i;
    decode_survivor<DT_IN, DT_VAL, ROT>(g, f, angle_v, angle_is_logit, b, l2, a_loc);
  }
}

// One survivor anchor: 64 independent bin-row loads, in-register softmax expectation per side (block.py:250-253),
// dist2bbox / dist2rbox, x stride, rounding through the value dtype, corners (nms.py:86).
template <int DT_IN, int DT_VAL, bool ROT>
__device__ __forceinline__ void decode_survivor(const HeadGeom& g, const FilterArgs& f, const void* angle_v,
                                                int angle_is_logit, int b, int l, int a_local) {
  using TI = typename DType<DT_IN>::type;
  using DV = DType<DT_VAL>;
  const long long cs = g.cstride[l];
  const TI* src = static_cast<const TI*>(g.ptr[l]) + static_cast<long long>(b) * g.bstride[l] + a_local;
  typename DType<DT_IN>::type raw[64];
#pragma unroll
  for (int kk = 0; kk < 64; ++kk) raw[kk] = src[static_cast<long long>(kk) * cs];  // 64 independent loads in flight
  float d[4];
#pragma unroll
  for (int side = 0; side < 4; ++side) {
    float v[16];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) v[kk] = DType<DT_IN>::to_f(raw[side * 16 + kk]);
    d[side] = dfl_expect<16>(v);
  }
  const int W = g.w[l];
  const int gy = a_local / W, gx = a_local - gy * W;
  const float ax = static_cast<float>(gx) + 0.5f, ay = static_cast<float>(gy) + 0.5f;
  const float stride = g.stride[l];
  const long long slot = static_cast<long long>(b) * g.anchors + g.anchor_start[l] + a_local;
  if constexpr (ROT) {
    float tt = DType<DT_IN>::to_f(static_cast<const TI*>(angle_v)[slot]);
    float theta = angle_is_logit ? DV::rnd(activate_angle(tt)) : tt;
    BoxXYWH bx = decode_rotated(d[0], d[1], d[2], d[3], theta, ax, ay, stride);
    f.cand_box[slot] = make_float4(DV::rnd(bx.cx), DV::rnd(bx.cy), DV::rnd(bx.w), DV::rnd(bx.h));
    f.cand_ang[slot] = theta;
  } else {
    BoxXYWH bx = decode_axis_aligned(d[0], d[1], d[2], d[3], ax, ay, stride, false);
    f.cand_box[slot] = corners_in_dtype<DT_VAL>(DV::rnd(bx.cx), DV::rnd(bx.cy), DV::rnd(bx.w), DV::rnd(bx.h));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused path, kernel 2: box decode of the octets that hold a survivor (head.py:167-168 restricted to them)
//
//   quarter-warp = one octet (8 consecutive anchors) per iteration, lane = one anchor: every bin row costs the warp four
//   32-byte sectors, the 64 rows are independent loads, and the 4 x 16-bin softmax expectation (block.py:250-253) is an
//   in-register reduction that flagged lanes alone execute.  Then dist2bbox / dist2rbox, x stride, rounding through
//   the value dtype and the corner conversion (nms.py:86).  Grid-stride over the octet list kernel 1 built, so the
//   work is spread over the whole GPU whatever the per-image candidate counts are.
// ---------------------------------------------------------------------------------------------------------------
template <int DT_IN, int DT_VAL, int VEC, bool ROT>
__global__ void __launch_bounds__(DEC_THREADS)
decode_tiles_kernel(const __grid_constant__ HeadGeom g, const void* __restrict__ angle_v, int angle_is_logit,
                    const __grid_constant__ FilterArgs f, int G) {
  using TI = typename DType<DT_IN>::type;
  using DV = DType<DT_VAL>;
  constexpr int LPO = VEC >= 8 ? 1 : 8 / VEC;  // kernel-1 lanes per octet
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (DEC_THREADS / 32);
  const int noct = min(*f.tile_count, f.tile_cap);
  for (int t = (blockIdx.x * (DEC_THREADS / 32) + (threadIdx.x >> 5)) * 4 + (lane >> 3); t < noct; t += nwarps * 4) {
    const int oct = f.tile_list[t];
    const int k = lane & 7;                      // anchor inside the octet
    const int lane_idx = oct * LPO + (VEC >= 8 ? 0 : k / VEC);  // kernel-1 lane (anchor group) holding this anchor
    const int i = VEC >= 8 ? k : k % VEC;        // anchor inside the group
    const uint32_t flags = f.tile_flags[lane_idx];
    if (!((flags >> i) & 1u)) continue;          // flagged anchors are always inside the image
    const int b = lane_idx / G;
    const int grp = lane_idx - b * G;
    const int l = find_level(g, grp);
    decode_survivor<DT_IN, DT_VAL, ROT>(g, f, angle_v, angle_is_logit, b, l, (grp - g.group_start[l]) * VEC + i);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// filter + compaction of an already decoded (B, 4+nc+extra, A) tensor with arbitrary strides (nms.py:76-131)
// ---------------------------------------------------------------------------------------------------------------
template <int DT, bool ROT, bool MULTI>
__global__ void __launch_bounds__(DEC_THREADS)
filter_from_dense_kernel(const __grid_constant__ ypb_dense_desc d, const __grid_constant__ FilterArgs f) {
  using T = typename DType<DT>::type;
  using D = DType<DT>;
  const int slot_i = blockIdx.x * DEC_THREADS + threadIdx.x;
  const int b = blockIdx.y;
  const int nc = f.nc;
  const float conf = f.conf_per_image ? f.conf_per_image[b] : f.conf;
  // one thread per anchor - or per entry of the anchor subset (ypb_dense_desc.anchor_subset): the anchor is then read through
  // the list; an index outside [0, anchors) (a padding entry) is no candidate
  int a = slot_i;
  bool active = slot_i < d.anchors;
  if (d.anchor_subset) {
    active = slot_i < d.subset_len;
    if (active) {
      const long long v = d.anchor_subset[static_cast<long long>(b) * d.subset_len + slot_i];
      active = v >= 0 && v < d.anchors;
      a = active ? static_cast<int>(v) : 0;
    }
  }
  const T* p = static_cast<const T*>(d.ptr) + static_cast<long long>(b) * d.stride_b + static_cast<long long>(a) * d.stride_a;
  const T* pc = p + 4 * d.stride_c;
  const long long sc = d.stride_c;

  int my_rows = 0, bcls = 0;
  float best = 0.f;
  if (active) {
    if constexpr (MULTI) {
      bool has_nan = false;
#pragma unroll 8
      for (int c = 0; c < nc; ++c) {
        float s = D::to_f(pc[c * sc]);
        has_nan |= (s != s);
        my_rows += (s > conf && class_allowed(f.class_mask, c)) ? 1 : 0;
      }
      if (has_nan) my_rows = 0;
    } else {
      // torch max: NaN wins; otherwise first index of the maximum (nms.py:120)
      float bs = -INFINITY;
      int bc = 0;
      bool has_nan = false;
#pragma unroll 8
      for (int c = 0; c < nc; ++c) {
        float s = D::to_f(pc[c * sc]);
        has_nan |= (s != s);
        if (s > bs) { bs = s; bc = c; }
      }
      best = bs;
      bcls = bc;
      my_rows = (!has_nan && bs > conf && class_allowed(f.class_mask, bc)) ? 1 : 0;
    }
    if (my_rows > 0) {
      const long long slot = static_cast<long long>(b) * d.anchors + a;
      float cx = D::to_f(p[0]), cy = D::to_f(p[sc]), w = D::to_f(p[2 * sc]), h = D::to_f(p[3 * sc]);
      if constexpr (ROT) {
        f.cand_box[slot] = make_float4(cx, cy, w, h);
        f.cand_ang[slot] = D::to_f(p[static_cast<long long>(d.channels - 1) * sc]);  // nms.py:146 x[:, -1:]
      } else {
        f.cand_box[slot] = f.boxes_xyxy ? make_float4(cx, cy, w, h) : corners_in_dtype<DT>(cx, cy, w, h);  // nms.py:86
      }
    }
  }

  int total;
  int off = block_exclusive_scan(my_rows, total);
  if (total == 0) return;
  const int base = reserve_rows(f.row_count, b, total);
  if (my_rows == 0) return;
  uint64_t* keys = f.keys + static_cast<long long>(b) * f.rows_cap;
  int pos = base + off;
  const uint32_t row0 = static_cast<uint32_t>(a) << f.cls_bits;
  if constexpr (MULTI) {
    for (int c = 0; c < nc; ++c) {
      float s = D::to_f(pc[c * sc]);
      if (s > conf && class_allowed(f.class_mask, c)) {
        if (pos < f.rows_cap) keys[pos] = make_key(s, row0 + c);
        ++pos;
      }
    }
  } else {
    if (pos < f.rows_cap) keys[pos] = make_key(best, row0 + bcls);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// stand-alone pieces of the decode, for callers that compose it themselves (YOLOEDetect.forward_lrpc, head.py:1777-1813,
// calls self.dfl and self.decode_bboxes directly):
//   dfl_kernel        DFL.forward, block.py:250-253: (B, 4*16, A) -> (B, 4, A), softmax expectation per side
//   dist2bbox_kernel  Detect.decode_bboxes == tal.py:367-376 dist2bbox(dim=1) and OBB.decode_bboxes == tal.py:385-403
//                     dist2rbox(dim=1): every torch op of the reference is one separately rounded operation in the
//                     tensor dtype T, so the axis-aligned result is bit-identical to the reference (IEEE add/sub/mul).
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256)
dfl_kernel(const void* __restrict__ x_v, long long sb, long long sc, int anchors, void* __restrict__ out_v, long long osb,
           long long osc) {
  using T = typename DType<DT>::type;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= anchors) return;
  const int side = blockIdx.y, b = blockIdx.z;
  const T* x = static_cast<const T*>(x_v) + static_cast<long long>(b) * sb + static_cast<long long>(side * 16) * sc + a;
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = DType<DT>::to_f(x[static_cast<long long>(k) * sc]);
  static_cast<T*>(out_v)[static_cast<long long>(b) * osb + static_cast<long long>(side) * osc + a] = DType<DT>::from_f(dfl_expect<16>(v));
}

template <int DT>
__global__ void __launch_bounds__(256)
dist2bbox_kernel(const Dist2BoxArgs d) {
  using T = typename DType<DT>::type;
  using D = DType<DT>;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= d.anchors) return;
  const int b = blockIdx.y;
  const T* dist = static_cast<const T*>(d.dist) + static_cast<long long>(b) * d.dsb + a;
  const T* ap = static_cast<const T*>(d.anchor_points) + static_cast<long long>(b) * d.asb + static_cast<long long>(a) * d.asa;
  T* out = static_cast<T*>(d.out) + static_cast<long long>(b) * d.osb + a;
  const float l = D::to_f(dist[0]), t = D::to_f(dist[d.dsc]), r = D::to_f(dist[2 * d.dsc]), bt = D::to_f(dist[3 * d.dsc]);
  const float ax = D::to_f(ap[0]), ay = D::to_f(ap[d.asc]);
  float o0, o1, o2, o3;
  if (d.angle) {  // tal.py:397-403
    const float th = D::to_f(static_cast<const T*>(d.angle)[static_cast<long long>(b) * d.angle_sb + a]);
    const float co = D::rnd(cosf(th)), si = D::rnd(sinf(th));
    const float xf = D::rnd(__fmul_rn(D::rnd(__fsub_rn(r, l)), 0.5f)), yf = D::rnd(__fmul_rn(D::rnd(__fsub_rn(bt, t)), 0.5f));
    const float x = D::rnd(__fsub_rn(D::rnd(__fmul_rn(xf, co)), D::rnd(__fmul_rn(yf, si))));
    const float y = D::rnd(__fadd_rn(D::rnd(__fmul_rn(xf, si)), D::rnd(__fmul_rn(yf, co))));
    o0 = D::rnd(__fadd_rn(x, ax)); o1 = D::rnd(__fadd_rn(y, ay));
    o2 = D::rnd(__fadd_rn(l, r)); o3 = D::rnd(__fadd_rn(t, bt));
  } else {        // tal.py:369-376
    const float x1 = D::rnd(__fsub_rn(ax, l)), y1 = D::rnd(__fsub_rn(ay, t));
    const float x2 = D::rnd(__fadd_rn(ax, r)), y2 = D::rnd(__fadd_rn(ay, bt));
    if (d.xywh) {
      o0 = D::rnd(__fmul_rn(D::rnd(__fadd_rn(x1, x2)), 0.5f)); o1 = D::rnd(__fmul_rn(D::rnd(__fadd_rn(y1, y2)), 0.5f));
      o2 = D::rnd(__fsub_rn(x2, x1)); o3 = D::rnd(__fsub_rn(y2, y1));
    } else {
      o0 = x1; o1 = y1; o2 = x2; o3 = y2;
    }
  }
  out[0] = D::from_f(o0); out[d.osc] = D::from_f(o1); out[2 * d.osc] = D::from_f(o2); out[3 * d.osc] = D::from_f(o3);
}

cudaError_t launch_dfl(const void* x, int dtype, int batch, int anchors, long long sb, long long sc, void* out, long long osb,
                       long long osc, cudaStream_t st) {
  if (batch <= 0 || anchors <= 0) return cudaSuccess;
  dim3 grid((anchors + 255) / 256, 4, batch);
  switch (dtype) {
    case YPB_F32: dfl_kernel<YPB_F32><<<grid, 256, 0, st>>>(x, sb, sc, anchors, out, osb, osc); break;
    case YPB_F16: dfl_kernel<YPB_F16><<<grid, 256, 0, st>>>(x, sb, sc, anchors, out, osb, osc); break;
    case YPB_BF16: dfl_kernel<YPB_BF16><<<grid, 256, 0, st>>>(x, sb, sc, anchors, out, osb, osc); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_dist2bbox(const Dist2BoxArgs& d, int dtype, cudaStream_t st) {
  if (d.batch <= 0 || d.anchors <= 0) return cudaSuccess;
  dim3 grid((d.anchors + 255) / 256, d.batch);
  switch (dtype) {
    case YPB_F32: dist2bbox_kernel<YPB_F32><<<grid, 256, 0, st>>>(d); break;
    case YPB_F16: dist2bbox_kernel<YPB_F16><<<grid, 256, 0, st>>>(d); break;
    case YPB_BF16: dist2bbox_kernel<YPB_BF16><<<grid, 256, 0, st>>>(d); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// exhaustive monotonicity check of round_T(sigmoid_f(x)) over every finite fp32 x
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void sigmoid_monotone_kernel(unsigned long long* violations) {
  const unsigned long long total = 0xffffffffull;  // pairs (o, o+1) in orderable space
  unsigned long long bad = 0;
  for (unsigned long long o = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    float x = from_orderable_bits(static_cast<uint32_t>(o));
    float y = from_orderable_bits(static_cast<uint32_t>(o + 1));
    if (!isfinite(x) || !isfinite(y)) continue;
    float sx = DType<DT>::rnd(sigmoid_f(x)), sy = DType<DT>::rnd(sigmoid_f(y));
    if (!(sx <= sy)) ++bad;
  }
  if (bad) atomicAdd(violations, bad);
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
template <int DT_IN, int DT_OUT, int VEC>
static cudaError_t decode_dense_dispatch(const HeadGeom& g, const void* angle, int angle_is_logit, int append_angle,
                                         int xyxy, void* out, long long osb, long long osc, cudaStream_t st) {
  const int groups = g.group_start[g.num_levels];
  dim3 grid((groups + DEC_THREADS - 1) / DEC_THREADS, g.batch);
#ifndef YPB_DD16_OLD
  if constexpr (DT_IN != YPB_F32 && DT_IN == DT_OUT && VEC == 4) {
    if (angle)
      decode_dense16_kernel<DT_IN, MODE_ROT><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, append_angle, out, osb, osc);
    else if (xyxy)
      decode_dense16_kernel<DT_IN, MODE_XYXY><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, append_angle, out, osb, osc);
    else
      decode_dense16_kernel<DT_IN, MODE_XYWH><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, append_angle, out, osb, osc);
    return cudaGetLastError();
  }
#endif
  if (angle)
    decode_dense_kernel<DT_IN, DT_OUT, VEC, 16, MODE_ROT><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, append_angle, out, osb, osc);
  else if (xyxy)
    decode_dense_kernel<DT_IN, DT_OUT, VEC, 16, MODE_XYXY><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, append_angle, out, osb, osc);
  else
    decode_dense_kernel<DT_IN, DT_OUT, VEC, 16, MODE_XYWH><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, append_angle, out, osb, osc);
  return cudaGetLastError();
}

cudaError_t launch_decode_dense(const HeadGeom& g, int in_dtype, const void* angle, int angle_is_logit,
                                int append_angle, int xyxy, void* out, int out_dtype, long long osb, long long osc,
                                int vec, cudaStream_t st) {
  if (in_dtype != out_dtype) return cudaErrorInvalidValue;  // Detect._inference returns the input dtype (head.py:169)
#define YPB_DD(DT, V) return decode_dense_dispatch<DT, DT, V>(g, angle, angle_is_logit, append_angle, xyxy, out, osb, osc, st)
  if (in_dtype == YPB_F32) {
    if (vec == 4) YPB_DD(YPB_F32, 4);
    if (vec == 1) YPB_DD(YPB_F32, 1);
  } else if (in_dtype == YPB_BF16) {
    if (vec == 8) YPB_DD(YPB_BF16, 8);
    if (vec == 4) YPB_DD(YPB_BF16, 4);
    if (vec == 1) YPB_DD(YPB_BF16, 1);
  } else if (in_dtype == YPB_F16) {
    if (vec == 8) YPB_DD(YPB_F16, 8);
    if (vec == 4) YPB_DD(YPB_F16, 4);
    if (vec == 1) YPB_DD(YPB_F16, 1);
  }
#undef YPB_DD
  return cudaErrorInvalidValue;
}

template <int DT_IN, int DT_VAL, int VEC>
static cudaError_t filter_head_dispatch(const HeadGeom& g, const void* angle, int angle_is_logit, const FilterArgs& f,
                                        int which, cudaStream_t st) {
  const int groups = g.group_start[g.num_levels];
  const int blocks_x = (groups + DEC_THREADS - 1) / DEC_THREADS;
  if (which == 1) {
    dim3 grid(blocks_x, g.batch);
#define YPB_SC(M, R) scan_classes_kernel<DT_IN, DT_VAL, VEC, M, R><<<grid, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, f)
    if (f.multi_label) { if (f.rotated) YPB_SC(true, true); else YPB_SC(true, false); }
    else               { if (f.rotated) YPB_SC(false, true); else YPB_SC(false, false); }
#undef YPB_SC
    return cudaGetLastError();
  }
  if (f.fuse_decode) return cudaSuccess;  // kernel 1 decoded its own survivors
  // kernel 2: grid-stride over the octet list; enough CTAs to cover the GPU, never more than there can be octets
  const int G = blocks_x * DEC_THREADS;
  int blocks = 148 * 8;
  const int max_blocks = (f.tile_cap + 15) / 16;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  if (f.rotated) decode_tiles_kernel<DT_IN, DT_VAL, VEC, true><<<blocks, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, f, G);
  else           decode_tiles_kernel<DT_IN, DT_VAL, VEC, false><<<blocks, DEC_THREADS, 0, st>>>(g, angle, angle_is_logit, f, G);
  return cudaGetLastError();
}

cudaError_t launch_filter_from_head(const HeadGeom& g, int in_dtype, int value_dtype, const void* angle,
                                    int angle_is_logit, const FilterArgs& f, int vec, int which, cudaStream_t st) {
  if (in_dtype != value_dtype) return cudaErrorInvalidValue;
#define YPB_FD(DT, V) return filter_head_dispatch<DT, DT, V>(g, angle, angle_is_logit, f, which, st)
  if (in_dtype == YPB_F32) {
    if (vec == 4) YPB_FD(YPB_F32, 4);
    if (vec == 1) YPB_FD(YPB_F32, 1);
  } else if (in_dtype == YPB_BF16) {
    if (vec == 8) YPB_FD(YPB_BF16, 8);
    if (vec == 1) YPB_FD(YPB_BF16, 1);
  } else if (in_dtype == YPB_F16) {
    if (vec == 8) YPB_FD(YPB_F16, 8);
    if (vec == 1) YPB_FD(YPB_F16, 1);
  }
#undef YPB_FD
  return cudaErrorInvalidValue;
}

template <int DT>
static cudaError_t filter_dense_dispatch(const ypb_dense_desc& d, const FilterArgs& f, cudaStream_t st) {
  const int lanes = d.anchor_subset ? d.subset_len : d.anchors;
  if (lanes <= 0) return cudaSuccess;
  dim3 grid((lanes + DEC_THREADS - 1) / DEC_THREADS, d.batch);
  // (a form that stages the rows of 128 anchors in shared memory with coalesced loads for channel-contiguous tensors - the
  // (B, A, 4+nc) layout of the end2end head - was built and measured: no faster than the per-anchor walk, whose sectors are
  // served by L1 after the first touch; removed)
#define YPB_FDN(R, M) filter_from_dense_kernel<DT, R, M><<<grid, DEC_THREADS, 0, st>>>(d, f)
  if (f.rotated) { if (f.multi_label) YPB_FDN(true, true); else YPB_FDN(true, false); }
  else           { if (f.multi_label) YPB_FDN(false, true); else YPB_FDN(false, false); }
#undef YPB_FDN
  return cudaGetLastError();
}

cudaError_t launch_filter_from_dense(const ypb_dense_desc& d, const FilterArgs& f, cudaStream_t st) {
  switch (d.dtype) {
    case YPB_F32: return filter_dense_dispatch<YPB_F32>(d, f, st);
    case YPB_F16: return filter_dense_dispatch<YPB_F16>(d, f, st);
    case YPB_BF16: return filter_dense_dispatch<YPB_BF16>(d, f, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_sigmoid_selftest(int dtype, unsigned long long* violations, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(violations, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  const int blocks = 148 * 16;
  switch (dtype) {
    case YPB_F32: sigmoid_monotone_kernel<YPB_F32><<<blocks, 256, 0, st>>>(violations); break;
    case YPB_F16: sigmoid_monotone_kernel<YPB_F16><<<blocks, 256, 0, st>>>(violations); break;
    case YPB_BF16: sigmoid_monotone_kernel<YPB_BF16><<<blocks, 256, 0, st>>>(violations); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace ypb
