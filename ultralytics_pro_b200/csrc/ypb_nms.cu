// Sort + suppression + gather kernel of libyolopost_b200 (sm_100a): one CTA per image.
//
//   stage 1  rank the candidate rows of the image by their unique 64-bit key (score desc, row asc):
//            <= SORT_SMEM_MAX rows -> bitonic network in shared memory; more -> 8-bit LSD radix passes in global
//            memory (ping-pong), skipping digit positions that are constant over the image.       (nms.py:137-141)
//   stage 2  walk the ranks in chunks of CH: each rank is first tested against the rows already kept (<= max_det,
//            the only ones that can still matter), then the chunk resolves its internal order dependence with a
//            bitmask fix-point; stops as soon as max_det rows are kept (nms.py:157 keeps only those).  Fast-NMS rules
//            (rotated ProbIoU, box_iou) have no dependence: a rank is tested against every higher rank. (nms.py:143-156)
//   stage 3  gather the kept rows (nms.py:159-161).
//
// All IoU arithmetic uses explicit round-to-nearest intrinsics in the reference's operation order (no FMA
// contraction), so kept sets are bit-exact against torchvision.ops.nms / TorchNMS (SURVEY.md section 7 "Hard parts").
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

#include "ypb_common.cuh"

namespace cg = cooperative_groups;

namespace ypb {

constexpr int NT = 512;              // threads per CTA
constexpr int NW = NT / 32;          // warps per CTA
constexpr int CH = 128;              // ranks per chunk
constexpr int CW = CH / 32;          // mask words per rank of a chunk
constexpr int PARTS = NT / CH;       // threads cooperating on one rank (== CW: part q owns mask word q)
constexpr int SORT_SMEM_MAX = 4096;  // rows sorted in shared memory
constexpr int RANK_COUNT_MAX = 1024;  // kept rows ranked by counting; beyond that they are sorted
constexpr int CW_CLASS_MAX = 1024;    // class histogram bins of the class-wise path
constexpr int KEPT_SMEM = 1024;      // kept rows held in shared memory (larger max_det spills the list to the workspace)
static_assert(KEPT_SMEM * 16 >= 2 * RANK_COUNT_MAX * 8, "kept-key scratch");
static_assert(PARTS == CW && CW == 4, "one thread per (rank, mask word); rows are read as one uint4");

struct __align__(16) Smem {
  union {
    uint64_t keys[SORT_SMEM_MAX];  // small path: sorted keys live here for the whole kernel
    struct {
      uint32_t hist[256];
      uint32_t wc[NW * 256];
    } rx;
  } u;
  alignas(16) uint32_t mask[CW * CH];  // mask[t * CW + w]: bit i set iff rank t of the chunk suppresses rank w*32+i (> t)
  union {
    struct {
      float4 box[CH];
      float area[CH];
    } g;
    float rec[CH * 8];
  } c;
  int dead[CH];  // rank already suppressed by a row outside the chunk
  float4 kbox[KEPT_SMEM];  // class-offset boxes of the rows kept so far (greedy rule, max_det <= KEPT_SMEM)
  float karea[KEPT_SMEM];
  // class-wise path: every candidate of the image, in (class, score desc) order
  float4 cbox[SORT_SMEM_MAX];
  float carea[SORT_SMEM_MAX];
  uint32_t kbits[SORT_SMEM_MAX / 32];  // kept flags by slot in (class, rank) order
  uint16_t order[SORT_SMEM_MAX];       // slot in (class, rank) order -> scattered position
  int ccount[CW_CLASS_MAX];            // rows per class
  int ckept[CW_CLASS_MAX];             // kept rows per class so far
  uint16_t krank[SORT_SMEM_MAX];       // per class (at its slot range): ranks of the kept rows, in order
  int cstart[CW_CLASS_MAX + 1];        // first slot of each class
  float red_min[NW], red_max[NW];
  uint16_t clist[CW_CLASS_MAX];        // non-empty classes
  uint16_t clist2[CW_CLASS_MAX];       // the same, longest first
  int nseg, kcount, next_seg, longest;
  uint32_t alive_bits[CW];
  uint32_t kept_bits[CW];
  uint32_t undec_bits[CW];
  unsigned long long red_and, red_or;
};

// ---------------------------------------------------------------------------------------------------------------
// pair predicates
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float box_area(const float4& b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

// torchvision nms_kernel_impl / nms.py:276-294: suppress iff fl(inter / (area_i + area_j - inter)) > thr, no eps.
// The IEEE quotient is never formed: for inter > 0 and union >= 0, round-to-nearest-even gives
//   fl(inter/union) > thr  <=>  inter/union > mid  (or == mid when thr's mantissa is odd),  mid = (thr + next(thr)) / 2,
// and  inter >< mid * union  is decided exactly in fp64 (25-bit x 24-bit product).  Branch-free, so the 32 pair tests of
// a mask word pipeline.  inter == 0 gives a quotient of +-0 or NaN (never > thr >= 0); a negative union a negative
// quotient; NaN operands compare false - all "not suppressed", as in the reference.
struct GreedyThr {
  double mid;
  bool tie_up;
  float lo, hi;  // thr * (1 -+ 2^-20): an fp32 product against these brackets decides all but borderline pairs
};
__device__ __forceinline__ GreedyThr make_greedy_thr(float thr) {
  const float nxt = __uint_as_float(__float_as_uint(thr) + 1u);  // thr >= 0
  GreedyThr g;
  g.mid = (static_cast<double>(thr) + static_cast<double>(nxt)) * 0.5;
  g.tie_up = (__float_as_uint(thr) & 1u) != 0u;
  // tiny thresholds: no bracket (every overlapping pair takes the exact test)
  g.lo = thr > 1e-6f ? thr * (1.0f - 9.5367431640625e-7f) : 0.f;
  g.hi = thr > 1e-6f ? thr * (1.0f + 9.5367431640625e-7f) : INFINITY;
  return g;
}
// exact decision (see above); only reached by borderline pairs
__device__ __noinline__ bool greedy_exact(float inter, float uni, double mid, bool tie_up) {
  const double di = static_cast<double>(inter), lim = mid * static_cast<double>(uni);
  const bool over = tie_up ? di >= lim : di > lim;
  return inter > 0.f && uni >= 0.f && over;
}
struct PairGeom { float inter, uni; };
__device__ __forceinline__ PairGeom pair_geom(const float4& a, float aa, const float4& b, float ab) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  PairGeom g;
  g.inter = __fmul_rn(w, h);
  g.uni = __fsub_rn(__fadd_rn(aa, ab), g.inter);
  return g;
}
// Classifies a pair with fp32 products only: `sure` = suppresses, `maybe` = borderline (needs greedy_exact), neither =
// surely not.  With uni > 1e-30 the products thr*(1+-2^-20)*uni carry a relative error <= 2^-23, far inside the 2^-20
// bracket around mid = thr*(1 + ~2^-24).  Written with non-short-circuit logic so it compiles to predicates, not branches.
__device__ __forceinline__ void greedy_flags(const PairGeom& g, const GreedyThr& t, unsigned& sure, unsigned& maybe) {
  const unsigned normal = g.uni > 1e-30f;
  const unsigned above = g.inter > __fmul_rn(t.hi, g.uni);
  const unsigned below = g.inter < __fmul_rn(t.lo, g.uni);
  const unsigned pos = g.inter > 0.f;
  sure = normal & above;
  const unsigned never = (pos ^ 1u) | (normal & below);
  maybe = (sure | never) ^ 1u;
}
__device__ __forceinline__ bool greedy_suppresses(const float4& a, float aa, const float4& b, float ab, const GreedyThr& t) {
  const PairGeom g = pair_geom(a, aa, b, ab);
  unsigned sure, maybe;
  greedy_flags(g, t, sure, maybe);
  if (maybe) return greedy_exact(g.inter, g.uni, t.mid, t.tie_up);
  return sure != 0u;
}

// metrics.py:54-75 with eps in the denominator.
__device__ __forceinline__ float boxiou_value(const float4& a, float aa, const float4& b, float ab) {
  float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(aa, ab), inter), 1e-7f));
}
// rule ">= thr" (nms.py:223)
__device__ __forceinline__ bool boxiou_suppresses(const float4& a, float aa, const float4& b, float ab, float thr) {
  return boxiou_value(a, aa, b, ab) >= thr;
}

struct ObbRec { float x, y, a, b, c, det; };

// metrics.py:187-203 on one xywhr box (+ the per-box factor of metrics.py:279).
__device__ __forceinline__ ObbRec obb_record(float x, float y, float w, float h, float r) {
  ObbRec o;
  o.x = x; o.y = y;
  float a = __fdiv_rn(__fmul_rn(w, w), 12.f), b = __fdiv_rn(__fmul_rn(h, h), 12.f);
  float co = cosf(r), si = sinf(r);
  float c2 = __fmul_rn(co, co), s2 = __fmul_rn(si, si);
  o.a = __fadd_rn(__fmul_rn(a, c2), __fmul_rn(b, s2));
  o.b = __fadd_rn(__fmul_rn(a, s2), __fmul_rn(b, c2));
  o.c = __fmul_rn(__fmul_rn(__fsub_rn(a, b), co), si);
  o.det = fmaxf(__fsub_rn(__fmul_rn(o.a, o.b), __fmul_rn(o.c, o.c)), 0.f);
  return o;
}

// metrics.py:251-284, p = obb1, q = obb2.
__device__ __forceinline__ float probiou_value(const ObbRec& p, const ObbRec& q) {
  const float eps = 1e-7f;
  float sa = __fadd_rn(p.a, q.a), sb = __fadd_rn(p.b, q.b), sc = __fadd_rn(p.c, q.c);
  float dy = __fsub_rn(p.y, q.y), dx = __fsub_rn(p.x, q.x);
  float den = __fsub_rn(__fmul_rn(sa, sb), __fmul_rn(sc, sc));
  float dene = __fadd_rn(den, eps);
  float t1 = __fmul_rn(__fdiv_rn(__fadd_rn(__fmul_rn(sa, __fmul_rn(dy, dy)), __fmul_rn(sb, __fmul_rn(dx, dx))), dene), 0.25f);
  float t2 = __fmul_rn(__fdiv_rn(__fmul_rn(__fmul_rn(sc, __fsub_rn(q.x, p.x)), dy), dene), 0.5f);
  float root = sqrtf(__fmul_rn(p.det, q.det));
  float t3 = __fmul_rn(logf(__fadd_rn(__fdiv_rn(den, __fadd_rn(__fmul_rn(4.f, root), eps)), eps)), 0.5f);
  float bd = __fadd_rn(__fadd_rn(t1, t2), t3);
  bd = bd != bd ? bd : fminf(fmaxf(bd, eps), 100.f);
  float hd = sqrtf(__fadd_rn(__fsub_rn(1.0f, expf(-bd)), eps));
  return __fsub_rn(1.0f, hd);
}
// p = higher-ranked row, q = lower-ranked; rule ">= thr" (nms.py:223).
__device__ __forceinline__ bool probiou_suppresses(const ObbRec& p, const ObbRec& q, float thr) {
  return probiou_value(p, q) >= thr;
}

// ---------------------------------------------------------------------------------------------------------------
// stage 1: ranking
// ---------------------------------------------------------------------------------------------------------------
// Bitonic sort of n <= SORT_SMEM_MAX keys, ascending, result in s[0..P), P = max(256, next power of two).
// Blocked layout: thread tid holds the KPT = 8 consecutive keys at positions tid*8 .. tid*8+7 in registers, so a
// compare-exchange at distance j is
//   j < 8          : between two registers of the same thread (27 of the 55 stages at P = 1024; full ILP)
//   8 <= j < 256   : a lane exchange (shuffle) between threads of one warp, 8 independent exchanges in flight
//   j >= 256       : a round trip through shared memory (2 barriers) - 3 stages at P = 1024, none at P = 256.
// Only P/8 threads take part (one warp up to 256 keys); the others just meet the barriers.
constexpr int KPT = 8;

__device__ __forceinline__ uint64_t cmpx(uint64_t mine, uint64_t other, bool want_min) {
  const bool take_other = want_min ? other < mine : other > mine;
  return take_other ? other : mine;
}

struct KeyIdentity {
  __device__ __forceinline__ uint64_t operator()(uint64_t k) const { return k; }
};
// (score desc, row asc) key -> (class asc, score desc, anchor asc): same order inside a class, classes contiguous.
struct KeyClassMajor {
  int cbits, abits;
  __device__ __forceinline__ uint64_t operator()(uint64_t k) const {
    const uint32_t row = static_cast<uint32_t>(k);
    const uint64_t cls = row & ((1u << cbits) - 1u), anchor = row >> cbits;
    return (cls << (32 + abits)) | ((k >> 32) << abits) | anchor;
  }
  __device__ __forceinline__ uint64_t inverse(uint64_t c) const {
    const uint64_t anchor = c & ((1ull << abits) - 1ull), cls = c >> (32 + abits);
    const uint64_t hi = (c >> abits) & 0xffffffffull;
    return (hi << 32) | (anchor << cbits) | cls;
  }
};

template <typename XF>
__device__ __forceinline__ void bitonic_sort(uint64_t* s, const uint64_t* src, int n, XF xf) {
  int P = 32 * KPT;
  while (P < n) P <<= 1;
  const int tid = threadIdx.x;
  const int T = P / KPT;  // participating threads (a multiple of 32: whole warps take part or idle)
  const bool active = __all_sync(0xffffffffu, tid < T) != 0;  // vote result: warp-uniform for the compiler too
  uint64_t key[KPT];
#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int p = tid * KPT + u;
    key[u] = (active && p < n) ? xf(src[p]) : KEY_SENTINEL;
  }
  __syncthreads();  // src may alias shared memory that the exchange stages are about to overwrite
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < KPT) {
        if (active) {
#pragma unroll
          for (int jb = KPT / 2; jb >= 1; jb >>= 1) {  // compile-time distance: registers stay registers
            if (j == jb) {
#pragma unroll
              for (int u = 0; u < KPT; ++u) {
                if ((u & jb) == 0) {
                  const bool up = ((tid * KPT + u) & k) == 0;
                  const uint64_t a = key[u], c = key[u | jb];
                  const bool swap = (a > c) == up;
                  key[u] = swap ? c : a;
                  key[u | jb] = swap ? a : c;
                }
              }
            }
          }
        }
      } else if (j < 32 * KPT) {
        if (active) {
          const int lane_xor = j / KPT;
          const bool lower = (tid & lane_xor) == 0;
#pragma unroll
          for (int u = 0; u < KPT; ++u) {
            const bool up = ((tid * KPT + u) & k) == 0;
            const uint64_t other = __shfl_xor_sync(0xffffffffu, key[u], lane_xor);
            key[u] = cmpx(key[u], other, lower == up);
          }
        }
      } else {
        if (active) {
#pragma unroll
          for (int u = 0; u < KPT; ++u) s[tid * KPT + u] = key[u];
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int u = 0; u < KPT; ++u) {
            const int p = tid * KPT + u;
            const bool up = (p & k) == 0, lower = (p & j) == 0;
            key[u] = cmpx(key[u], s[p ^ j], lower == up);
          }
        }
        __syncthreads();
      }
    }
  }
  if (active) {
#pragma unroll
    for (int u = 0; u < KPT; ++u) s[tid * KPT + u] = key[u];
  }
  __syncthreads();
}

// One stable 8-bit LSD pass src -> dst over n keys.
__device__ void radix_pass(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n, int shift, Smem& sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* hist = sm.u.rx.hist;
  uint32_t* wc = sm.u.rx.wc;
  for (int i = threadIdx.x; i < 256; i += NT) hist[i] = 0;
  __syncthreads();
  for (int base = 0; base < n; base += NT) {
    const int i = base + threadIdx.x;
    const bool act = i < n;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if (act) {
      const uint32_t d = static_cast<uint32_t>(src[i] >> shift) & 255u;
      const unsigned peers = __match_any_sync(am, d);
      if ((peers & lt_mask) == 0) atomicAdd(&hist[d], __popc(peers));
    }
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 256 bins
    uint32_t v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = hist[lane * 8 + k]; s += v[k]; }
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    uint32_t run = inc - s;
#pragma unroll
    for (int k = 0; k < 8; ++k) { hist[lane * 8 + k] = run; run += v[k]; }
  }
  __syncthreads();
  for (int base = 0; base < n; base += NT) {
    for (int k = threadIdx.x; k < NW * 256; k += NT) wc[k] = 0;
    __syncthreads();
    const int i = base + threadIdx.x;
    const bool act = i < n;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    uint64_t key = 0;
    uint32_t d = 0, rank = 0;
    if (act) {
      key = src[i];
      d = static_cast<uint32_t>(key >> shift) & 255u;
      const unsigned peers = __match_any_sync(am, d);
      rank = __popc(peers & lt_mask);
      if (rank == 0) wc[warp * 256 + d] = __popc(peers);
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t run = hist[threadIdx.x];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        uint32_t c = wc[w * 256 + threadIdx.x];
        wc[w * 256 + threadIdx.x] = run;
        run += c;
      }
      hist[threadIdx.x] = run;
    }
    __syncthreads();
    if (act) dst[wc[warp * 256 + d] + rank] = key;
    __syncthreads();
  }
}

// Full ascending sort of n keys in global memory; returns the buffer holding the result.
__device__ const uint64_t* radix_sort_global(uint64_t* a, uint64_t* b, int n, Smem& sm) {
  // digit positions that are constant over the image need no pass
  unsigned long long vand = ~0ull, vor = 0ull;
  for (int i = threadIdx.x; i < n; i += NT) { uint64_t k = a[i]; vand &= k; vor |= k; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vand &= __shfl_xor_sync(0xffffffffu, vand, o);
    vor |= __shfl_xor_sync(0xffffffffu, vor, o);
  }
  if (threadIdx.x == 0) { sm.red_and = ~0ull; sm.red_or = 0ull; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { atomicAnd(&sm.red_and, vand); atomicOr(&sm.red_or, vor); }
  __syncthreads();
  const unsigned long long varying = sm.red_and ^ sm.red_or;
  uint64_t* src = a;
  uint64_t* dst = b;
  for (int shift = 0; shift < 64; shift += 8) {
    if (((varying >> shift) & 255ull) == 0) continue;  // uniform branch
    radix_pass(src, dst, n, shift, sm);
    uint64_t* t = src; src = dst; dst = t;
  }
  return src;
}

// Radix select of a PREFIX of the ascending key order: finds a threshold T such that k = |{key <= T}| lies in [lo, hi]
// (1 <= lo <= hi <= n) and copies those k keys, unordered, to out.  MSD digits of SEL_BITS bits: one histogram pass over
// the keys still matching the chosen high bits per digit; the pass ends the search as soon as a bucket boundary falls
// inside [lo, hi] - typically the first or second digit - else the straddling bucket is refined.  Keys are unique, so the
// search ends at the latest when all 64 bits are fixed.
constexpr int SEL_BITS = 11;
__device__ int select_prefix(const uint64_t* __restrict__ keys, int n, int lo, int hi, uint64_t* __restrict__ out, Smem& sm) {
  static_assert((1 << SEL_BITS) * 4 <= NW * 256 * 4, "digit histogram fits the radix scratch");
  uint32_t* hist = sm.u.rx.wc;  // (1 << SEL_BITS) bins
  __shared__ unsigned long long s_thr;
  __shared__ int s_state[3];    // [0] rows below the current bucket, [1] done flag, [2] compaction cursor
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned long long prefix = 0;  // chosen high bits, right-aligned
  int fixed = 0;                  // how many high bits are chosen
  int below = 0;                  // keys strictly below the current bucket
  while (true) {
    const int w = min(SEL_BITS, 64 - fixed);
    const int shift = 64 - fixed - w;
    const uint32_t dmask = (1u << w) - 1u;
    for (int i = tid; i < (1 << SEL_BITS); i += NT) hist[i] = 0;
    __syncthreads();
    for (int i0 = tid; i0 < n; i0 += 4 * NT) {  // 4 independent (L2) loads in flight per thread
      unsigned long long k4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) k4[u] = i0 + u * NT < n ? keys[i0 + u * NT] : 0ull;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned long long k = k4[u];
        if (i0 + u * NT < n && (fixed == 0 || (k >> (64 - fixed)) == prefix)) atomicAdd(&hist[static_cast<uint32_t>(k >> shift) & dmask], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {  // one warp scans the bins in order: first bin whose inclusive cumulative count reaches lo
      int run = below, found = -1, cum_at = 0, before_at = 0;
      for (int d0 = 0; d0 < (1 << w) && found < 0; d0 += 32) {
        const int v = static_cast<int>(hist[d0 + lane]);
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const unsigned hitm = __ballot_sync(0xffffffffu, run + inc >= lo);
        if (hitm) {
          const int l = __ffs(hitm) - 1;
          found = d0 + l;
          cum_at = run + __shfl_sync(0xffffffffu, inc, l);
          before_at = cum_at - __shfl_sync(0xffffffffu, v, l);
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) {
        // found >= 0 always: the bucket holds >= lo - below keys by construction
        const unsigned long long np = (prefix << w) | static_cast<unsigned long long>(found);
        if (cum_at <= hi || shift == 0) {
          s_thr = shift == 0 ? np : ((np << shift) | ((1ull << shift) - 1ull));
          s_state[0] = cum_at;
          s_state[1] = 1;
        } else {
          s_thr = np;
          s_state[0] = before_at;
          s_state[1] = 0;
        }
        s_state[2] = 0;
      }
    }
    __syncthreads();
    const bool done = s_state[1] != 0;
    if (done) break;
    prefix = s_thr;
    below = s_state[0];
    fixed += w;
    __syncthreads();
  }
  const unsigned long long thr = s_thr;
  const int k = s_state[0];
  for (int base = 0; base < n; base += NT) {  // unordered compaction, one shared-memory atomic per warp
    const int i = base + tid;
    const unsigned long long key = i < n ? keys[i] : ~0ull;
    const bool take = i < n && key <= thr;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int pos = 0;
    if (lane == 0 && m) pos = atomicAdd(&s_state[2], __popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (take) out[pos + __popc(m & ((1u << lane) - 1u))] = key;
  }
  __syncthreads();
  return k;
}

// Optional phase timestamps (diagnostic, ypb_debug_set_phase_buffer): per CTA 32 clock64 marks.
static std::atomic<long long*> g_phase_buf{nullptr};
void set_phase_buffer(long long* p) { g_phase_buf.store(p, std::memory_order_relaxed); }
#define YPB_MARK(slot)                                                                   \
  do {                                                                                   \
    if (a.dbg && threadIdx.x == 0 && (slot) < 32) a.dbg[blockIdx.x * 32 + (slot)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// class-wise greedy NMS (fast path of the class-aware rule)
//
// nms.py:143-149 makes NMS class-aware by adding cls * max_wh to every coordinate.  When all un-offset coordinates of
// the image span no more than max_wh, boxes of different classes cannot intersect after the offset (the offsets differ
// by >= max_wh and fp32 rounding is monotone), so their IoU is exactly 0 and the greedy walk decomposes EXACTLY into
// independent walks per class:
//   1. counting sort of the candidates by class (shared-memory histogram + scatter), boxes gathered on the way;
//   2. every row is ranked inside its class by counting (all threads), then one warp per class (handed out
//      dynamically, longest class first) walks it with lane = row: the lowest surviving lane is kept and one ballot
//      strikes what it suppresses - one step per KEPT box instead of one IoU per pair of candidates;
//   3. the kept rows of all classes are ranked by the original (score desc, row asc) key by counting and cut at
//      max_det (nms.py:157).
// Same fp32 arithmetic on the same offset boxes as the dense path, so results are bit-identical.  Returns the number
// of kept rows (their keys, in order, in sm.u.keys) or -1 when a precondition fails (span, a class with more than
// CW_SEG_MAX rows, non-finite boxes): the caller then takes the dense walk.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CW_SEG_MAX = 256;     // rows of one class a single warp ranks in registers (8 per lane)

// One greedy test of every lane's row against the same higher-ranked row (kb, ka) of its class.  Branch-free for the
// whole warp; only a borderline pair (|IoU - thr| within 2^-20 relative) takes the exact fp64 decision.
__device__ __forceinline__ bool lanes_suppressed_by(const float4& kb, float ka, const float4& mine, float marea,
                                                    const GreedyThr& t, bool consider) {
  const PairGeom g = pair_geom(kb, ka, mine, marea);
  unsigned su, mb;
  greedy_flags(g, t, su, mb);
  bool hit = consider && su != 0u;
  const bool border = consider && mb != 0u;
  if (__any_sync(0xffffffffu, border)) {
    if (border) hit = greedy_exact(g.inter, g.uni, t.mid, t.tie_up);
  }
  return hit;
}

__device__ int classwise_greedy(Smem& sm, const SuppressArgs& a, const uint64_t* ka, int n, const float4* cand_box,
                                const GreedyThr& gthr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int cbits = a.cls_bits;
  const uint32_t cmask = (1u << cbits) - 1u;
  const int nbins = 1 << cbits;
  YPB_MARK(16);

  // ---- 1. counting sort by class ----------------------------------------------------------------------------------
  for (int c = tid; c < nbins; c += NT) { sm.ccount[c] = 0; sm.ckept[c] = 0; }
  if (tid == 0) { sm.kcount = 0; sm.next_seg = 0; }
  for (int w = tid; w < (n + 31) / 32; w += NT) sm.kbits[w] = 0;
  __syncthreads();
  constexpr int PER = SORT_SMEM_MAX / NT;  // rows per thread
  uint64_t mykey[PER];
  int myslot[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int i = u * NT + tid;
    if (i < n) {
      mykey[u] = ka[i];
      myslot[u] = atomicAdd(&sm.ccount[static_cast<uint32_t>(mykey[u]) & cmask], 1);
    }
  }
  __syncthreads();
  if (warp == 0) {  // exclusive prefix over the class bins (32 per step); also the longest class
    int run = 0, longest = 0, nlist = 0;
    for (int c0 = 0; c0 < nbins; c0 += 32) {
      const int c = c0 + lane;
      const int v = c < nbins ? sm.ccount[c] : 0;
      longest = max(longest, v);
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (c < nbins) sm.cstart[c] = run + inc - v;
      run += __shfl_sync(0xffffffffu, inc, 31);
      const unsigned ne = __ballot_sync(0xffffffffu, v > 0);
      if (v > 0) sm.clist[nlist + __popc(ne & lt_mask)] = static_cast<uint16_t>(c);
      nlist += __popc(ne);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, o));
    if (lane == 0) { sm.cstart[nbins] = run; sm.longest = longest; sm.nseg = nlist; }
  }
  __syncthreads();
  // longest classes first: they bound the critical path of the dynamic hand-out (rank by counting, ties by class id)
  {
    const int nl = sm.nseg;
    for (int i = tid; i < nl; i += NT) {
      const int c = sm.clist[i], v = sm.ccount[c];
      int r = 0;
      for (int j = 0; j < nl; ++j) {
        const int cj = sm.clist[j], vj = sm.ccount[cj];
        r += (vj > v || (vj == v && cj < c)) ? 1 : 0;
      }
      sm.clist2[r] = static_cast<uint16_t>(c);
    }
  }
  __syncthreads();
  if (sm.longest > CW_SEG_MAX) return -1;  // uniform: a class too long for one warp - dense walk
  float lo = INFINITY, hi = -INFINITY;
  bool finite = true;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int i = u * NT + tid;
    if (i < n) {
      const uint32_t row = static_cast<uint32_t>(mykey[u]);
      const uint32_t cls = row & cmask, anchor = row >> cbits;
      const int pos = sm.cstart[cls] + myslot[u];
      const float4 bx = cand_box[anchor];
      lo = fminf(lo, fminf(fminf(bx.x, bx.y), fminf(bx.z, bx.w)));
      hi = fmaxf(hi, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
      finite &= isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w);
      const float off = __fmul_rn(static_cast<float>(cls), a.max_wh);  // nms.py:143
      const float4 ob = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
      sm.u.keys[pos] = mykey[u];
      sm.cbox[pos] = ob;
      sm.carea[pos] = box_area(ob);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { sm.red_min[warp] = lo; sm.red_max[warp] = hi; }
  const int bad = __syncthreads_or(!finite);
  float glo = sm.red_min[0], ghi = sm.red_max[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) { glo = fminf(glo, sm.red_min[w]); ghi = fmaxf(ghi, sm.red_max[w]); }
  if (bad || !(ghi - glo <= 0.999f * a.max_wh)) return -1;  // uniform
  YPB_MARK(19);

  // ---- 2a. rank every row inside its class, all threads in parallel: rank = number of rows of the class with a smaller
  //          (score desc, anchor asc) key.  Rows are class-contiguous, so the lanes of a warp mostly share a class and
  //          read the same shared-memory words (broadcast).  order[slot of (class, rank)] = scattered position.
  for (int p = tid; p < n; p += NT) {
    const uint64_t k = sm.u.keys[p];
    const uint32_t cls = static_cast<uint32_t>(k) & cmask;
    const uint64_t mine = (k & 0xffffffff00000000ull) | (static_cast<uint32_t>(k) >> cbits);
    const int s = sm.cstart[cls], e = sm.cstart[cls + 1];
    int rank = 0;
#pragma unroll 4
    for (int q = s; q < e; ++q) {
      const uint64_t o = sm.u.keys[q];
      rank += ((o & 0xffffffff00000000ull) | (static_cast<uint32_t>(o) >> cbits)) < mine ? 1 : 0;
    }
    sm.order[s + rank] = static_cast<uint16_t>(p);
  }
  __syncthreads();
  YPB_MARK(18);

  // ---- 2b. one warp per class, handed out dynamically (longest first) ---------------------------------------------------
  const int nseg = sm.nseg;
  while (true) {
    int sid = 0;
    if (lane == 0) sid = atomicAdd(&sm.next_seg, 1);
    sid = __shfl_sync(0xffffffffu, sid, 0);
    if (sid >= nseg) break;
    const int cid = sm.clist2[sid];
    const int s = sm.cstart[cid], e = sm.cstart[cid + 1];
    const int L = e - s;
    // walk the class in rank order, 32 ranks (one per lane) at a time
    for (int r0 = 0; r0 < L; r0 += 32) {
      const bool in = r0 + lane < L;
      float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
      float marea = 0.f;
      if (in) { const int p = sm.order[s + r0 + lane]; mine = sm.cbox[p]; marea = sm.carea[p]; }
      bool alive = in;
      // rows of this class kept in earlier groups: their boxes were parked in rank order at the front of the segment
      const int nk = sm.ckept[cid];
      for (int k0 = 0; k0 < nk; k0 += 32) {
        // lane l fetches kept row k0 + l (independent 3-hop shared-memory chains), then they are broadcast by shuffle
        float4 kb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ka = 0.f;
        if (k0 + lane < nk) {
          const int pk = sm.order[s + sm.krank[s + k0 + lane]];
          kb = sm.cbox[pk];
          ka = sm.carea[pk];
        }
        const int cnt = min(32, nk - k0);
        for (int k = 0; k < cnt; ++k) {
          const float4 b4 = make_float4(__shfl_sync(0xffffffffu, kb.x, k), __shfl_sync(0xffffffffu, kb.y, k),
                                        __shfl_sync(0xffffffffu, kb.z, k), __shfl_sync(0xffffffffu, kb.w, k));
          const float a1 = __shfl_sync(0xffffffffu, ka, k);
          const bool hit = lanes_suppressed_by(b4, a1, mine, marea, gthr, alive);  // all lanes call
          alive = alive && !hit;
        }
      }
      // inside the group: the lowest surviving lane is kept; its box reaches the others by shuffle and one ballot
      // strikes the lanes it suppresses
      uint32_t m = __ballot_sync(0xffffffffu, alive), keptw = 0;
      while (m) {
        const int i = __ffs(m) - 1;
        keptw |= 1u << i;
        const float4 kb = make_float4(__shfl_sync(0xffffffffu, mine.x, i), __shfl_sync(0xffffffffu, mine.y, i),
                                      __shfl_sync(0xffffffffu, mine.z, i), __shfl_sync(0xffffffffu, mine.w, i));
        const float ka = __shfl_sync(0xffffffffu, marea, i);
        const bool hit = lanes_suppressed_by(kb, ka, mine, marea, gthr, lane > i && ((m >> lane) & 1u) != 0u);
        m &= ~((1u << i) | __ballot_sync(0xffffffffu, hit));
      }
      // record the group's kept ranks (kept flag by slot, and the per-class list used by later groups)
      if ((keptw >> lane) & 1u) {
        sm.krank[s + nk + __popc(keptw & lt_mask)] = static_cast<uint16_t>(r0 + lane);
      }
      // every lane read ckept[cid] at the top of this group: order those reads before lane 0's update (the ballots in between
      // converge the warp but are not memory barriers; compute-sanitizer racecheck reports the pair without this)
      __syncwarp();
      if (lane == 0) {
        sm.ckept[cid] = nk + __popc(keptw);
        if (keptw) {
          // slots s+r0 .. s+r0+31 straddle at most two kbits words
          const int q0 = s + r0;
          atomicOr(&sm.kbits[q0 >> 5], keptw << (q0 & 31));
          if ((q0 & 31) && ((q0 >> 5) + 1) < SORT_SMEM_MAX / 32) atomicOr(&sm.kbits[(q0 >> 5) + 1], keptw >> (32 - (q0 & 31)));
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  YPB_MARK(20);

  // ---- 3. kept rows, ranked by (score desc, row asc) -------------------------------------------------------------------
  uint64_t* kkeys = reinterpret_cast<uint64_t*>(sm.kbox);  // KEPT_SMEM float4 = 2 * RANK_COUNT_MAX keys
  const int nwords = (n + 31) >> 5;
  for (int w = warp; w < nwords; w += NW) {
    const uint32_t kb = sm.kbits[w];
    int base = 0;
    if (lane == 0 && kb) base = atomicAdd(&sm.kcount, __popc(kb));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (((kb >> lane) & 1u) && base + __popc(kb & lt_mask) < RANK_COUNT_MAX * 2)
      kkeys[base + __popc(kb & lt_mask)] = sm.u.keys[sm.order[32 * w + lane]];
  }
  __syncthreads();
  const int kc = sm.kcount;
  YPB_MARK(21);
  if (kc > RANK_COUNT_MAX * 2) return -1;  // uniform; more kept rows than the ranking scratch holds - dense walk
  // the kept rows in (score desc, row asc) order; the first max_det are the result.  Few rows: rank by counting (rank =
  // number of kept rows with a smaller key, all threads in parallel).  Many (val mode: most of a 4096-row prefix survives):
  // counting is O(kc^2) - 100 us at kc = 2000 - so the keys take the same bitonic network as the candidates.
  uint64_t* out_keys = reinterpret_cast<uint64_t*>(sm.cbox);  // boxes are no longer needed
  __syncthreads();
  if (kc <= 256) {
    for (int i = tid; i < kc; i += NT) {
      const uint64_t mine = kkeys[i];
      int rank = 0;
#pragma unroll 4
      for (int j = 0; j < kc; ++j) rank += kkeys[j] < mine ? 1 : 0;
      if (rank < a.max_det) out_keys[rank] = mine;
    }
    __syncthreads();
  } else {
    bitonic_sort(out_keys, kkeys, kc, KeyIdentity{});  // ascending keys == rank order; ends with a barrier
  }
  YPB_MARK(22);
  return min(kc, a.max_det);
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ float load_as_float(const void* p, long long off) {
  return DType<DT>::to_f(static_cast<const typename DType<DT>::type*>(p)[off]);
}
__device__ __forceinline__ float load_pred(const void* p, int dt, long long off) {
  if (dt == YPB_F32) return load_as_float<YPB_F32>(p, off);
  if (dt == YPB_F16) return load_as_float<YPB_F16>(p, off);
  return load_as_float<YPB_BF16>(p, off);
}

// One-sided gather over NVLink (ypb_nms_out.peer_*), image b: the kept rows are contiguous floats in the local result buffer;
// copy them and the count into ring entry (seq % depth) of every peer's buffer (this rank's own included) with plain
// peer-mapped stores, then the last CTA of the launch publishes the launch sequence number in every peer's arrival flag.
// Back-pressure: entry seq % depth was last filled by launch seq - depth; a peer has released it once its acknowledgement
// (written into OUR buffer by its ypb_peer_wait) reached seq - depth.  Called by all `nthr` threads of a CTA.
__device__ void peer_push(const SuppressArgs& a, int b, int kept_n, int nthr) {
  const int tid = threadIdx.x;
  const int cols = 6 + a.extra;
    __shared__ int s_seq;
    if (tid == 0) s_seq = a.peer_state[1] + 1;  // stable during the launch: only its LAST CTA advances peer_state[1]
    __syncthreads();  // also: the local rows of this image are complete
    const int seq = s_seq;
    if (tid < a.num_peers && a.peer_ack && !(a.peer_debug & 4)) {
      // bounded (~2 s): a consumer that never calls ypb_peer_wait must not hang the GPU - the entry is then overwritten and
      // the overrun is recorded in peer_state[3] for the host to see
      const volatile int32_t* ack = a.peer_ack + tid;
      int spins = 0;
      while (*ack - (seq - a.peer_depth) < 0) {
        __nanosleep(128);
        if (++spins > (1 << 24)) { a.peer_state[3] = seq; break; }
      }
    }
    __syncthreads();
    const int nfl = kept_n * cols;
    const long long img_off = static_cast<long long>(b) * a.max_det * cols;
    const long long entry = static_cast<long long>(seq % a.peer_depth) * a.peer_entry_stride;
    const float* src = a.out_rows + img_off;
    for (int p = 0; p < a.num_peers; ++p) {
      if ((a.peer_debug & 1) && p != a.my_rank) continue;  // diagnostic: local ring only
      float* dst = a.peer_rows[p] + entry + img_off;
      if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
        for (int i = tid; i < (nfl >> 2); i += nthr) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
        for (int i = (nfl & ~3) + tid; i < nfl; i += nthr) dst[i] = src[i];
      } else {
        for (int i = tid; i < nfl; i += nthr) dst[i] = src[i];
      }
      if (tid == 0) reinterpret_cast<int32_t*>(reinterpret_cast<float*>(a.peer_count[p]) + entry)[b] = kept_n;
    }
    // every thread's remote stores are ordered before the barrier; ONE system-scope fence by the thread that then
    // publishes (fences are cumulative), instead of 512 fences each waiting for its own remote acknowledgements
    __syncthreads();
    if (tid == 0) {
      if (!(a.peer_debug & 2)) __threadfence_system();
      const int prev = atomicAdd(&a.peer_state[0], 1);
      if (prev == a.batch - 1) {  // last image of the launch
        a.peer_state[0] = 0;
        a.peer_state[1] = seq;
        __threadfence_system();
        for (int p = 0; p < a.num_peers; ++p) *reinterpret_cast<volatile int32_t*>(a.peer_flag[p] + a.my_rank) = seq;
      }
    }
}

// The push as a kernel of its own (YPB_PEER_PUSH_SPLIT=1): light CTAs (no shared memory) wait for the remote stores'
// acknowledgement instead of the suppression CTAs with their 170 KB of shared memory.
__global__ void __launch_bounds__(128) peer_push_kernel(const __grid_constant__ SuppressArgs a) {
  peer_push(a, blockIdx.x, min(a.out_count[blockIdx.x], a.max_det), 128);
}

// ---------------------------------------------------------------------------------------------------------------
// stage 3: gather the kept rows (nms.py:159-161), riders, optional rescale, zero padding and the one-sided peer push.
// kk = kept keys in rank order (shared or global memory), kept_n <= max_det.  Called by every thread of ONE CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int RULE>
__device__ void gather_stage(const SuppressArgs& a, int b, const uint64_t* kk, int kept_n) {
  const int tid = threadIdx.x;
  const int cbits = a.cls_bits;
  const uint32_t cmask = (1u << cbits) - 1u;
  const float4* cand_box = a.cand_box + static_cast<long long>(b) * a.anchors;
  if (tid == 0) {
    a.out_count[b] = kept_n;
    if (a.out_count_host) a.out_count_host[b] = kept_n;  // posted write over PCIe; host-visible once the stream is synchronised
  }
  const int cols = 6 + a.extra;
  const float* cand_ang3 = a.cand_ang ? a.cand_ang + static_cast<long long>(b) * a.anchors : nullptr;
  for (int k = tid; k < kept_n; k += NT) {
    const uint64_t key = kk[k];
    const uint32_t row = key_row(key);
    const uint32_t anchor = row >> cbits, cls = row & cmask;
    if (a.out_idx) a.out_idx[static_cast<long long>(b) * a.max_det + k] = a.idx_as_row ? row : anchor;
    if (a.out_rows) {
      float* o = a.out_rows + (static_cast<long long>(b) * a.max_det + k) * cols;
      float4 bx = cand_box[anchor];
      o[4] = key_score(key);
      o[5] = static_cast<float>(cls);
      if (a.pred) {
        const long long base = static_cast<long long>(b) * a.pred_sb + static_cast<long long>(anchor) * a.pred_sa;
        for (int e = 0; e < a.extra; ++e)
          o[6 + e] = load_pred(a.pred, a.pred_dtype, base + static_cast<long long>(4 + a.nc + e) * a.pred_sc);
      } else if (a.extra == 1 && cand_ang3) {
        o[6] = cand_ang3[anchor];
      }
      // (riders of the fused path are gathered below, one (row, channel) pair per thread)
      if (a.scale_xforms) {
        // detect/predict.py:120 (scale_boxes) or obb/predict.py:59-60 (regularize_rboxes + scale_boxes(xywh=True)),
        // fused into the gather: the angle is the row's last column (nms.py:146)
        const ypb_scale_xform xf = a.scale_xforms[b];
        if constexpr (RULE == YPB_NMS_FAST_PROBIOU)
          scale_box(bx.x, bx.y, bx.z, bx.w, o + cols - 1, xf, YPB_BOXES_XYWHR, a.scale_padding);
        else
          scale_box(bx.x, bx.y, bx.z, bx.w, nullptr, xf, YPB_BOXES_XYXY, a.scale_padding);
      }
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
    }
  }
  if (a.pad_zero && a.out_rows) {  // exporter.py:1478-1479: rows past the kept count are zeros
    float* o = a.out_rows + static_cast<long long>(b) * a.max_det * cols;
    for (int i = kept_n * cols + tid; i < a.max_det * cols; i += NT) o[i] = 0.f;
    if (a.out_idx)  // and their indices are -1 (no anchor): a list a later call may take as its anchor subset
      for (int i = kept_n + tid; i < a.max_det; i += NT) a.out_idx[static_cast<long long>(b) * a.max_det + i] = -1;
  }
  if (a.rider && a.out_rows) {
    // Segment / Pose riders (head.py:837 mask coefficients, head.py:1252 decoded keypoints): only the kept anchors'
    // values are ever read.  Keypoints are decoded here (head.py:1254-1273) and, when the gather rescales to the original
    // image, scaled like pose/predict.py:73-75 (scale_coords).
    const int total = kept_n * a.extra;
    for (int t = tid; t < total; t += NT) {
      const int k = t / a.extra, e = t - k * a.extra;
      const uint32_t anchor = key_row(kk[k]) >> cbits;
      float v = load_pred(a.rider, a.rider_dtype, static_cast<long long>(b) * a.rider_sb + static_cast<long long>(e) * a.rider_sc + anchor);
      if (a.rider_kind == YPB_RIDER_KEYPOINTS) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < YPB_MAX_LEVELS; ++i)
          if (i < a.lv_n && static_cast<int>(anchor) >= a.lv_start[i]) l = i;
        const int loc = static_cast<int>(anchor) - a.lv_start[l];
        const int gy = loc / a.lv_w[l], gx = loc - gy * a.lv_w[l];
        const int d = e % a.rider_ndim;
        const float fx = static_cast<float>(gx) + 0.5f, fy = static_cast<float>(gy) + 0.5f;
        if (a.rider_dtype == YPB_F32) v = kpt_value<YPB_F32>(v, d, fx, fy, a.lv_stride[l]);
        else if (a.rider_dtype == YPB_F16) v = kpt_value<YPB_F16>(v, d, DType<YPB_F16>::rnd(fx), DType<YPB_F16>::rnd(fy), a.lv_stride[l]);
        else v = kpt_value<YPB_BF16>(v, d, DType<YPB_BF16>::rnd(fx), DType<YPB_BF16>::rnd(fy), a.lv_stride[l]);
        if (a.scale_xforms) v = scale_coord(v, d, a.scale_xforms[b], a.scale_padding);
      }
      a.out_rows[(static_cast<long long>(b) * a.max_det + k) * cols + 6 + e] = v;
    }
  }
  if (a.num_peers > 0 && a.out_rows) peer_push(a, b, kept_n, NT);
}

template <int RULE>
__device__ int suppress_image(Smem& sm, const SuppressArgs& a, int b, int n, const uint64_t** kk_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;

  uint64_t* ka = a.keys_a + static_cast<long long>(b) * a.rows_cap;
  uint64_t* kb = a.keys_b + static_cast<long long>(b) * a.rows_cap;

  const float4* cand_box = a.cand_box + static_cast<long long>(b) * a.anchors;
  const float thr = a.iou_thr;
  const GreedyThr gthr = make_greedy_thr(thr);
  const int cbits = a.cls_bits;
  const uint32_t cmask = (1u << cbits) - 1u;
  int kept_n = 0;
  const uint64_t* kk = nullptr;  // kept keys in rank order (stage 3 input)

  // More rows than shared memory ranks at once (val mode: conf 0.001, multi_label - tens of thousands of rows): the walk
  // stops at max_det kept rows (nms.py:157), which it normally reaches within the first few thousand ranks.  So first take a
  // PREFIX of the ranking - the nn best keys, SORT_SMEM_MAX/2 <= nn <= SORT_SMEM_MAX, found by a radix select that never sorts
  // the rest - and run the ordinary shared-memory paths on it.  The kept rows of a prefix are the kept rows of the full
  // walk restricted to those ranks (a row is only ever suppressed by higher-ranked rows), so if the prefix yields max_det
  // kept rows - or was the whole ranking - the result is exact; otherwise the full sort + walk below runs after all.
  uint64_t* kin = ka;
  int nn = n;
  bool partial = false;
  if constexpr (RULE == YPB_NMS_GREEDY) {
    if (n > SORT_SMEM_MAX && a.max_det <= SORT_SMEM_MAX / 4) {  // a prefix can only pay off if max_det rows fit well inside it
      const int L = min(n, a.max_nms);
      nn = select_prefix(ka, n, min(L, SORT_SMEM_MAX / 2), min(L, SORT_SMEM_MAX), kb, sm);
      kin = kb;
      partial = nn < L;
    }
  }
  if constexpr (RULE == YPB_NMS_GREEDY) {
    // A threshold >= 1 suppresses nothing (inter / union never exceeds 1): the kept rows are simply the first ranks.  The end2end
    // top-k (head.py:193-214) runs through here twice; the greedy walk over rows that cannot touch each other was 60 % of it.
    if (thr >= 1.0f && nn <= SORT_SMEM_MAX && a.box_div == 0.f) {
      bitonic_sort(sm.u.keys, kin, nn, KeyIdentity{});
      __syncthreads();
      *kk_out = sm.u.keys;
      return min(min(nn, a.max_nms), a.max_det);
    }
  }
  while (true) {
  kept_n = 0;
  // ---- fast path: class-wise walk (class-aware greedy rule, everything fits shared memory) ----------------------------
  bool done = false;
  if constexpr (RULE == YPB_NMS_GREEDY) {
    if (a.max_wh > 0.f && a.box_div == 0.f && nn <= SORT_SMEM_MAX && nn <= a.max_nms && (1 << a.cls_bits) <= CW_CLASS_MAX &&
        a.anchor_bits <= 20 && nn > 0) {
      const int r = classwise_greedy(sm, a, kin, nn, cand_box, gthr);
      if (r >= 0) { kept_n = r; kk = reinterpret_cast<const uint64_t*>(sm.cbox); done = true; }
    }
  }
  YPB_MARK(1);

  if (!done) {
  // ---- stage 1 ---------------------------------------------------------------------------------------------------
  const uint64_t* sorted;
  if (nn <= SORT_SMEM_MAX) {
    bitonic_sort(sm.u.keys, kin, nn, KeyIdentity{});
    sorted = sm.u.keys;
  } else {
    sorted = radix_sort_global(kin, kb, nn, sm);  // only reached with kin == ka
  }
  const int m = min(nn, a.max_nms);

  // ---- stage 2 ---------------------------------------------------------------------------------------------------
  const float* cand_ang = a.cand_ang ? a.cand_ang + static_cast<long long>(b) * a.anchors : nullptr;
  float4* kept_box = a.kept_box + static_cast<long long>(b) * a.max_det;
  float* kept_area = a.kept_area + static_cast<long long>(b) * a.max_det;
  uint64_t* kept_key = a.kept_key + static_cast<long long>(b) * a.max_det;
  float* rec_g = RULE == YPB_NMS_GREEDY ? nullptr : a.rec + static_cast<long long>(b) * min(a.rows_cap, a.max_nms) * 8;

  // thread (t, q): rank t of the chunk, part q.  A warp holds 32 consecutive ranks of ONE part, so the row it tests
  // against (kept row k, or chunk rank i) is the same for all lanes: shared-memory broadcast loads, no divergence.
  const int t = tid & (CH - 1);
  const int q = tid / CH;
  float4* kbox = a.max_det <= KEPT_SMEM ? sm.kbox : kept_box;
  float* karea = a.max_det <= KEPT_SMEM ? sm.karea : kept_area;

  // rank r -> (key, un-offset box); fetched one chunk ahead so the dependent global loads overlap the current chunk
  auto fetch = [&](int r, uint64_t& key, float4& bx, float& ang) {
    key = 0; bx = make_float4(0.f, 0.f, 0.f, 0.f); ang = 0.f;
    if (r < m) {
      key = sorted[r];
      const uint32_t anchor = key_row(key) >> cbits;
      bx = cand_box[anchor];
      if constexpr (RULE == YPB_NMS_FAST_PROBIOU) ang = cand_ang[anchor];
    }
  };
  uint64_t key_n; float4 bx_n; float ang_n;
  fetch(t, key_n, bx_n, ang_n);

  for (int c0 = 0; c0 < m && kept_n < a.max_det; c0 += CH) {
    const int r = c0 + t;
    const bool valid = r < m;
    const uint64_t key = key_n;
    const float4 bx = bx_n;
    const float ang = ang_n;
    fetch(r + CH, key_n, bx_n, ang_n);
    float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
    float area = 0.f;
    ObbRec me{};
    if (valid) {
      const uint32_t row = key_row(key);
      const uint32_t cls = row & cmask;
      const float off = __fmul_rn(static_cast<float>(cls), a.max_wh);  // nms.py:143
      if constexpr (RULE == YPB_NMS_FAST_PROBIOU) {
        if (q == 0) me = obb_record(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), bx.z, bx.w, ang);  // nms.py:146
      } else {
        float4 nb = bx;
        if (a.box_div > 0.f)  // exporter.py:1444: multiplier * (box / max(imgsz))
          nb = make_float4(__fmul_rn(a.box_mult, __fdiv_rn(bx.x, a.box_div)), __fmul_rn(a.box_mult, __fdiv_rn(bx.y, a.box_div)),
                           __fmul_rn(a.box_mult, __fdiv_rn(bx.z, a.box_div)), __fmul_rn(a.box_mult, __fdiv_rn(bx.w, a.box_div)));
        ob = make_float4(__fadd_rn(nb.x, off), __fadd_rn(nb.y, off), __fadd_rn(nb.z, off), __fadd_rn(nb.w, off));  // nms.py:149
        area = box_area(ob);
      }
    }
    if (q == 0) sm.dead[t] = valid ? 0 : 1;
    bool alive;

    if constexpr (RULE == YPB_NMS_GREEDY) {
      if (c0 == CH) YPB_MARK(8);
      if (q == 0) { sm.c.g.box[t] = ob; sm.c.g.area[t] = area; }
      __syncthreads();
      if (c0 == CH) YPB_MARK(9);
      // (a) against the rows kept in earlier chunks: part q takes every PARTS-th kept row (no early exit: ILP)
      bool hit = false;
      if (valid) {
        unsigned hs = 0, hm = 0;
#pragma unroll 4
        for (int k = q; k < kept_n; k += PARTS) {
          unsigned su, mb;
          greedy_flags(pair_geom(kbox[k], karea[k], ob, area), gthr, su, mb);
          hs |= su;
          hm |= mb;
        }
        hit = hs != 0u;
        if (!hit && hm) {  // some borderline pair and no sure hit: settle exactly (rare)
          for (int k = q; k < kept_n && !hit; k += PARTS) hit = greedy_suppresses(kbox[k], karea[k], ob, area, gthr);
        }
      }
      if (c0 == CH) YPB_MARK(10);
      if (hit) sm.dead[t] = 1;
      __syncthreads();
      if (c0 == CH) YPB_MARK(11);
      alive = sm.dead[t] == 0;
      if (q == 0) {
        const unsigned ab = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) sm.alive_bits[warp] = ab;
      }
      // (b) word q of the suppression row of rank t: which LOWER ranks 32q..32q+31 of this chunk t would suppress if it
      //     is kept (IoU is symmetric, so this is the transpose of "who suppresses t").  32 branch-free fp32 tests; the
      //     rare borderline pairs are settled exactly afterwards.
      const int tw = t >> 5;
      if (alive && q >= tw) {
        uint32_t word = 0, maybe = 0;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const int idx = q * 32 + i;
          unsigned su, mb;
          greedy_flags(pair_geom(ob, area, sm.c.g.box[idx], sm.c.g.area[idx]), gthr, su, mb);
          word |= su << i;
          maybe |= mb << i;
        }
        while (maybe) {
          const int i = __ffs(maybe) - 1;
          maybe &= maybe - 1;
          const int idx = q * 32 + i;
          const PairGeom g = pair_geom(ob, area, sm.c.g.box[idx], sm.c.g.area[idx]);
          if (greedy_exact(g.inter, g.uni, gthr.mid, gthr.tie_up)) word |= 1u << i;
        }
        if (q == tw) word &= ~(lt_mask | (1u << lane));  // only ranks below t
        sm.mask[t * CW + q] = word;
      }
      if (c0 == CH) YPB_MARK(12);
      __syncthreads();
      if (c0 == CH) YPB_MARK(13);
      // (c) resolve the chunk: the lowest rank still standing has every higher-ranked row decided, so it is kept and
      //     its row strikes the ranks it suppresses - one step per KEPT rank.  Every warp runs the same scalar loop on
      //     broadcast shared-memory reads, so nobody waits for the answer (no barrier, no warp collective).
      uint32_t rem[CW], kept[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) { rem[j] = sm.alive_bits[j]; kept[j] = 0; }
#pragma unroll
      for (int w = 0; w < CW; ++w) {
        while (rem[w]) {
          const int i = __ffs(rem[w]) - 1;
          kept[w] |= 1u << i;
          rem[w] &= ~(1u << i);
          const uint4 row = *reinterpret_cast<const uint4*>(&sm.mask[(w * 32 + i) * CW]);
          const uint32_t rw[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
          for (int j = w; j < CW; ++j) rem[j] &= ~rw[j];
        }
      }
      if (c0 == CH) YPB_MARK(14);
      alive = q == 0 && ((kept[0] >> lane) & 1u) != 0;
#pragma unroll
      for (int j = 1; j < CW; ++j) alive = (q == 0 && tw == j) ? ((kept[j] >> lane) & 1u) != 0 : alive;
      if (tid < CW) {
        uint32_t kw = kept[0];
#pragma unroll
        for (int j = 1; j < CW; ++j) kw = tid == j ? kept[j] : kw;
        sm.kept_bits[tid] = kw;
      }
      __syncthreads();
      if (c0 == CH) YPB_MARK(15);
    } else {
      // Fast-NMS: a rank is dropped iff ANY higher rank (kept or not) overlaps it >= thr, nms.py:221-223.
      float* my = sm.c.rec + t * 8;
      if (q == 0) {
        if constexpr (RULE == YPB_NMS_FAST_PROBIOU) {
          my[0] = me.x; my[1] = me.y; my[2] = me.a; my[3] = me.b; my[4] = me.c; my[5] = me.det;
        } else {
          my[0] = ob.x; my[1] = ob.y; my[2] = ob.z; my[3] = ob.w; my[4] = area; my[5] = 0.f;
        }
        if (valid) {
          float* gr = rec_g + static_cast<long long>(r) * 8;
#pragma unroll
          for (int k = 0; k < 6; ++k) gr[k] = my[k];
        }
      }
      __syncthreads();
      if constexpr (RULE == YPB_NMS_FAST_PROBIOU) me = ObbRec{my[0], my[1], my[2], my[3], my[4], my[5]};
      bool hit = false;
      if (valid) {
        for (int i = q; i < r; i += PARTS) {
          const float* o = i < c0 ? rec_g + static_cast<long long>(i) * 8 : sm.c.rec + (i - c0) * 8;
          if constexpr (RULE == YPB_NMS_FAST_PROBIOU) {
            ObbRec hi{o[0], o[1], o[2], o[3], o[4], o[5]};
            hit |= probiou_suppresses(hi, me, thr);
          } else {
            hit |= boxiou_suppresses(make_float4(o[0], o[1], o[2], o[3]), o[4], ob, area, thr);
          }
        }
      }
      if (hit) sm.dead[t] = 1;
      __syncthreads();
      alive = q == 0 && sm.dead[t] == 0;
      if (q == 0) {
        const unsigned kbits = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) sm.kept_bits[warp] = kbits;
      }
      __syncthreads();
    }

    // (d) append the chunk's kept rows in rank order (part 0 owns the ranks)
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CW; ++w) {
      const int pc = __popc(sm.kept_bits[w]);
      if (w < (t >> 5)) before += pc;
      total += pc;
    }
    if (q == 0 && alive) {
      const int pos = kept_n + before + __popc(sm.kept_bits[t >> 5] & lt_mask);
      if (pos < a.max_det) {
        kept_key[pos] = key;
        if constexpr (RULE == YPB_NMS_GREEDY) { kbox[pos] = ob; karea[pos] = area; }
      }
    }
    kept_n = min(a.max_det, kept_n + total);
    __syncthreads();  // kept_* visible to the block, smem reusable
    YPB_MARK(2 + c0 / CH);
  }

    __syncthreads();
    kk = kept_key;
  }  // !done
  if (!partial || kept_n >= a.max_det) break;  // uniform
  partial = false;  // the prefix did not yield max_det rows: rank and walk everything
  kin = ka;
  nn = n;
  __syncthreads();
  }  // attempts

  *kk_out = kk;
  return kept_n;
}

template <int RULE>
__device__ __noinline__ int legacy_fast_walk(Smem& sm, const SuppressArgs& a, int b, int n, const uint64_t** kk_out) {
  return suppress_image<RULE>(sm, a, b, n, kk_out);
}

template <int RULE>
__global__ void __launch_bounds__(NT, 1) sort_suppress_kernel(const __grid_constant__ SuppressArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  YPB_MARK(0);
  const int emitted = a.row_count[b];
  const int n = min(emitted, a.rows_cap);
  if (tid == 0 && a.out_cand) a.out_cand[b] = emitted;
  const uint64_t* kk = nullptr;
  const int kept_n = suppress_image<RULE>(sm, a, b, n, &kk);
  // ---- stage 3 ---------------------------------------------------------------------------------------------------
  YPB_MARK(30);
  gather_stage<RULE>(a, b, kk, kept_n);
  if (tid == 0) {  // clean on exit: this CTA was the only reader of its image's counter; the octet list was consumed by the previous kernel
    a.row_count[b] = 0;
    if (b == 0 && a.tile_counter) *a.tile_counter = 0;
  }
  __syncthreads();
  YPB_MARK(31);
}

// ---------------------------------------------------------------------------------------------------------------
// Fast-NMS over a thread-block CLUSTER (rotated ProbIoU / box_iou rules, nms.py:187-236)
//
// Fast-NMS has no sequential dependence: rank j is dropped iff ANY higher rank i < j (kept or not) overlaps it >= thr.
// A cluster of FC (1, 2, 4 or 8, chosen at launch so that the grid fills the GPU once) CTAs serves one image:
//   1. every CTA ranks the (<= SORT_SMEM_MAX) keys redundantly in its OWN shared memory (bitonic network) and builds the
//      per-rank records (covariance terms of metrics.py:187-203) as structure-of-arrays - no cross-CTA traffic in the hot loop;
//   2. the (target j, source i < j) tests are spread over all FC x NT threads: targets are dealt to the CTAs round-robin
//      (the cost of a target grows with its rank) and P threads share one target, each taking every P-th source;
//   3. a pair first takes a CHEAP EXACT-SAFE rejection (see fast_cheap_reject) - a handful of FMAs - and only the few
//      survivors are queued per warp and then evaluated 32 at a time with the full formula (3 IEEE divisions, 2 square
//      roots, log, exp), so the expensive path always runs with full lanes;
//   4. the CTAs OR their "dropped" bitmaps into CTA 0's shared memory through distributed shared memory
//      (cluster.map_shared_rank), one cluster barrier, and CTA 0 gathers the first max_det survivors in rank order.
// Images with more rows than shared memory holds take the single-CTA kernel's chunked walk on CTA 0.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FC_MAX = 8;       // CTAs per image at most (portable cluster size limit)
constexpr int FQ = 64;          // per-warp queue of pairs awaiting the full test

struct __align__(16) FastSmem {
  uint64_t keys[SORT_SMEM_MAX];   // sorted keys; afterwards the kept keys in rank order
  float f0[SORT_SMEM_MAX];        // probiou: x        box_iou: x1
  float f1[SORT_SMEM_MAX];        //          y                 y1
  float f2[SORT_SMEM_MAX];        //          trace T (< 0: never reject cheaply)   x2
  float f3[SORT_SMEM_MAX];        //          a                 y2
  float f4[SORT_SMEM_MAX];        //          b                 area
  float f5[SORT_SMEM_MAX];        //          c
  float f6[SORT_SMEM_MAX];        //          det
  uint32_t dead[SORT_SMEM_MAX / 32];
  uint32_t queue[NW][FQ];
  int wsum[NW];
};

union FastOrLegacy {
  FastSmem f;
  Smem legacy;
};

// Largest Bhattacharyya distance that still suppresses: 1 - sqrt(1 - exp(-bd) + eps) >= thr  <=>  bd <= -ln(1 + eps - (1-thr)^2).
// Cheap rejection of a pair of REGULAR boxes (finite, w, h > 0, aspect ratio <= 50, so every determinant below is
// well conditioned in fp32):  t1 + t2 = 0.25 * d^T (S1+S2)^-1 d >= 0.25 |d|^2 / trace(S1+S2)  (smallest eigenvalue of the
// inverse), t3 >= -1e-3, fp32 evaluation error of t1 + t2 < 1 % under the aspect bound - so
//     |d|^2 > K (T1 + T2),  K = (bd_max + 0.02) / 0.24   ==>   bd > bd_max   ==>   probiou < thr
// with margin; the pair is then skipped without evaluating metrics.py:251-284.  Irregular boxes (trace stored as -1) and
// thresholds below 1e-3 never take the shortcut.
__device__ __forceinline__ float probiou_reject_factor(float thr) {
  if (!(thr >= 1e-3f) || !(thr <= 1.0f)) return INFINITY;
  const double om = 1.0 - static_cast<double>(thr);
  const double arg = 1.0 + 1e-7 - om * om;
  const double bd_max = -log(arg);
  return static_cast<float>((bd_max + 0.02) / 0.24);
}

template <int RULE>
__global__ void __launch_bounds__(NT, 1) fast_nms_cluster_kernel(const __grid_constant__ SuppressArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastSmem& fs = reinterpret_cast<FastOrLegacy*>(smem_raw)->f;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = static_cast<int>(cluster.block_rank());
  const int FC = static_cast<int>(cluster.dim_blocks().x);  // CTAs serving this image
  const int b = blockIdx.x / FC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;

  const int emitted = a.row_count[b];
  const int n = min(emitted, a.rows_cap);
  if (crank == 0 && tid == 0 && a.out_cand) a.out_cand[b] = emitted;
  const int cbits = a.cls_bits;
  const uint32_t cmask = (1u << cbits) - 1u;
  const float thr = a.iou_thr;

  cluster.sync();  // every CTA of the cluster has read the image's counter: CTA 0 may clear it (clean on exit)
  if (crank == 0 && tid == 0) {
    a.row_count[b] = 0;
    if (b == 0 && a.tile_counter) *a.tile_counter = 0;
  }
  if (n > SORT_SMEM_MAX) {
    // uniform over the cluster: CTA 0 runs the single-CTA path (global radix sort + chunked walk), the others leave
    if (crank != 0) return;
    Smem& sm = reinterpret_cast<FastOrLegacy*>(smem_raw)->legacy;
    const uint64_t* kk = nullptr;
    const int kept_n = legacy_fast_walk<RULE>(sm, a, b, n, &kk);
    gather_stage<RULE>(a, b, kk, kept_n);
    return;
  }

  uint64_t* ka = a.keys_a + static_cast<long long>(b) * a.rows_cap;
  const float4* cand_box = a.cand_box + static_cast<long long>(b) * a.anchors;
  const float* cand_ang = a.cand_ang ? a.cand_ang + static_cast<long long>(b) * a.anchors : nullptr;

  // ---- 1. rank + records (every CTA, redundantly) ------------------------------------------------------------------------
  for (int w = tid; w < SORT_SMEM_MAX / 32; w += NT) fs.dead[w] = 0u;
  bitonic_sort(fs.keys, ka, n, KeyIdentity{});
  const int m = min(n, a.max_nms);
  for (int r = tid; r < m; r += NT) {
    const uint32_t row = key_row(fs.keys[r]);
    const uint32_t anchor = row >> cbits, cls = row & cmask;
    const float4 bx = cand_box[anchor];
    const float off = __fmul_rn(static_cast<float>(cls), a.max_wh);  // nms.py:143
    if constexpr (RULE == YPB_NMS_FAST_PROBIOU) {
      const float ang = cand_ang[anchor];
      const ObbRec o = obb_record(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), bx.z, bx.w, ang);  // nms.py:146
      const float w = bx.z, h = bx.w;
      const bool regular = isfinite(o.x) && isfinite(o.y) && isfinite(ang) && w > 0.f && h > 0.f && w <= 1e5f && h <= 1e5f &&
                           fmaxf(w, h) <= 50.f * fminf(w, h) && fabsf(o.x) <= 1e8f && fabsf(o.y) <= 1e8f;
      fs.f0[r] = o.x; fs.f1[r] = o.y;
      fs.f2[r] = regular ? (w * w + h * h) * (1.0f / 12.0f) : -1.f;
      fs.f3[r] = o.a; fs.f4[r] = o.b; fs.f5[r] = o.c; fs.f6[r] = o.det;
    } else {
      const float4 ob = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
      fs.f0[r] = ob.x; fs.f1[r] = ob.y; fs.f2[r] = ob.z; fs.f3[r] = ob.w; fs.f4[r] = box_area(ob);
    }
  }
  cluster.sync();  // records visible in this CTA; every CTA of the cluster is running and has cleared its bitmap

  // ---- 2. pair tests ----------------------------------------------------------------------------------------------------
  const float K = RULE == YPB_NMS_FAST_PROBIOU ? probiou_reject_factor(thr) : 0.f;
  int P = 1;                                    // threads per target: as many as the cluster has to spare
  while (P < 32 && 2 * P * m <= FC * NT) P <<= 1;
  const int q = tid & (P - 1);
  const int slots = NT / P;                     // targets per CTA per pass
  uint32_t* queue = fs.queue[warp];
  int qn = 0;                                   // warp-uniform
  auto full_test = [&](uint32_t packed) {
    const int j = packed >> 16, i = packed & 0xffffu;
    bool hit;
    if constexpr (RULE == YPB_NMS_FAST_PROBIOU) {
      const ObbRec hi{fs.f0[i], fs.f1[i], fs.f3[i], fs.f4[i], fs.f5[i], fs.f6[i]};
      const ObbRec me{fs.f0[j], fs.f1[j], fs.f3[j], fs.f4[j], fs.f5[j], fs.f6[j]};
      hit = probiou_suppresses(hi, me, thr);
    } else {
      hit = boxiou_suppresses(make_float4(fs.f0[i], fs.f1[i], fs.f2[i], fs.f3[i]), fs.f4[i],
                              make_float4(fs.f0[j], fs.f1[j], fs.f2[j], fs.f3[j]), fs.f4[j], thr);
    }
    if (hit) atomicOr(&fs.dead[j >> 5], 1u << (j & 31));
  };
  for (int base = 0; base < m; base += slots * FC) {
    const int j = base + (tid / P) * FC + crank;
    const bool valid = j < m;
    float jx = 0.f, jy = 0.f, jt = -1.f, jx2 = 0.f, jy2 = 0.f;
    if (valid) {
      jx = fs.f0[j]; jy = fs.f1[j]; jt = fs.f2[j];
      if constexpr (RULE != YPB_NMS_FAST_PROBIOU) { jx2 = fs.f2[j]; jy2 = fs.f3[j]; }
    }
    // longest chain of the warp (targets grow with the lane): uniform trip count, lanes past their own j idle
    const int jmax = __reduce_max_sync(0xffffffffu, valid ? j : 0);
    for (int i0 = 0; i0 < jmax; i0 += P) {  // uniform trip count: the warp votes inside
      const int i = i0 + q;
      bool near = false;
      if (valid && i < j) {
        if constexpr (RULE == YPB_NMS_FAST_PROBIOU) {
          const float dx = jx - fs.f0[i], dy = jy - fs.f1[i];
          const float ti = fs.f2[i];
          near = !(jt >= 0.f && ti >= 0.f && dx * dx + dy * dy > K * (jt + ti));
        } else {
          // box_iou: disjoint boxes have inter == 0 exactly -> iou == 0 < thr (thr > 0); NaNs fall through to the full test
          near = !(fs.f0[i] >= jx2 || fs.f2[i] <= jx || fs.f1[i] >= jy2 || fs.f3[i] <= jy) || !(thr > 0.f);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, near);
      if (bal) {
        if (near) queue[qn + __popc(bal & lt_mask)] = (static_cast<uint32_t>(j) << 16) | static_cast<uint32_t>(i);
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 32) {
          qn -= 32;
          full_test(queue[qn + lane]);
          __syncwarp();
        }
      }
    }
  }
  if (lane < qn) full_test(queue[lane]);
  __syncthreads();

  // ---- 3. merge the bitmaps in CTA 0 (distributed shared memory) -----------------------------------------------------------
  if (crank != 0) {
    uint32_t* remote = cluster.map_shared_rank(fs.dead, 0);
    for (int w = tid; w < (m + 31) / 32; w += NT) {
      const uint32_t v = fs.dead[w];
      if (v) atomicOr(remote + w, v);
    }
  }
  cluster.sync();
  if (crank != 0) return;

  // ---- 4. first max_det survivors in rank order, then the gather ----------------------------------------------------------
  const int nwords = (m + 31) / 32;
  int kept_total = 0;
  for (int w0 = 0; w0 < nwords && kept_total < a.max_det; w0 += NT) {
    const int w = w0 + tid;
    uint32_t alive = 0u;
    if (w < nwords) {
      alive = ~fs.dead[w];
      if (w == nwords - 1 && (m & 31)) alive &= (1u << (m & 31)) - 1u;
    }
    const int c = __popc(alive);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) fs.wsum[warp] = inc;
    __syncthreads();
    int before = kept_total + inc - c, total = 0;
    for (int x = 0; x < NW; ++x) { const int v = fs.wsum[x]; if (x < warp) before += v; total += v; }
    // the kept keys overwrite the records' f0/f1 arrays (dead after the tests): max_det <= SORT_SMEM_MAX keys of 8 bytes
    uint64_t* kkw = reinterpret_cast<uint64_t*>(fs.f0);
    uint32_t rem = alive;
    int pos = before;
    while (rem && pos < a.max_det) {
      const int bit = __ffs(rem) - 1;
      rem &= rem - 1;
      kkw[pos++] = fs.keys[w * 32 + bit];
    }
    kept_total += total;
    __syncthreads();
  }
  const int kept_n = min(kept_total, a.max_det);
  gather_stage<RULE>(a, b, reinterpret_cast<const uint64_t*>(fs.f0), kept_n);
}

struct PeerAckPtrs { int32_t* p[YPB_MAX_PEERS]; };
__global__ void peer_wait_kernel_byval(const int32_t* flags, int world, int32_t* state, int lag, int depth, PeerAckPtrs acks,
                                       int has_ack, int my_rank, long long* slot_index) {
  const int want = lag < 0 ? state[2] + 1 : state[1] - lag;  // lag < 0: the batch after the one handed out last (in order)
  const int done = state[2];
  if (has_ack && threadIdx.x < world) *reinterpret_cast<volatile int32_t*>(acks.p[threadIdx.x] + my_rank) = done;
  if (threadIdx.x == 0) {
    if (want > 0) {
      for (int r = 0; r < world; ++r) {
        const volatile int32_t* f = flags + r;
        int spins = 0;
        while (*f - want < 0) {  // bounded (~4 s): a peer that died must not hang this GPU; recorded in state[3] (negative)
          __nanosleep(256);
          if (++spins > (1 << 24)) { state[3] = -want; break; }
        }
      }
    }
    __threadfence_system();
    state[2] = want > done ? want : done;
    if (slot_index) *slot_index = want > 0 ? want % depth : 0;
  }
}

// ypb_peer_wait with a consumer attached: after the wait, all threads copy the returned ring entry (world x slot floats)
// into `out` - one kernel instead of wait + gather, for consumers that just want the gathered batch in a stable buffer.
__global__ void __launch_bounds__(1024)
peer_wait_copy_kernel(const int32_t* flags, int world, int32_t* state, int lag, int depth, PeerAckPtrs acks, int has_ack,
                      int my_rank, long long* slot_index, const float* ring, long long entry_floats, float* out) {
  __shared__ int s_slot;
  if (threadIdx.x < 32) {
    const int want = lag < 0 ? state[2] + 1 : state[1] - lag;  // lag < 0: the batch after the one handed out last (in order)
    const int done = state[2];
    if (has_ack && threadIdx.x < world) *reinterpret_cast<volatile int32_t*>(acks.p[threadIdx.x] + my_rank) = done;
    if (threadIdx.x == 0) {
      if (want > 0) {
        for (int r = 0; r < world; ++r) {
          const volatile int32_t* f = flags + r;
          int spins = 0;
          while (*f - want < 0) {
            __nanosleep(256);
            if (++spins > (1 << 24)) { state[3] = -want; break; }
          }
        }
      }
      __threadfence_system();
      state[2] = want > done ? want : done;
      s_slot = want > 0 ? want % depth : 0;
      if (slot_index) *slot_index = s_slot;
    }
  }
  __syncthreads();
  const float4* src = reinterpret_cast<const float4*>(ring + static_cast<long long>(s_slot) * entry_floats);
  float4* dst = reinterpret_cast<float4*>(out);
  for (long long i = threadIdx.x; i < entry_floats / 4; i += blockDim.x) dst[i] = src[i];
}

cudaError_t launch_peer_wait_copy(const int32_t* flags, int world, int32_t* state, int lag, int depth, int32_t* const* peer_ack_host,
                                  int my_rank, long long* slot_index, const float* ring, long long entry_floats, float* out,
                                  cudaStream_t st) {
  PeerAckPtrs acks{};
  if (peer_ack_host)
    for (int i = 0; i < world && i < YPB_MAX_PEERS; ++i) acks.p[i] = peer_ack_host[i];
  peer_wait_copy_kernel<<<1, 1024, 0, st>>>(flags, world, state, lag, depth > 0 ? depth : 1, acks, peer_ack_host ? 1 : 0, my_rank,
                                            slot_index, ring, entry_floats, out);
  return cudaGetLastError();
}

cudaError_t launch_peer_wait(const int32_t* flags, int world, int32_t* state, int lag, int depth, int32_t* const* peer_ack_host,
                             int my_rank, long long* slot_index, cudaStream_t st) {
  PeerAckPtrs acks{};
  if (peer_ack_host)
    for (int i = 0; i < world && i < YPB_MAX_PEERS; ++i) acks.p[i] = peer_ack_host[i];
  peer_wait_kernel_byval<<<1, 32, 0, st>>>(flags, world, state, lag, depth > 0 ? depth : 1, acks, peer_ack_host ? 1 : 0, my_rank, slot_index);
  return cudaGetLastError();
}

__global__ void boxes_prep_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n, int box_dim,
                                  uint64_t* keys, float4* cand_box, float* cand_ang, int32_t* row_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) row_count[0] = n;
  if (i >= n) return;
  keys[i] = make_key(scores[i], static_cast<uint32_t>(i));
  const float* p = boxes + static_cast<long long>(i) * box_dim;
  cand_box[i] = make_float4(p[0], p[1], p[2], p[3]);
  if (box_dim == 5) cand_ang[i] = p[4];
}

// nms.py:159-161 returns one tensor per image.  The suppression kernel writes fixed-stride (B, max_det, cols) rows; this
// kernel packs the kept rows of all images back to back (image order) so the host can cut the per-image views with ONE
// split instead of B slicing calls: CTA b sums count[0..b) and copies its image's rows / anchor indices.
__global__ void __launch_bounds__(256)
compact_results_kernel(const float* __restrict__ rows, const long long* __restrict__ idx, const int32_t* __restrict__ count,
                       int batch, int max_det, int cols, float* __restrict__ out_rows, long long* __restrict__ out_idx,
                       int32_t* __restrict__ out_offsets) {
  __shared__ int warp_part[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  int part = 0;
  for (int i = tid; i < b; i += 256) part += min(count[i], max_det);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) warp_part[tid >> 5] = part;
  __syncthreads();
  int prefix = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) prefix += warp_part[w];
  const int n = min(count[b], max_det);
  if (out_offsets && tid == 0) {
    out_offsets[b] = prefix;
    if (b == batch - 1) out_offsets[batch] = prefix + n;
  }
  if (out_rows) {
    const float* src = rows + static_cast<long long>(b) * max_det * cols;
    float* dst = out_rows + static_cast<long long>(prefix) * cols;
    const int nfl = n * cols;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
      for (int i = tid; i < (nfl >> 2); i += 256) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
      for (int i = (nfl & ~3) + tid; i < nfl; i += 256) dst[i] = src[i];
    } else {
      for (int i = tid; i < nfl; i += 256) dst[i] = src[i];
    }
  }
  if (out_idx && idx) {
    const long long* src = idx + static_cast<long long>(b) * max_det;
    for (int i = tid; i < n; i += 256) out_idx[prefix + i] = src[i];
  }
}

cudaError_t launch_compact_results(const float* rows, const long long* idx, const int32_t* count, int batch, int max_det,
                                   int cols, float* out_rows, long long* out_idx, int32_t* out_offsets, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  compact_results_kernel<<<batch, 256, 0, st>>>(rows, idx, count, batch, max_det, cols, out_rows, out_idx, out_offsets);
  return cudaGetLastError();
}

// metrics.py:54-75 box_iou / metrics.py:251-284 batch_probiou as free functions: out[i * m + j] = iou(a_i, b_j).
__global__ void __launch_bounds__(256)
pairwise_iou_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m, int box_dim, float* __restrict__ out) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(n) * m) return;
  const int i = static_cast<int>(t / m), j = static_cast<int>(t - static_cast<long long>(i) * m);
  const float* p = a + static_cast<long long>(i) * box_dim;
  const float* q = b + static_cast<long long>(j) * box_dim;
  if (box_dim == 5) {
    out[t] = probiou_value(obb_record(p[0], p[1], p[2], p[3], p[4]), obb_record(q[0], q[1], q[2], q[3], q[4]));
  } else {
    const float4 ba = make_float4(p[0], p[1], p[2], p[3]), bb = make_float4(q[0], q[1], q[2], q[3]);
    out[t] = boxiou_value(ba, box_area(ba), bb, box_area(bb));
  }
}

cudaError_t launch_pairwise_iou(const float* a, int n, const float* b, int m, int box_dim, float* out, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * m;
  if (total <= 0) return cudaSuccess;
  pairwise_iou_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(a, n, b, m, box_dim, out);
  return cudaGetLastError();
}

template <int RULE>
static cudaError_t launch_fast_cluster(const SuppressArgs& a, int dev, cudaStream_t st) {
  static std::atomic<bool> configured[64];
  const size_t smem = sizeof(FastOrLegacy);
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(fast_nms_cluster_kernel<RULE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  int fc = FC_MAX;  // the largest cluster that still lets every image start in the first wave of 148 SMs
  while (fc > 1 && static_cast<long long>(a.batch) * fc > 148) fc >>= 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(a.batch) * fc);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = fc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fast_nms_cluster_kernel<RULE>, a);
}

cudaError_t launch_sort_suppress(const SuppressArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(Smem);
  // opt in to > 48 KB dynamic shared memory once per (device, rule); not a stream operation, safe under graph capture
  static std::atomic<bool> configured[64][3];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (a.rule < 0 || a.rule > 2) return cudaErrorInvalidValue;
  if (a.batch <= 0) return cudaSuccess;
  SuppressArgs aa = a;
  aa.dbg = g_phase_buf.load(std::memory_order_relaxed);
  static const int peer_debug = [] { const char* e = std::getenv("YPB_PEER_DEBUG"); return e ? std::atoi(e) : 0; }();
  aa.peer_debug = peer_debug;
  static const bool split_push = [] { const char* e = std::getenv("YPB_PEER_PUSH_SPLIT"); return e && e[0] == '1'; }();
  const bool push_after = split_push && a.num_peers > 0 && a.out_rows;
  if (push_after) aa.num_peers = 0;  // the suppression kernel writes the local rows only; peer_push_kernel follows
  if (a.rule == YPB_NMS_FAST_PROBIOU || a.rule == YPB_NMS_FAST_BOXIOU) {
    aa.dbg = nullptr;
    e = a.rule == YPB_NMS_FAST_PROBIOU ? launch_fast_cluster<YPB_NMS_FAST_PROBIOU>(aa, dev, st)
                                       : launch_fast_cluster<YPB_NMS_FAST_BOXIOU>(aa, dev, st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && push_after) { peer_push_kernel<<<a.batch, 128, 0, st>>>(a); e = cudaGetLastError(); }
    return e;
  }
  const bool need_cfg = dev < 0 || dev >= 64 || !configured[dev][a.rule].load(std::memory_order_acquire);
  if (need_cfg) {
    e = cudaFuncSetAttribute(sort_suppress_kernel<YPB_NMS_GREEDY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev][a.rule].store(true, std::memory_order_release);
  }
  sort_suppress_kernel<YPB_NMS_GREEDY><<<a.batch, NT, smem, st>>>(aa);
  e = cudaGetLastError();
  if (e == cudaSuccess && push_after) { peer_push_kernel<<<a.batch, 128, 0, st>>>(a); e = cudaGetLastError(); }
  return e;
}

cudaError_t launch_boxes_prep(const float* boxes, const float* scores, int n, int box_dim, uint64_t* keys,
                              float4* cand_box, float* cand_ang, int32_t* row_count, cudaStream_t st) {
  const int blocks = n > 0 ? (n + 255) / 256 : 1;
  boxes_prep_kernel<<<blocks, 256, 0, st>>>(boxes, scores, n, box_dim, keys, cand_box, cand_ang, row_count);
  return cudaGetLastError();
}

}  // namespace ypb
