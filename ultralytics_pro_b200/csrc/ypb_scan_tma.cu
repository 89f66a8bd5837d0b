// Fused path, kernel 1 as a PERSISTENT, TMA-fed pipeline (sm_100a): the class scan + confidence filter + row compaction of
// nms.py:76-131 evaluated on head.py:169's scores without materialising them - same outputs as scan_classes_kernel
// (ypb_decode.cu), different data movement:
//
//   grid    one CTA per SM, each walking the tile list t = blockIdx.x, + gridDim.x, ...; a tile = TW consecutive anchors of one
//           level of one image x ALL nc class rows
//   loads   ONE cp.async.bulk.tensor.3d (TMA, tensor map over the level tensor (B, 64+nc, H*W)) per tile, issued by a
//           single producer thread into a ring of S shared-memory stages; completion is signalled on an mbarrier with the
//           byte count (expect_tx).  No load occupies a register: the bytes in flight per SM are S-1 stages (>= 80 KB),
//           independent of occupancy - which is what the register-staged LDG form runs out of for 16-bit heads
//   math    NCW consumer warps; a warp owns a whole tile (lane = VEC anchors, 16 B per row) and walks the nc rows in shared
//           memory (conflict-free 128-bit LDS).  Same per-element arithmetic as the LDG kernel: NaN-propagating max, first
//           argmax and runner-up (packed x2 for 16-bit inputs); ONE sigmoid per anchor; ties re-read from the stage
//   tail    per-warp epilogue (ballot compaction, one atomicAdd per warp-tile, keys + octet list) runs while the TMA engine
//           is already filling the other stages: the one-wave-grid epilogue bubble of the LDG kernel disappears
//
// Out-of-range anchors of a level's last tile are zero-filled by the TMA unit and masked per lane.
#include <cuda.h>

#include <cstdlib>

#include "ypb_common.cuh"

namespace ypb {

namespace {

constexpr int TMA_MAX_NCW = 16;      // consumer warps at most (+ 1 producer warp)
constexpr int TMA_MAX_STAGES = 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ bool class_ok(const uint32_t* mask, int c) {
  return mask == nullptr || ((mask[c >> 5] >> (c & 31)) & 1u);
}
__device__ __forceinline__ float max_nan_f(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

struct TmaScanGeom {
  int num_levels, batch, nc, tiles_per_image;
  int tile_start[YPB_MAX_LEVELS + 1];    // prefix of ceil(H*W / TW)
  int level_anchors[YPB_MAX_LEVELS];     // H*W
  int anchor_start[YPB_MAX_LEVELS + 1];
  int group_start[YPB_MAX_LEVELS + 1];   // prefix of H*W / VEC (the octet index space shared with decode_tiles_kernel)
  int G;                                 // groups reserved per image in that space
  int stages, ncw;                       // ring depth, consumer warps
};

struct TmaMaps { CUtensorMap m[YPB_MAX_LEVELS]; };

// LB = bytes per lane and class row (16, 8 or 4): a tile row is 32 * LB bytes, a lane owns VEC = LB / sizeof(T) anchors.  Narrower
// lanes mean smaller tiles, hence more stages and more consumer warps per SM for the same shared memory - the scan is a
// dependent max chain per anchor, so it wants several warps per scheduler to hide its own latency.
template <int DT_IN, bool MULTI, int LB>
__global__ void __launch_bounds__(32 * (TMA_MAX_NCW + 1), 1)
scan_classes_tma_kernel(const __grid_constant__ TmaMaps maps, const __grid_constant__ TmaScanGeom g, const __grid_constant__ FilterArgs f) {
  using TI = typename DType<DT_IN>::type;
  using DV = DType<DT_IN>;
  constexpr int TMA_ROW_BYTES = 32 * LB;
  constexpr int VEC = LB / static_cast<int>(sizeof(TI));            // anchors per lane
  constexpr int VEC_CANON = 16 / static_cast<int>(sizeof(TI));      // anchors per group of the octet index space (decode_tiles_kernel)
  constexpr int RG = VEC_CANON / VEC;                               // lanes per canonical group
  constexpr int TW = TMA_ROW_BYTES / static_cast<int>(sizeof(TI));
  constexpr int LPO = 8 / VEC;                                      // lanes per octet (8 anchors)
  constexpr int LPO_CANON = VEC_CANON >= 8 ? 1 : 8 / VEC_CANON;     // canonical groups per octet
  constexpr int NOCT = 32 / LPO;
  static_assert(VEC >= 1 && VEC <= 8 && RG >= 1 && LPO >= 1, "lane width");
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = g.stages, nc = g.nc, NCW = g.ncw;
  const int stage_bytes = nc * TMA_ROW_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(S) * stage_bytes);
  uint64_t* empty = full + TMA_MAX_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const long long total = static_cast<long long>(g.batch) * g.tiles_per_image;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int total32 = static_cast<int>(total);  // the launcher guarantees total < 2^31
  if (warp == NCW) {
    // ---- producer: one elected thread, one TMA per tile.  Stage and phase advance incrementally (no modulo per tile). -------
    if (lane == 0) {
      int s = 0, fill = 0;  // fill = how many times stage s has been filled before
      for (int t = blockIdx.x; t < total32; t += gridDim.x) {
        if (fill > 0) mbar_wait(&empty[s], (fill - 1) & 1);
        const int b = t / g.tiles_per_image, r = t - b * g.tiles_per_image;
        int l = 0;
#pragma unroll
        for (int i = 1; i < YPB_MAX_LEVELS; ++i)
          if (i < g.num_levels && r >= g.tile_start[i]) l = i;
        mbar_expect_tx(&full[s], static_cast<uint32_t>(stage_bytes));
        tma_load_3d(smem + static_cast<size_t>(s) * stage_bytes, &maps.m[l], (r - g.tile_start[l]) * TW, 64, b, &full[s]);
        if (++s == S) { s = 0; ++fill; }
      }
    }
    return;
  }

  // ---- consumers: warp w takes this CTA's tiles number w, w + NCW, ... (whole tiles); tile number `it` sits in stage it % S,
  //      filled for the (it / S)-th time - both tracked incrementally ------------------------------------------------------------
  const float conf = f.conf;
  int s = warp % S, fill = warp / S;
  const int s_step = NCW % S, fill_step = NCW / S;
  for (long long tt = blockIdx.x + static_cast<long long>(warp) * gridDim.x; tt < total; tt += static_cast<long long>(NCW) * gridDim.x) {
    const int t = static_cast<int>(tt);
    const int b = t / g.tiles_per_image, r = t - b * g.tiles_per_image;
    int l = 0;
#pragma unroll
    for (int i = 1; i < YPB_MAX_LEVELS; ++i)
      if (i < g.num_levels && r >= g.tile_start[i]) l = i;
    const int a_local = (r - g.tile_start[l]) * TW + lane * VEC;   // first anchor of this lane inside the level
    const bool in_level = a_local < g.level_anchors[l];            // VEC divides H*W: a lane is inside or outside as a whole
    const int a_glob = g.anchor_start[l] + a_local;
    mbar_wait(&full[s], fill & 1);
    const unsigned char* tile = smem + static_cast<size_t>(s) * stage_bytes + lane * LB;
    auto row = [&](int c) { return *reinterpret_cast<const Pack<TI, VEC>*>(tile + static_cast<size_t>(c) * TMA_ROW_BYTES); };

    int rows[VEC];
    float score[VEC];
    int cls[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { rows[i] = 0; score[i] = 0.f; cls[i] = 0; }

    if constexpr (MULTI) {
      float nanacc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) nanacc[i] = -INFINITY;
#pragma unroll 4
      for (int c = 0; c < nc; ++c) {
        const Pack<TI, VEC> p = row(c);
        const bool ok = class_ok(f.class_mask, c);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float v = DV::to_f(p.v[i]);
          nanacc[i] = max_nan_f(nanacc[i], v);
          rows[i] += (DV::rnd(sigmoid_f(v)) > conf && ok) ? 1 : 0;
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (nanacc[i] != nanacc[i] || !in_level) rows[i] = 0;  // amax -> NaN -> not a candidate (nms.py:76)
    } else {
      float m[VEC], m2[VEC];
      if constexpr (DT_IN != YPB_F32) {
        // 16-bit inputs: the whole scan stays in packed x2 arithmetic (comparisons and max/min of 16-bit floats are exact)
        using T2 = typename Packed2<DT_IN>::type;
        constexpr int NP = VEC / 2;
        T2 pm[NP], pm2[NP];
        uint32_t pidx[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) { pm[j] = Packed2<DT_IN>::neg_inf(); pm2[j] = pm[j]; pidx[j] = 0u; }
#pragma unroll 4
        for (int c = 0; c < nc; ++c) {
          const Pack<TI, VEC> p = row(c);
          const T2* v2 = reinterpret_cast<const T2*>(&p);
          const uint32_t cc = static_cast<uint32_t>(c) * 0x00010001u;
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const uint32_t gt = __hgt2_mask(v2[j], pm[j]);
            pm2[j] = __hmax2(pm2[j], __hmin2(pm[j], v2[j]));
            pm[j] = __hmax2_nan(pm[j], v2[j]);
            pidx[j] = (cc & gt) | (pidx[j] & ~gt);
          }
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          m[2 * j] = Packed2<DT_IN>::lo(pm[j]);   m[2 * j + 1] = Packed2<DT_IN>::hi(pm[j]);
          m2[2 * j] = Packed2<DT_IN>::lo(pm2[j]); m2[2 * j + 1] = Packed2<DT_IN>::hi(pm2[j]);
          cls[2 * j] = static_cast<int>(pidx[j] & 0xffffu); cls[2 * j + 1] = static_cast<int>(pidx[j] >> 16);
        }
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { m[i] = -INFINITY; m2[i] = -INFINITY; }
#pragma unroll 8
        for (int c = 0; c < nc; ++c) {
          const Pack<TI, VEC> p = row(c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float v = DV::to_f(p.v[i]);
            const bool gt = v > m[i];
            m2[i] = fmaxf(m2[i], fminf(m[i], v));
            m[i] = max_nan_f(m[i], v);
            cls[i] = gt ? c : cls[i];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float sc = DV::rnd(sigmoid_f(m[i]));
        if (in_level && sc > conf) {  // false for a NaN max
          if (DV::rnd(sigmoid_f(m2[i])) == sc) {
            // the runner-up rounds to the same score: take the FIRST class that reaches it (nms.py:120) - from the stage
            for (int c = 0; c < cls[i]; ++c) {
              const TI* p = reinterpret_cast<const TI*>(tile + static_cast<size_t>(c) * TMA_ROW_BYTES);
              if (DV::rnd(sigmoid_f(DV::to_f(p[i]))) == sc) { cls[i] = c; break; }
            }
          }
          score[i] = sc;
          rows[i] = class_ok(f.class_mask, cls[i]) ? 1 : 0;  // nms.py:127-131
        }
      }
    }

    if constexpr (!MULTI) {
      // single-label: everything the epilogue needs is in registers now - hand the stage back BEFORE the global atomics of
      // the epilogue (two round trips to L2), so the TMA engine refills it while this warp waits for them
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // ---- per-warp epilogue: compaction of the keys, octet work list for decode_tiles_kernel ------------------------------
    int my_rows = 0;
    uint32_t flags = 0;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { my_rows += rows[i]; flags |= rows[i] > 0 ? 1u << i : 0u; }
    const unsigned bal = __ballot_sync(0xffffffffu, flags != 0);
    if (bal) {  // warp-uniform
      uint32_t oct_mask = 0;
#pragma unroll
      for (int o = 0; o < NOCT; ++o)
        if ((bal >> (o * LPO)) & ((1u << LPO) - 1u)) oct_mask |= 1u << o;
      int inc = my_rows;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tt = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += tt;
      }
      const int total_rows = __shfl_sync(0xffffffffu, inc, 31);
      int base_rows = 0, base_oct = 0;
      if (lane == 0) {
        base_rows = atomicAdd(&f.row_count[b], total_rows);
        base_oct = atomicAdd(f.tile_count, __popc(oct_mask));
      }
      base_rows = __shfl_sync(0xffffffffu, base_rows, 0);
      base_oct = __shfl_sync(0xffffffffu, base_oct, 0);
      // octet ids index the dense space of CANONICAL anchor groups (16 bytes of a row: VEC_CANON anchors) shared with
      // decode_tiles_kernel: group index = b * G + group-in-image; an octet = LPO_CANON consecutive groups = LPO lanes here
      const int group0 = b * g.G + g.group_start[l] + (r - g.tile_start[l]) * (TW / VEC_CANON);  // first group of this tile
      if (lane < NOCT && ((oct_mask >> lane) & 1u)) {
        const int slot = base_oct + __popc(oct_mask & lt_mask);
        if (slot < f.tile_cap) f.tile_list[slot] = group0 / LPO_CANON + lane;
      }
      // the group's flag byte: bit i = anchor i of the group survived; RG lanes hold one group
      uint32_t gflags = flags << ((lane % RG) * VEC);
#pragma unroll
      for (int o = 1; o < RG; o <<= 1) gflags |= __shfl_xor_sync(0xffffffffu, gflags, o);
      // (a lane past the level's end would alias the next level's groups: not written)
      if (in_level && (lane % RG) == 0) f.tile_flags[group0 + lane / RG] = static_cast<uint8_t>(gflags);
      if (my_rows > 0) {
        uint64_t* keys = f.keys + static_cast<long long>(b) * f.rows_cap;
        int rpos = base_rows + inc - my_rows;
        if constexpr (MULTI) {
          int cur[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) { cur[i] = rpos; rpos += rows[i]; }
          for (int c = 0; c < nc; ++c) {  // second pass over the tile, still resident in the stage
            if (!class_ok(f.class_mask, c)) continue;
            const Pack<TI, VEC> p = row(c);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              const float sc = DV::rnd(sigmoid_f(DV::to_f(p.v[i])));
              if (rows[i] > 0 && sc > conf) {
                if (cur[i] < f.rows_cap) keys[cur[i]] = make_key(sc, (static_cast<uint32_t>(a_glob + i) << f.cls_bits) + c);
                ++cur[i];
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if (rows[i] == 0) continue;
            if (rpos < f.rows_cap) keys[rpos] = make_key(score[i], (static_cast<uint32_t>(a_glob + i) << f.cls_bits) + cls[i]);
            ++rpos;
          }
        }
      }
    }
    if constexpr (MULTI) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);  // the stage may be refilled (the key emission above re-read it)
    }
    s += s_step;
    fill += fill_step;
    if (s >= S) { s -= S; ++fill; }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

template <int DT_IN, int LB>
cudaError_t launch_tma(const TmaMaps& maps, const TmaScanGeom& g, const FilterArgs& f, size_t smem, int grid, cudaStream_t st) {
  cudaError_t e;
  const int threads = 32 * (g.ncw + 1);
  if (f.multi_label) {
    e = cudaFuncSetAttribute(scan_classes_tma_kernel<DT_IN, true, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    scan_classes_tma_kernel<DT_IN, true, LB><<<grid, threads, smem, st>>>(maps, g, f);
  } else {
    e = cudaFuncSetAttribute(scan_classes_tma_kernel<DT_IN, false, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    scan_classes_tma_kernel<DT_IN, false, LB><<<grid, threads, smem, st>>>(maps, g, f);
  }
  return cudaGetLastError();
}

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e && e[0] ? std::atoi(e) : dflt;
}

}  // namespace

// Returns cudaErrorNotSupported when the geometry does not fit the TMA form (the caller then launches the LDG kernel).
cudaError_t launch_scan_classes_tma(const HeadGeom& hg, int in_dtype, const FilterArgs& f, int vec, cudaStream_t st) {
  const int es = in_dtype == YPB_F32 ? 4 : 2;
  const int wide = 16 / es;
  EncodeTiledFn enc = encode_fn();
  if (!enc || vec != wide || hg.nc > 256 || hg.nc < 1 || hg.batch < 1) return cudaErrorNotSupported;
  // tuning knobs (diagnostic): bytes per lane and row, consumer warps
  static const int lb_env = env_int("YPB_TMA_LB", 0), ncw_env = env_int("YPB_TMA_NCW", 0);
  static const int multi_env = env_int("YPB_TMA_MULTI", 0);
  // multi-label (val mode) evaluates a sigmoid per ELEMENT and walks every tile twice: compute-bound in the consumer warps,
  // where the one-wave LDG kernel with its ~30 resident warps per SM is the faster form (measured, profiles/r02_scan_tune.jsonl)
  if (f.multi_label && !multi_env) return cudaErrorNotSupported;
  int lb = lb_env == 16 || lb_env == 8 || lb_env == 4 ? lb_env : 8;
  int ncw = ncw_env >= 1 && ncw_env <= TMA_MAX_NCW ? ncw_env : 12;
  const int row_bytes = 32 * lb;
  const size_t stage_bytes = static_cast<size_t>(hg.nc) * row_bytes;
  // ring size: all the shared memory of an SM by default; YPB_TMA_SMEM_KB caps it (leaves room for other streams' CTAs)
  static const int smem_kb_env = env_int("YPB_TMA_SMEM_KB", 0);
  const size_t ring_bytes = (smem_kb_env >= 16 && smem_kb_env <= 200 ? smem_kb_env : 200) * 1024u;
  int stages = static_cast<int>(ring_bytes / stage_bytes);
  if (stages > TMA_MAX_STAGES) stages = TMA_MAX_STAGES;
  if (stages < 2) return cudaErrorNotSupported;
  if (ncw > stages) ncw = stages;
  const int tw = row_bytes / es;
  TmaScanGeom g{};
  TmaMaps maps{};
  g.num_levels = hg.num_levels; g.batch = hg.batch; g.nc = hg.nc; g.stages = stages; g.ncw = ncw;
  int ts = 0;
  for (int l = 0; l < hg.num_levels; ++l) {
    const int hw = hg.h[l] * hg.w[l];
    g.tile_start[l] = ts;
    ts += (hw + tw - 1) / tw;
    g.level_anchors[l] = hw;
    g.anchor_start[l] = hg.anchor_start[l];
    g.group_start[l] = hg.group_start[l];
    if (hg.group_start[l] % (wide >= 8 ? 1 : 8 / wide) || hw % 8) return cudaErrorNotSupported;  // octets must not straddle levels
    // tensor map over (B, 64 + nc, H*W): innermost = anchors
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(hw), static_cast<cuuint64_t>(64 + hg.nc), static_cast<cuuint64_t>(hg.batch)};
    const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(hg.cstride[l]) * es, static_cast<cuuint64_t>(hg.bstride[l]) * es};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(hg.nc), 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (gstr[0] % 16 || (hg.batch > 1 && gstr[1] % 16) || gstr[0] >= (1ull << 40) || gstr[1] >= (1ull << 40)) return cudaErrorNotSupported;
    const CUtensorMapDataType dt = in_dtype == YPB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : in_dtype == YPB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUresult r = enc(&maps.m[l], dt, 3, const_cast<void*>(hg.ptr[l]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorNotSupported;
  }
  for (int l = hg.num_levels; l <= YPB_MAX_LEVELS; ++l) {
    g.tile_start[l] = ts;
    g.anchor_start[l] = hg.anchor_start[l];
    g.group_start[l] = hg.group_start[l];
  }
  g.tiles_per_image = ts;
  // the same per-image span of the octet index space as the LDG kernel (its grid covers ceil(groups / 128) * 128 lanes), so
  // decode_tiles_kernel is launched identically after either scan
  g.G = (hg.group_start[hg.num_levels] + 127) / 128 * 128;
  const long long total = static_cast<long long>(hg.batch) * ts;
  if (total >= (1ll << 31)) return cudaErrorNotSupported;
  int grid = sm_count();
  if (grid > total) grid = static_cast<int>(total);
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 2 * TMA_MAX_STAGES * sizeof(uint64_t);
#define YPB_TMA(DT)                                                                  \
  do {                                                                               \
    if (lb == 16) return launch_tma<DT, 16>(maps, g, f, smem, grid, st);             \
    if (lb == 8) return launch_tma<DT, 8>(maps, g, f, smem, grid, st);               \
    return launch_tma<DT, 4>(maps, g, f, smem, grid, st);                            \
  } while (0)
  switch (in_dtype) {
    case YPB_F32: YPB_TMA(YPB_F32);
    case YPB_F16: YPB_TMA(YPB_F16);
    case YPB_BF16: YPB_TMA(YPB_BF16);
  }
#undef YPB_TMA
  return cudaErrorInvalidValue;
}

}  // namespace ypb
