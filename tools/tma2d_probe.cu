// Diagnostic (not part of the library): persistent ring fed by ONE cp.async.bulk.tensor.3d (TMA, tensor map) per tile, in the
// access shape of the class scan: box = TW anchors x 80 class rows of one image.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tma2d tools/tma2d_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int TWB, int NR, int S, int NCW>   // TWB = bytes per tile row
__global__ void __launch_bounds__(32 * (NCW + 1)) ring_kernel(const __grid_constant__ CUtensorMap tm, int tw_elems, int cols, int B, int row0, int* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * NR * TWB);
  uint64_t* empty = full + S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = (long long)B * cols;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NCW) {
    if (lane == 0) {
      int it = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int s = it % S;
        if (it >= S) mbar_wait(&empty[s], ((it / S) - 1) & 1);
        const int b = (int)(t / cols), cb = (int)(t % cols);
        mbar_expect(&full[s], NR * TWB);
        tma_load_3d(smem + (size_t)s * NR * TWB, &tm, cb * tw_elems, row0, b, &full[s]);
      }
    }
  } else {
    int acc = 0, it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % S;
      if (s % NCW != warp) continue;
      mbar_wait(&full[s], (it / S) & 1);
      const char* tile = reinterpret_cast<const char*>(smem) + (size_t)s * NR * TWB;
      if (TWB >= 512) {
        for (int r = 0; r < NR; ++r)
          for (int c = lane * 16; c < TWB; c += 512) { int4 v = *reinterpret_cast<const int4*>(tile + r * TWB + c); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      } else if (TWB >= 256) {
        for (int r = 0; r < NR; ++r) { int2 v = *reinterpret_cast<const int2*>(tile + r * TWB + lane * 8); acc ^= v.x ^ v.y; }
      } else {
        for (int r = 0; r < NR; ++r) { int v = *reinterpret_cast<const int*>(tile + r * TWB + lane * 4); acc ^= v; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678) *sink = acc;
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int B = 64, C = 144, NC = 80, A = 8400;
  const int es = argc > 1 ? atoi(argv[1]) : 4;  // element size
  const size_t bytes = (size_t)B * C * A * es;
  const int NBUF = 3;
  char* buf[NBUF];
  for (int i = 0; i < NBUF; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], i + 1, bytes)); }
  int* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int IT = 30;
#define RUN(TW, S, NCW, GRID, L2P)                                                                                       \
  do {                                                                                                                   \
    constexpr int TWB_ = TW;                                                                                             \
    const int tw = TWB_ / es;                                                                                            \
    CUtensorMap tms[NBUF];                                                                                               \
    bool ok = true;                                                                                                      \
    for (int i = 0; i < NBUF; ++i) {                                                                                     \
      cuuint64_t gdim[3] = {(cuuint64_t)A, (cuuint64_t)C, (cuuint64_t)B};                                                \
      cuuint64_t gstr[2] = {(cuuint64_t)A * es, (cuuint64_t)C * A * es};                                                 \
      cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)NC, 1};                                                           \
      cuuint32_t estr[3] = {1, 1, 1};                                                                                    \
      CUresult r = encode(&tms[i], es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, buf[i], gdim, gstr, box, estr, \
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, L2P, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); \
      if (r != CUDA_SUCCESS) { printf("encode failed %d (TWB=%d)\n", (int)r, TWB_); ok = false; }                        \
    }                                                                                                                    \
    if (ok) {                                                                                                            \
      const int cols = (A + tw - 1) / tw;                                                                                \
      size_t sm = (size_t)S * NC * TWB_ + 2 * S * 8;                                                                     \
      CK(cudaFuncSetAttribute(ring_kernel<TWB_, NC, S, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
      for (int w = 0; w < 3; ++w) ring_kernel<TWB_, NC, S, NCW><<<GRID, 32 * (NCW + 1), sm>>>(tms[w % NBUF], tw, cols, B, 64, sink); \
      CK(cudaDeviceSynchronize());                                                                                       \
      CK(cudaEventRecord(e0));                                                                                           \
      for (int w = 0; w < IT; ++w) ring_kernel<TWB_, NC, S, NCW><<<GRID, 32 * (NCW + 1), sm>>>(tms[w % NBUF], tw, cols, B, 64, sink); \
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());                                     \
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                                                   \
      double nb = (double)B * NC * A * es;                                                                               \
      printf("tma3d rowbytes=%4d S=%d NCW=%d grid=%3d smem=%3zuKB l2p=%d  %7.2f us  %7.1f GB/s\n", TWB_, S, NCW, GRID, sm >> 10, (int)L2P, ms * 1e3 / IT, nb * IT / (ms * 1e-3) / 1e9); \
    }                                                                                                                    \
  } while (0)
  RUN(512, 4, 4, 148, CU_TENSOR_MAP_L2_PROMOTION_NONE);
  RUN(512, 4, 4, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(512, 4, 4, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  RUN(512, 5, 5, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(512, 4, 2, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(512, 2, 2, 296, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(256, 8, 4, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(256, 8, 8, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(256, 4, 4, 296, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(256, 3, 3, 444, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(128, 8, 8, 296, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(128, 16, 8, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  RUN(1024, 2, 2, 148, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  printf("done\n");
  return 0;
}
