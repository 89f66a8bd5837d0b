// Diagnostic (not part of the library): throughput of a persistent producer/consumer ring fed by cp.async.bulk row copies issued
// by all 32 lanes of a producer warp, in the access shape of the class scan (NR rows of ROWB bytes per tile, rows `pitch` apart).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/bulk_ring tools/bulk_ring_probe.cu
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// tile t -> (image, column block): rows of `rowlen` bytes, ROWB bytes per tile column block
template <int ROWB, int NR, int S, int NCW>
__global__ void __launch_bounds__(32 * (NCW + 1)) ring_kernel(const char* __restrict__ base, long long img_stride, long long pitch, int rowlen, int B, int* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * NR * ROWB);
  uint64_t* empty = full + S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cols = rowlen / ROWB;
  const long long total = (long long)B * cols;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NCW) {  // producer
    int it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % S;
      if (it >= S) mbar_wait(&empty[s], ((it / S) - 1) & 1);
      const int b = (int)(t / cols), cb = (int)(t % cols);
      const char* src = base + (long long)b * img_stride + (long long)cb * ROWB;
      if (lane == 0) mbar_expect(&full[s], NR * ROWB);
      __syncwarp();
      for (int r = lane; r < NR; r += 32) bulk_g2s(smem + ((size_t)s * NR + r) * ROWB, src + (long long)r * pitch, ROWB, &full[s]);
    }
  } else {
    int acc = 0, it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % S;
      if (s % NCW != warp) continue;
      mbar_wait(&full[s], (it / S) & 1);
      const char* tile = reinterpret_cast<const char*>(smem) + (size_t)s * NR * ROWB;
      if (ROWB >= 512) {
        for (int r = 0; r < NR; ++r)
          for (int c = lane * 16; c < ROWB; c += 512) { int4 v = *reinterpret_cast<const int4*>(tile + r * ROWB + c); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      } else {
        for (int r = 0; r < NR; ++r) { int2 v = *reinterpret_cast<const int2*>(tile + r * ROWB + lane * 8); acc ^= v.x ^ v.y; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678) *sink = acc;
  }
}
int main(int argc, char** argv) {
  const int B = 64, C = 144, NC = 80;
  const int rowlen = argc > 1 ? atoi(argv[1]) : 33600;  // bytes per row: 8400 fp32 = 33600, bf16 = 16800
  const long long pitch = rowlen, img_stride = (long long)C * rowlen;
  const size_t bytes = (size_t)B * C * rowlen;
  const int NBUF = 3;
  char* buf[NBUF];
  for (int i = 0; i < NBUF; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], i + 1, bytes)); }
  int* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int IT = 30;
#define RUN(ROWB, S, NCW, GRID)                                                                                          \
  do {                                                                                                                   \
    if (rowlen % ROWB == 0) {                                                                                            \
      size_t sm = (size_t)S * NC * ROWB + 2 * S * 8;                                                                     \
      CK(cudaFuncSetAttribute(ring_kernel<ROWB, NC, S, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
      for (int w = 0; w < 3; ++w) ring_kernel<ROWB, NC, S, NCW><<<GRID, 32 * (NCW + 1), sm>>>(buf[w % NBUF] + 64LL * rowlen, img_stride, pitch, rowlen, B, sink); \
      CK(cudaDeviceSynchronize());                                                                                       \
      CK(cudaEventRecord(e0));                                                                                           \
      for (int w = 0; w < IT; ++w) ring_kernel<ROWB, NC, S, NCW><<<GRID, 32 * (NCW + 1), sm>>>(buf[w % NBUF] + 64LL * rowlen, img_stride, pitch, rowlen, B, sink); \
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());                                     \
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                                                   \
      double nb = (double)B * NC * rowlen;                                                                               \
      printf("ring ROWB=%4d S=%d NCW=%d grid=%3d smem=%3zuKB  %7.2f us  %7.1f GB/s\n", ROWB, S, NCW, GRID, sm >> 10, ms * 1e3 / IT, nb * IT / (ms * 1e-3) / 1e9); \
    }                                                                                                                    \
  } while (0)
  RUN(256, 8, 4, 148); RUN(256, 8, 8, 148); RUN(256, 4, 4, 296); RUN(256, 4, 2, 296); RUN(256, 10, 5, 148);
  RUN(512, 4, 4, 148); RUN(512, 4, 2, 148); RUN(512, 5, 5, 148); RUN(512, 2, 2, 296); RUN(512, 2, 1, 296);
  RUN(1024, 2, 2, 148); RUN(1024, 2, 1, 148);
  RUN(128, 8, 8, 148); RUN(128, 16, 8, 148); RUN(128, 8, 4, 296);
  RUN(640, 4, 4, 148); RUN(320, 8, 4, 148); RUN(160, 16, 8, 148);
  RUN(336, 6, 6, 148); RUN(672, 4, 4, 148); RUN(336, 3, 3, 296); RUN(168, 12, 6, 148);
  RUN(192, 12, 6, 148); RUN(224, 10, 5, 148); RUN(448, 5, 5, 148); RUN(448, 4, 4, 148); RUN(480, 4, 4, 148); RUN(960, 2, 2, 148); RUN(1344, 2, 2, 148);
  RUN(672, 2, 2, 296); RUN(448, 2, 2, 296); RUN(336, 6, 3, 148); RUN(672, 4, 2, 148);
  printf("done\n");
  return 0;
}
