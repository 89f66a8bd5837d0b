// Diagnostic (not part of the library): what a read-only stream of the class rows can reach on this GPU, by technique.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/read_bw tools/read_bw_bench.cu && gpurun_out/read_bw
// Pattern "rows": B images x 80 class rows x A anchors (fp32) inside a (B,144,A) level block, i.e. exactly what
// scan_classes_kernel reads (80 rows of 4*A bytes, stride 144 rows between images).  "flat": the same bytes contiguous.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ int4 ldnc(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// flat grid-stride read, U loads in flight per thread
template <int U>
__global__ void flat_kernel(const int4* __restrict__ p, long long n, int* sink) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  int acc = 0;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldnc(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n; i += stride) { int4 v = ldnc(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678) *sink = acc;
}

// rows pattern, one thread = 4 anchors, walks the 80 rows with U loads in flight (the scan kernel's shape)
template <int U>
__global__ void rows_kernel(const float* __restrict__ base, int A, int rows, long long img_stride, int* sink) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 4 >= A) return;
  const float* p = base + (long long)blockIdx.y * img_stride + g * 4;
  int acc = 0;
  for (int c = 0; c < rows; c += U) {
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldnc(reinterpret_cast<const int4*>(p + (long long)(c + u) * A));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678) *sink = acc;
}

// rows pattern, software-pipelined: the next U loads are issued before the current U are consumed
template <int U>
__global__ void rows_pipe_kernel(const float* __restrict__ base, int A, int rows, long long img_stride, int* sink) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 4 >= A) return;
  const float* p = base + (long long)blockIdx.y * img_stride + g * 4;
  int acc = 0;
  int4 cur[U], nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) cur[u] = ldnc(reinterpret_cast<const int4*>(p + (long long)u * A));
  for (int c = U; c < rows; c += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) nxt[u] = ldnc(reinterpret_cast<const int4*>(p + (long long)(c + u) * A));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= cur[u].x ^ cur[u].y ^ cur[u].z ^ cur[u].w;
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
  }
#pragma unroll
  for (int u = 0; u < U; ++u) acc ^= cur[u].x ^ cur[u].y ^ cur[u].z ^ cur[u].w;
  if (acc == 0x12345678) *sink = acc;
}

// rows pattern through the bulk-copy engine: persistent CTAs, tile = ROWS_T rows x TW anchors, STAGES smem stages filled by
// cp.async.bulk (one 1-D copy per row) and consumed from shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TW, int ROWS_T, int STAGES>
__global__ void __launch_bounds__(256) rows_bulk_kernel(const float* __restrict__ base, int A, int rows, long long img_stride, int B,
                                                         int* sink) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* buf = reinterpret_cast<float*>(smem_raw);                                   // STAGES x ROWS_T x TW
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * ROWS_T * TW * 4);  // STAGES
  const int tid = threadIdx.x;
  const int tiles_a = A / TW, row_tiles = rows / ROWS_T;
  const long long total = (long long)B * tiles_a * row_tiles;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long long t, int s) {
    const int rt = (int)(t % row_tiles);
    const long long t2 = t / row_tiles;
    const int ta = (int)(t2 % tiles_a), b = (int)(t2 / tiles_a);
    const float* src = base + (long long)b * img_stride + (long long)rt * ROWS_T * A + ta * TW;
    mbar_expect(&full[s], ROWS_T * TW * 4);
    for (int r = 0; r < ROWS_T; ++r) bulk_g2s(buf + ((size_t)s * ROWS_T + r) * TW, src + (long long)r * A, TW * 4, &full[s]);
  };
  long long t = blockIdx.x;
  if (tid == 0) {
    long long tt = t;
    for (int s = 0; s < STAGES && tt < total; ++s, tt += gridDim.x) issue(tt, s);
  }
  int acc = 0, it = 0;
  for (; t < total; t += gridDim.x, ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const int4* tile = reinterpret_cast<const int4*>(buf + (size_t)s * ROWS_T * TW);
    for (int i = tid; i < ROWS_T * TW / 4; i += blockDim.x) { int4 v = tile[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    __syncthreads();  // everyone done with stage s
    const long long nt = t + (long long)STAGES * gridDim.x;
    if (tid == 0 && nt < total) issue(nt, s);
  }
  if (acc == 0x12345678) *sink = acc;
}

int main(int argc, char** argv) {
  const int B = 64, C = 144, NC = 80;
  const int A = argc > 1 ? atoi(argv[1]) : 6400 + 1600 + 400;  // 4200 emulates the bf16 rows (8400 x 2 B)  // level tensors concatenated per image for simplicity: rows of A
  const long long img_stride = (long long)C * A;
  const size_t bytes = (size_t)B * C * A * 4;
  const int NBUF = 3;
  float* buf[NBUF];
  for (int i = 0; i < NBUF; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], i + 1, bytes)); }
  int* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double rows_bytes = (double)B * NC * A * 4;
  const int IT = 30;
  auto report = [&](const char* name, double nbytes, float ms) { printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms * 1e3 / IT, nbytes * IT / (ms * 1e-3) / 1e9); };
#define TIME(name, nbytes, launch)                                             \
  do {                                                                         \
    for (int w = 0; w < 3; ++w) { const float* P = buf[w % NBUF]; (void)P; launch; } \
    CK(cudaDeviceSynchronize());                                               \
    CK(cudaEventRecord(e0));                                                   \
    for (int w = 0; w < IT; ++w) { const float* P = buf[w % NBUF]; (void)P; launch; } \
    CK(cudaEventRecord(e1));                                                   \
    CK(cudaEventSynchronize(e1));                                              \
    CK(cudaGetLastError());                                                    \
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                           \
    report(name, nbytes, ms);                                                  \
  } while (0)

  const long long n16 = (long long)(rows_bytes / 16);
  TIME("flat 172MB U=8  grid 148*8 x256", rows_bytes, (flat_kernel<8><<<148 * 8, 256>>>(reinterpret_cast<const int4*>(P), n16, sink)));
  TIME("flat 172MB U=16 grid 148*8 x256", rows_bytes, (flat_kernel<16><<<148 * 8, 256>>>(reinterpret_cast<const int4*>(P), n16, sink)));
  TIME("flat 172MB U=8  grid 148*16 x128", rows_bytes, (flat_kernel<8><<<148 * 16, 128>>>(reinterpret_cast<const int4*>(P), n16, sink)));
  TIME("flat 310MB U=8  grid 148*8 x256", (double)bytes, (flat_kernel<8><<<148 * 8, 256>>>(reinterpret_cast<const int4*>(P), (long long)(bytes / 16), sink)));
  {
    dim3 g((A / 4 + 127) / 128, B);
    TIME("rows U=8  128thr (scan shape)", rows_bytes, (rows_kernel<8><<<g, 128>>>(P + 64LL * A, A, NC, img_stride, sink)));
    TIME("rows U=16 128thr", rows_bytes, (rows_kernel<16><<<g, 128>>>(P + 64LL * A, A, NC, img_stride, sink)));
    TIME("rows U=4  128thr", rows_bytes, (rows_kernel<4><<<g, 128>>>(P + 64LL * A, A, NC, img_stride, sink)));
    TIME("rows pipelined U=4 128thr", rows_bytes, (rows_pipe_kernel<4><<<g, 128>>>(P + 64LL * A, A, NC, img_stride, sink)));
    TIME("rows pipelined U=8 128thr", rows_bytes, (rows_pipe_kernel<8><<<g, 128>>>(P + 64LL * A, A, NC, img_stride, sink)));
    dim3 g2((A / 4 + 63) / 64, B);
    TIME("rows U=8  64thr", rows_bytes, (rows_kernel<8><<<g2, 64>>>(P + 64LL * A, A, NC, img_stride, sink)));
    TIME("rows pipelined U=8 64thr", rows_bytes, (rows_pipe_kernel<8><<<g2, 64>>>(P + 64LL * A, A, NC, img_stride, sink)));
    dim3 g3((A / 4 + 255) / 256, B);
    TIME("rows U=8  256thr", rows_bytes, (rows_kernel<8><<<g3, 256>>>(P + 64LL * A, A, NC, img_stride, sink)));
  }
  {
    // bulk-copy tiles: A = 8400 = 2^4 * 525 -> TW must divide 8400: 240 (35 tiles), 400 (21), 560 (15), 1200 (7)
#define BULK(TW, RT, ST, GRID)                                                                                         \
    do {                                                                                                               \
      size_t sm = (size_t)ST * RT * TW * 4 + ST * 8;                                                                   \
      CK(cudaFuncSetAttribute(rows_bulk_kernel<TW, RT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));    \
      char nm[96]; snprintf(nm, sizeof nm, "rows bulk TW=%d RT=%d ST=%d grid=%d", TW, RT, ST, GRID);                   \
      TIME(nm, rows_bytes, (rows_bulk_kernel<TW, RT, ST><<<GRID, 256, sm>>>(P + 64LL * A, A, NC, img_stride, B, sink))); \
    } while (0)
    if (A == 8400) {
    BULK(240, 16, 4, 148);
    BULK(240, 16, 8, 148);
    BULK(240, 16, 4, 296);
    BULK(400, 16, 4, 148);
    BULK(400, 16, 6, 148);
    BULK(400, 16, 4, 296);
    BULK(560, 16, 4, 148);
    BULK(1200, 8, 4, 148);
    BULK(240, 80, 2, 148);
    BULK(400, 40, 3, 148);
    BULK(80, 80, 4, 148);
    BULK(80, 80, 6, 148);
    BULK(80, 80, 4, 296);
    BULK(80, 80, 3, 444);
    BULK(160, 80, 3, 148);
    BULK(160, 80, 2, 296);
    } else if (A == 4200) {
    BULK(168, 16, 4, 148);
    BULK(280, 16, 4, 148);
    BULK(280, 16, 6, 148);
    BULK(280, 16, 4, 296);
    BULK(600, 8, 4, 148);
    BULK(168, 80, 2, 148);
    BULK(168, 80, 3, 148);
    BULK(280, 40, 3, 148);
    BULK(40, 80, 6, 148);
    BULK(40, 80, 8, 296);
    BULK(40, 80, 4, 444);
    BULK(120, 80, 4, 148);
    BULK(120, 80, 3, 296);
    }
  }
  printf("done\n");
  return 0;
}
