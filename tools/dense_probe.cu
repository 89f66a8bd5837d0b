// Diagnostic (not part of the library): the dense decode kernels of ypb_decode.cu timed alone on C2 shapes, by variant.
// The translation unit is included as is, so every -D switch of the kernels (YPB_DD16_OLD, YPB_DFL_FORM, YPB_DD16_MIN_BLOCKS,
// YPB_DEC_THREADS ...) can be compared from binaries built side by side:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -prec-div=true -prec-sqrt=true [-D...] \
//        -o tools/_bin/dense_probe tools/dense_probe.cu && tools/_bin/dense_probe [batch] [reps]
// Values are checked against a double-precision host evaluation of head.py:151-169 on a sample of anchors.
#include "../ultralytics_pro_b200/csrc/ypb_decode.cu"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

template <typename T> static T host_from_float(float f);
template <> float host_from_float<float>(float f) { return f; }
template <> __half host_from_float<__half>(float f) { return __float2half_rn(f); }
template <> __nv_bfloat16 host_from_float<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }
static float host_to_float(float v) { return v; }
static float host_to_float(__half v) { return __half2float(v); }
static float host_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, int DT>
static int run(const char* name, int B, int reps) {
  const int nc = 80, C = 64 + nc, NL = 3, NSETS = 3;
  const int hw[NL] = {80, 40, 20};
  const float strides[NL] = {8.f, 16.f, 32.f};
  int A = 0;
  for (int l = 0; l < NL; ++l) A += hw[l] * hw[l];
  std::vector<T*> lv(NSETS * NL);
  std::vector<std::vector<T>> host(NL);
  uint32_t seed = 1234567u;
  for (int s = 0; s < NSETS; ++s)
    for (int l = 0; l < NL; ++l) {
      const size_t n = static_cast<size_t>(B) * C * hw[l] * hw[l];
      CK(cudaMalloc(&lv[s * NL + l], n * sizeof(T)));
      std::vector<T> h(n);
      for (size_t i = 0; i < n; ++i) h[i] = host_from_float<T>((static_cast<float>(lcg(seed) >> 8) / 16777216.f - 0.5f) * 14.f - 1.f);
      CK(cudaMemcpy(lv[s * NL + l], h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
      if (s == 0) host[l] = std::move(h);
    }
  const size_t out_n = static_cast<size_t>(B) * (4 + nc) * A;
  T* out[2];
  for (int i = 0; i < 2; ++i) CK(cudaMalloc(&out[i], out_n * sizeof(T)));
  const int vec = sizeof(T) == 2 ? 4 : 4;
  auto geom = [&](int s) {
    ypb::HeadGeom g{};
    g.num_levels = NL; g.batch = B; g.nc = nc; g.reg_max = 16;
    int as = 0, gs = 0;
    for (int l = 0; l < NL; ++l) {
      g.ptr[l] = lv[s * NL + l]; g.h[l] = hw[l]; g.w[l] = hw[l];
      g.cstride[l] = hw[l] * hw[l]; g.bstride[l] = static_cast<long long>(C) * hw[l] * hw[l];
      g.stride[l] = strides[l];
      g.anchor_start[l] = as; g.group_start[l] = gs;
      as += hw[l] * hw[l]; gs += hw[l] * hw[l] / vec;
    }
    for (int l = NL; l <= YPB_MAX_LEVELS; ++l) { g.anchor_start[l] = as; g.group_start[l] = gs; }
    g.anchors = A;
    return g;
  };
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int i = 0; i < 6; ++i)
    CK(ypb::launch_decode_dense(geom(i % NSETS), DT, nullptr, 0, 0, 0, out[i & 1], DT, static_cast<long long>(4 + nc) * A, A, vec, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; ++i)
    CK(ypb::launch_decode_dense(geom(i % NSETS), DT, nullptr, 0, 0, 0, out[i & 1], DT, static_cast<long long>(4 + nc) * A, A, vec, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1000.0 / reps;
  const double bytes = static_cast<double>(B) * (C + 4 + nc) * A * sizeof(T);
  // correctness on set 0 (run once more into out[0])
  CK(ypb::launch_decode_dense(geom(0), DT, nullptr, 0, 0, 0, out[0], DT, static_cast<long long>(4 + nc) * A, A, vec, st));
  CK(cudaStreamSynchronize(st));
  std::vector<T> ho(out_n);
  CK(cudaMemcpy(ho.data(), out[0], out_n * sizeof(T), cudaMemcpyDeviceToHost));
  double worst_box = 0, worst_cls = 0;
  int as = 0;
  for (int l = 0; l < NL; ++l) {
    const int n = hw[l] * hw[l];
    for (int b = 0; b < B; b += (B > 3 ? B / 3 : 1))
      for (int a = 0; a < n; a += 37) {
        const T* x = host[l].data() + static_cast<size_t>(b) * C * n + a;
        double d[4];
        for (int side = 0; side < 4; ++side) {
          double m = -1e30, se = 0, sp = 0;
          for (int k = 0; k < 16; ++k) m = std::fmax(m, static_cast<double>(host_to_float(x[static_cast<size_t>(side * 16 + k) * n])));
          for (int k = 0; k < 16; ++k) {
            const double e = std::exp(static_cast<double>(host_to_float(x[static_cast<size_t>(side * 16 + k) * n])) - m);
            se += e; sp += k * e;
          }
          d[side] = sp / se;
        }
        const double ax = a % hw[l] + 0.5, ay = a / hw[l] + 0.5;
        const double ref[4] = {(ax - d[0] + ax + d[2]) * 0.5 * strides[l], (ay - d[1] + ay + d[3]) * 0.5 * strides[l],
                               (d[0] + d[2]) * strides[l], (d[1] + d[3]) * strides[l]};
        const T* o = ho.data() + static_cast<size_t>(b) * (4 + nc) * A + as + a;
        for (int c = 0; c < 4; ++c) {
          const double got = host_to_float(o[static_cast<size_t>(c) * A]);
          worst_box = std::fmax(worst_box, std::fabs(got - ref[c]) / (std::fabs(ref[c]) + 1e-5 * 640));
        }
        for (int c = 0; c < nc; ++c) {
          const double r = 1.0 / (1.0 + std::exp(-static_cast<double>(host_to_float(x[static_cast<size_t>(64 + c) * n]))));
          const double got = host_to_float(o[static_cast<size_t>(4 + c) * A]);
          worst_cls = std::fmax(worst_cls, std::fabs(got - r) / r);
        }
      }
    as += n;
  }
  printf("{\"probe\": \"dense_decode\", \"dtype\": \"%s\", \"batch\": %d, \"us\": %.2f, \"gbs\": %.0f, \"frac\": %.3f, \"max_rel_box\": %.3g, \"max_rel_cls\": %.3g}\n",
         name, B, us, bytes / us / 1e3, bytes / us / 1e3 / 6550.4, worst_box, worst_cls);
  for (auto p : lv) cudaFree(p);
  for (int i = 0; i < 2; ++i) cudaFree(out[i]);
  return 0;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64;
  const int reps = argc > 2 ? atoi(argv[2]) : 60;
  const char* which = argc > 3 ? argv[3] : "all";
  int rc = 0;
  if (which[0] == 'a' || which[0] == 'b') rc |= run<__nv_bfloat16, YPB_BF16>("bf16", B, reps);
  if (which[0] == 'a' || which[0] == 'h') rc |= run<__half, YPB_F16>("fp16", B, reps);
  if (which[0] == 'a' || which[0] == 'f') rc |= run<float, YPB_F32>("fp32", B, reps);
  return rc;
}
