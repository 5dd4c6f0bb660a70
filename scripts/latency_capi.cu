// Wall-clock latency of the C ABI itself (what a Julia `ccall` pays), without the Python wrapper:
// single-pair dcol_proximity / dcol_proximity_jacobian with host and device pointers, and small cube-cube batches (BASELINE config 1).
// build: nvcc -O2 -o build_variants/latency_capi scripts/latency_capi.cu -Iinclude -Ldifferentiablecollisions.jl_b200 -ldcol_b200 -Xlinker -rpath,'$ORIGIN/../differentiablecollisions.jl_b200'
#include "dcol_b200.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  dcol_ctx* ctx = nullptr;
  if (dcol_create(&ctx, 0)) { printf("dcol_create failed: %s\n", dcol_last_error(nullptr)); return 1; }
  // shape table: 0 cone(2.0, 22 deg), 1 capsule(0.3, 1.2), 2 unit cube
  std::vector<double> P; std::vector<int64_t> off; std::vector<int32_t> type;
  auto tail = [&]() { for (int k = 0; k < 3; k++) P.push_back(0.0); for (int k = 0; k < 9; k++) P.push_back(k % 4 == 0 ? 1.0 : 0.0); };
  type.push_back(3); off.push_back((int64_t)P.size()); P.push_back(2.0); P.push_back(22.0 * M_PI / 180.0); tail();
  type.push_back(1); off.push_back((int64_t)P.size()); P.push_back(0.3); P.push_back(1.2); tail();
  type.push_back(0); off.push_back((int64_t)P.size()); P.push_back(6.0);
  const double A[18] = {1, 0, 0, 0, 1, 0, 0, 0, 1, -1, 0, 0, 0, -1, 0, 0, 0, -1};
  for (double v : A) P.push_back(v);
  for (int k = 0; k < 6; k++) P.push_back(0.5);
  tail();
  if (dcol_set_shapes(ctx, 3, type.data(), off.data(), P.data(), (int64_t)P.size())) { printf("set_shapes: %s\n", dcol_last_error(ctx)); return 1; }

  const int NMAX = 100000;
  std::mt19937_64 g(1234); std::normal_distribution<double> N01(0.0, 1.0);
  std::vector<int32_t> s1(NMAX), s2(NMAX), st(NMAX), it(NMAX);
  std::vector<double> r1(3 * NMAX), r2(3 * NMAX), q1(4 * NMAX), q2(4 * NMAX), alpha(NMAX), x(3 * NMAX), J(56 * (size_t)NMAX);
  for (int i = 0; i < NMAX; i++) {
    for (int k = 0; k < 3; k++) { r1[3 * i + k] = 2 * N01(g); r2[3 * i + k] = N01(g); }
    double n1 = 0, n2 = 0;
    for (int k = 0; k < 4; k++) { q1[4 * i + k] = N01(g); q2[4 * i + k] = N01(g); n1 += q1[4 * i + k] * q1[4 * i + k]; n2 += q2[4 * i + k] * q2[4 * i + k]; }
    for (int k = 0; k < 4; k++) { q1[4 * i + k] /= std::sqrt(n1); q2[4 * i + k] /= std::sqrt(n2); }
  }
  auto bench = [&](const char* what, int n, int sa, int sb, int kind, int mem, int reps) {
    for (int i = 0; i < n; i++) { s1[i] = sa; s2[i] = sb; }
    int32_t *ps1 = s1.data(), *ps2 = s2.data(), *pst = st.data(), *pit = it.data();
    double *pr1 = r1.data(), *pq1 = q1.data(), *pr2 = r2.data(), *pq2 = q2.data(), *pa = alpha.data(), *px = x.data(), *pJ = J.data();
    std::vector<void*> dev;
    if (mem == DCOL_MEM_DEVICE) {
      auto up = [&](const void* h, size_t bytes) { void* d; cudaMalloc(&d, bytes); if (h) cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); dev.push_back(d); return d; };
      ps1 = (int32_t*)up(ps1, 4 * n); ps2 = (int32_t*)up(ps2, 4 * n); pr1 = (double*)up(pr1, 24 * n); pq1 = (double*)up(pq1, 32 * n);
      pr2 = (double*)up(pr2, 24 * n); pq2 = (double*)up(pq2, 32 * n); pa = (double*)up(nullptr, 8 * n); px = (double*)up(nullptr, 24 * n);
      pJ = (double*)up(nullptr, 448 * (size_t)n); pst = (int32_t*)up(nullptr, 4 * n); pit = (int32_t*)up(nullptr, 4 * n);
    }
    auto call = [&]() {
      return kind == 0 ? dcol_proximity(ctx, n, ps1, ps2, pr1, pq1, pr2, pq2, DCOL_ATT_QUAT, 1e-6, mem, pa, px, pst, pit)
                       : dcol_proximity_jacobian(ctx, n, ps1, ps2, pr1, pq1, pr2, pq2, DCOL_ATT_QUAT, 1e-6, mem, pa, px, pJ, pst, pit);
    };
    for (int k = 0; k < 10; k++) if (call()) { printf("%s failed: %s\n", what, dcol_last_error(ctx)); return; }
    double t0 = now_us();
    for (int k = 0; k < reps; k++) call();
    double dt = (now_us() - t0) / reps;
    printf("%-34s n=%6d %-9s %-6s : %9.1f us per call   kernels %7.1f us   %8.3f Mpairs/s\n", what, n, kind == 0 ? "proximity" : "jacobian",
           mem == DCOL_MEM_HOST ? "host" : "device", dt, dcol_last_kernel_ms(ctx) * 1e3, n / dt);
    for (void* d : dev) cudaFree(d);
  };
  bench("single pair cone-capsule", 1, 0, 1, 0, DCOL_MEM_HOST, 500);
  bench("single pair cone-capsule", 1, 0, 1, 0, DCOL_MEM_DEVICE, 500);
  bench("single pair cone-capsule", 1, 0, 1, 2, DCOL_MEM_HOST, 500);
  bench("single pair cone-capsule", 1, 0, 1, 2, DCOL_MEM_DEVICE, 500);
  for (int n : {100, 1000, 10000, 100000}) {
    bench("cube-cube (config 1 at n=10000)", n, 2, 2, 0, DCOL_MEM_HOST, n >= 100000 ? 50 : 300);
    bench("cube-cube (config 1 at n=10000)", n, 2, 2, 0, DCOL_MEM_DEVICE, n >= 100000 ? 50 : 300);
  }
  bench("cube-cube", 10000, 2, 2, 2, DCOL_MEM_HOST, 100);
  dcol_destroy(ctx);
  return 0;
}
