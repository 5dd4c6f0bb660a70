// Variants of the FP64 peak probe (development aid): which DFMA-only loop saturates the FP64 pipe?  Run under ncu
// (sm__pipe_fp64_cycles_active) and plain (TFLOP/s from CUDA events).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, int ROUNDS, int MODE>
__global__ void probe(double* out, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
#pragma unroll 1
  for (int i = 0; i < iters; i += ROUNDS) {
#pragma unroll
    for (int u = 0; u < ROUNDS; u++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (MODE == 0) x[c] = fma(x[c], a, b);
        else if (MODE == 1) x[c] = fma(x[c], a, x[c]);
        else x[c] = fma(a, b, x[c]);
      }
    }
  }
  double s = 0; for (int c = 0; c < CH; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH, int ROUNDS, int MODE>
void run(const char* name, int blocks, int threads, double* d) {
  const int iters = 20480;
  probe<CH, ROUNDS, MODE><<<blocks, threads>>>(d, 2048, 1.0000001, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); probe<CH, ROUNDS, MODE><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  printf("%-34s grid %5d x %4d: %7.3f ms  %6.2f TFLOP/s\n", name, blocks, threads, best, 2.0 * CH * (double)iters * blocks * threads / (best * 1e-3) / 1e12);
}
int main() {
  double* d; cudaMalloc(&d, sizeof(double) * 148 * 16 * 1024);
  run<8, 16, 0>("ch8 r16 fma(x,a,b) [current]", 148 * 8, 256, d);
  run<8, 64, 0>("ch8 r64 fma(x,a,b)", 148 * 8, 256, d);
  run<8, 64, 1>("ch8 r64 fma(x,a,x)", 148 * 8, 256, d);
  run<8, 64, 2>("ch8 r64 fma(a,b,x)", 148 * 8, 256, d);
  run<16, 32, 0>("ch16 r32 fma(x,a,b)", 148 * 8, 256, d);
  run<4, 64, 0>("ch4 r64 fma(x,a,b)", 148 * 8, 256, d);
  run<8, 64, 0>("ch8 r64 fma(x,a,b) 4x512", 148 * 4, 512, d);
  run<8, 64, 0>("ch8 r64 fma(x,a,b) 2x1024", 148 * 2, 1024, d);
  run<8, 64, 0>("ch8 r64 fma(x,a,b) 4x256 (half occ)", 148 * 4, 256, d);
  run<8, 64, 0>("ch8 r64 fma(x,a,b) 4x128", 148 * 4, 128, d);
  return 0;
}
