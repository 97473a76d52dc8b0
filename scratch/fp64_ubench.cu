// Micro-benchmark: cycles per DFMA for one warp with ILP independent accumulator chains, and with several warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// LDS + DFMA mix like the apply loop: per "row" 5 rhs values (registers) x NC weights from shared memory
template <int NC>
__global__ void k2(double *out, long long *cyc, int iters) {
  __shared__ double w[32 * 32 * 4];
  for (int i = threadIdx.x; i < 32 * 32 * 4; i += blockDim.x) w[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double *ww = w + (warp & 3) * 32 * 32 + lane;
  double acc[NC][5];
  for (int c = 0; c < NC; ++c) for (int v = 0; v < 5; ++v) acc[c][v] = 0;
  double rhs[5] = {1.0 + lane, 2.0, 3.0, 4.0, 5.0};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double wv = ww[((r * NC + c) & 31) * 32];
#pragma unroll
        for (int v = 0; v < 5; ++v) acc[c][v] = fma(wv, rhs[v], acc[c][v]);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < NC; ++c) for (int v = 0; v < 5; ++v) s += acc[c][v];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
#define RUN(ILP, THREADS) { k<ILP><<<148, THREADS>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("ILP %2d warps/SM %2d: %.2f cycles per DFMA per warp, %.1f DFMA lanes/clk/SM\n", ILP, THREADS / 32, (double)h / (iters * ILP), 32.0 * (THREADS / 32) * iters * ILP / (double)h); }
  RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(16, 32) RUN(32, 32)
  RUN(8, 128) RUN(16, 128) RUN(32, 128) RUN(8, 256) RUN(8, 512) RUN(8, 1024) RUN(4, 1024)
#define RUN2(NC, THREADS) { k2<NC><<<148, THREADS>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("apply-like NC %d warps/SM %2d: %.2f cycles per DFMA per warp, %.1f DFMA lanes/clk/SM\n", NC, THREADS / 32, (double)h / (iters * 3 * NC * 5), 32.0 * (THREADS / 32) * iters * 15 * NC / (double)h); }
  RUN2(9, 32) RUN2(9, 128) RUN2(9, 256) RUN2(3, 128) RUN2(3, 512)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
