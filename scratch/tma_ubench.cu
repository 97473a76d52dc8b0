// Micro-benchmark: 1-D TMA bulk copies (cp.async.bulk global -> shared, UBLKCP) per SM: latency and throughput as a
// function of warps per CTA, ring depth and copy size.  Every warp streams its own region with a private ring, like
// recon_tile_kernel does, but consumes nothing.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_ubench tma_ubench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, int parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"((uint32_t)parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void k(const char *src, size_t region, int slots, int bytes, int iters, int spin, unsigned long long *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char *my = smem + (size_t)warp * (128 + (size_t)slots * bytes);
  uint64_t *bar = reinterpret_cast<uint64_t *>(my);
  unsigned char *ring = my + 128;
  const char *p = src + ((size_t)(blockIdx.x * nw + warp) * (size_t)iters * bytes) % region;
  if (lane == 0) {
    for (int s = 0; s < slots; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0)
    for (int s = 0; s < slots; ++s) { mbar_expect(&bar[s], bytes); bulk(ring + (size_t)s * bytes, p + (size_t)s * bytes, bytes, &bar[s]); }
  long long t0 = clock64(), waited = 0;
  int slot = 0, phase = 0;
  double acc = 0.0;
  for (int i = 0; i < iters; ++i) {
    long long a = clock64();
    mbar_wait(&bar[slot], phase);
    waited += clock64() - a;
    acc += reinterpret_cast<double *>(ring + (size_t)slot * bytes)[lane];
    for (int j = 0; j < spin; ++j) acc = fma(acc, 1.0000001, 1e-9);  // "compute" between copies (dependent DFMAs, 8 cycles each)
    __syncwarp();
    if (lane == 0 && i + slots < iters) { mbar_expect(&bar[slot], bytes); bulk(ring + (size_t)slot * bytes, p + (size_t)(i + slots) * bytes, bytes, &bar[slot]); }
    if (++slot == slots) { slot = 0; phase ^= 1; }
  }
  long long t1 = clock64();
  if (lane == 0) {
    atomicAdd(out, (unsigned long long)(t1 - t0));
    atomicAdd(out + 1, (unsigned long long)waited);
    if (acc == 12345.678) out[2] = 1;
  }
}

int main() {
  const size_t region_small = 48ull << 20, region_big = 8ull << 30;
  char *src;
  cudaMalloc(&src, region_big);
  cudaMemset(src, 0, region_big);
  unsigned long long *out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("%-6s %5s %5s %6s %5s | %10s %10s %10s\n", "mem", "warps", "slots", "bytes", "spin", "cyc/copy", "wait/copy", "B/clk/SM");
  for (int big = 0; big < 2; ++big)
    for (int warps : {1, 2, 4, 8})
      for (int slots : {1, 2, 3, 6})
        for (int bytes : {2304, 4608})
          for (int spin : {0, 40}) {
            if ((size_t)warps * (128 + (size_t)slots * bytes) > 227 * 1024) continue;
            const int iters = 2000;
            const size_t smem = (size_t)warps * (128 + (size_t)slots * bytes);
            for (int rep = 0; rep < 2; ++rep) {
              out[0] = out[1] = 0;
              k<<<148, 32 * warps, smem>>>(src, big ? region_big : region_small, slots, bytes, iters, spin, out);
              cudaDeviceSynchronize();
            }
            const double per = (double)out[0] / (148.0 * warps) / iters;
            const double wait = (double)out[1] / (148.0 * warps) / iters;
            printf("%-6s %5d %5d %6d %5d | %10.1f %10.1f %10.2f\n", big ? "hbm" : "l2", warps, slots, bytes, spin, per, wait, warps * bytes / per);
          }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
