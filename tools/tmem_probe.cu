// TMEM -> register bandwidth probe: how many bytes per clock one SM's warps can pull through tcgen05.ld (32x32b shape).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I effocr_b200/csrc -o tools/_bin/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace effocr;

template <int X>
__global__ void probe(long long* out, int iters, int wait_every) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t accum = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (X == 32) { uint32_t v[32]; tmem_ld_32x32b_x32(base + ((i * 32) & 511), v); accum ^= v[0] ^ v[31]; }
    else { uint32_t v[16]; tmem_ld_32x32b_x16(base + ((i * 16) & 511), v); accum ^= v[0] ^ v[15]; }
    if ((i + 1) % wait_every == 0) tmem_ld_wait();
  }
  tmem_ld_wait();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (accum == 0x12345678u) out[1] = accum;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int iters = 4096;
  for (int x : {16, 32})
    for (int warps : {1, 4, 8, 16})
      for (int we : {1, 4}) {
        if (x == 32) probe<32><<<1, warps * 32>>>(d, iters, we); else probe<16><<<1, warps * 32>>>(d, iters, we);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        const double bytes = double(iters) * warps * 32 * x * 4;
        printf("x%d warps %2d wait_every %d: %lld cycles, %.1f B/clk/SM, %.1f cycles per ld per warp  (%s)\n", x, warps, we, h[0],
               bytes / h[0], double(h[0]) / iters, cudaGetErrorString(e));
      }
  return 0;
}
