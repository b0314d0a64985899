// tcgen05.ld (TMEM -> registers) throughput per SM as a function of the number of reading warps: the attention softmax
// warps read every fp32 score (and dP) tile out of TMEM, 64 KB per 128 x 128 tile.
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I open_genie_b200/csrc -o /tmp/tmem_ld_rate scripts/microbench/tmem_ld_rate.cu && /tmp/tmem_ld_rate
#include <cstdio>
#include <cuda.h>
#include "og_ptx.cuh"
using namespace og;

template <int X16>
__global__ void k_ld(unsigned* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  unsigned acc = 0;
  const int colgroups = blockDim.x >> 7;            // warps per lane quarter
  const int c0 = (warp >> 2) * (512 / colgroups);   // each warp sweeps its own column range
  for (int i = 0; i < iters; ++i) {
#pragma unroll 1
    for (int c = 0; c < 512 / colgroups; c += (X16 ? 16 : 32)) {
      if (X16) {
        uint32_t v[16];
        tmem_ld_32x16(base + lane_addr + c0 + c, v);
        tmem_ld_wait();
        acc += v[0] ^ v[15];
      } else {
        uint32_t v[32];
        tmem_ld_32x32(base + lane_addr + c0 + c, v);
        tmem_ld_wait();
        acc += v[0] ^ v[31];
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(base, 512);
  }
}

// two loads in flight before the wait (what a pipelined epilogue does)
__global__ void k_ld2(unsigned* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  unsigned acc = 0;
  const int colgroups = blockDim.x >> 7;
  const int c0 = (warp >> 2) * (512 / colgroups);
  for (int i = 0; i < iters; ++i) {
#pragma unroll 1
    for (int c = 0; c < 512 / colgroups; c += 64) {
      uint32_t v[32], w[32];
      tmem_ld_32x32(base + lane_addr + c0 + c, v);
      tmem_ld_32x32(base + lane_addr + c0 + c + 32, w);
      tmem_ld_wait();
      acc += v[0] ^ w[31];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(base, 512);
  }
}

int main() {
  unsigned* o;
  cudaMalloc(&o, 1 << 24);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;
  for (int variant = 0; variant < 3; ++variant)
    for (int warps : {4, 8, 16}) {
      auto launch = [&](int n) {
        if (variant == 0) k_ld<0><<<148, warps * 32>>>(o, n);
        else if (variant == 1) k_ld<1><<<148, warps * 32>>>(o, n);
        else k_ld2<<<148, warps * 32>>>(o, n);
      };
      launch(10);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      launch(iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)iters * 128 * 512 * 4;  // whole TMEM (256 KB) per iteration per SM
      printf("%s, %2d warps/SM: %.3f ms  %.1f B/clk/SM at %.2f GHz nominal (%s)\n",
             variant == 0 ? "32x32b.x32 + wait" : variant == 1 ? "32x32b.x16 + wait" : "2 x (32x32b.x32) + wait", warps, ms,
             bytes / (ms * 1e-3 * clk * 1e3), clk * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
