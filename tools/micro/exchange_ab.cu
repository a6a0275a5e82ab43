// A/B of the exchange between two register stages of a 1024-point c64 line held by ONE warp (32 threads x 32 points, cfg1's
// kernel): after the first radix-32 stage thread t holds row t of a 32 x 32 matrix of complex values and needs column t.
//   A: padded shared memory -- 32 STS.64 + 32 LDS.64 per thread (what fft_kernel.cuh does; pitch 33 elements, conflict free)
//   B: warp shuffles        -- the log2(32) = 5-level butterfly transpose, 16 exchanges of one complex value per level
// Timed in isolation (no butterflies), ITER exchanges back to back, 4 warps per CTA, 16 CTAs per SM.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/exchange_ab tools/micro/exchange_ab.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2000;

__global__ void __launch_bounds__(128) ex_smem(float2* out, int iters) {
  __shared__ float2 sm[4][32 * 33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float2 v[32];
#pragma unroll
  for (int e = 0; e < 32; e++) v[e] = make_float2((float)(lane * 32 + e), (float)blockIdx.x);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int e = 0; e < 32; e++) sm[w][lane * 33 + e] = v[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = sm[w][e * 33 + lane];
    __syncwarp();
  }
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 32; e++) { acc.x += v[e].x; acc.y += v[e].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(128) ex_shfl(float2* out, int iters) {
  const int lane = threadIdx.x & 31;
  float2 v[32];
#pragma unroll
  for (int e = 0; e < 32; e++) v[e] = make_float2((float)(lane * 32 + e), (float)blockIdx.x);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
      const bool up = (lane & k) != 0;
#pragma unroll
      for (int e = 0; e < 32; e++) {
        if ((e & k) == 0) {
          // lanes with bit k clear keep v[e] and send v[e|k]; lanes with bit k set keep v[e|k] and send v[e]
          float2 send = up ? v[e] : v[e | k];
          send.x = __shfl_xor_sync(0xffffffffu, send.x, k);
          send.y = __shfl_xor_sync(0xffffffffu, send.y, k);
          if (up) v[e] = send; else v[e | k] = send;
        }
      }
    }
  }
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 32; e++) { acc.x += v[e].x * (e + 1); acc.y += v[e].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int grid = nsm * 16, threads = 128;
  float2* out; cudaMalloc(&out, (size_t)grid * threads * sizeof(float2));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int rep = 0; rep < 2; rep++)
    for (int which = 0; which < 2; which++) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      if (which == 0) ex_smem<<<grid, threads>>>(out, 10); else ex_shfl<<<grid, threads>>>(out, 10);
      cudaDeviceSynchronize();
      cudaEventRecord(a);
      if (which == 0) ex_smem<<<grid, threads>>>(out, ITER); else ex_shfl<<<grid, threads>>>(out, ITER);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      // per SM: 16 CTAs x 4 warps exchange ITER times 32 x 32 complex values (8 KB per warp-exchange)
      const double exch_per_sm = 64.0 * ITER, bytes = exch_per_sm * 8192.0;
      printf("%-6s %8.3f ms   %.1f ns per warp-exchange (64 warps/SM in flight)   %.1f GB/s per SM = %.1f B/clk at %d MHz max\n",
             which == 0 ? "smem" : "shfl", ms, ms * 1e6 / exch_per_sm * 64.0, bytes / ms / 1e6, bytes / (ms * 1e-3) / (clk * 1e3), clk / 1000);
    }
  return 0;
}
