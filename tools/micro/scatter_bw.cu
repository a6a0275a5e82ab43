// Micro-benchmark: bandwidth of page-scattered row pieces, the access pattern of a strided (column) FFT pass.
// A "tile" = R rows x RUN bytes; its rows are H/R rows apart (so every tile spans the whole H x PITCH array); tiles advance
// along the row index first (k1 fastest), then along the columns -- the order of the band kernel's phase-B stores.
//   mode 0: write only   mode 1: read only   mode 2: read + write in place
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/scatter_bw tools/micro/scatter_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC>   // bytes per lane: 8 or 16
__global__ void k(char* base, long long H, long long pitch, int R, int run, int mode, long long ntiles, float* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int lanes_per_row = run / VEC;            // lanes covering one row piece
  const int rows_per_instr = 32 / lanes_per_row;  // rows one warp instruction covers
  const long long stride_rows = H / R;
  const long long tiles_per_col = stride_rows;    // k1 = 0 .. H/R-1
  float acc = 0.f;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long k1 = t % tiles_per_col, cb = t / tiles_per_col;
    char* tb = base + k1 * pitch + cb * run;
    for (int r0 = warp * rows_per_instr; r0 < R; r0 += nw * rows_per_instr) {
      const int r = r0 + lane / lanes_per_row;
      char* p = tb + (long long)r * stride_rows * pitch + (lane % lanes_per_row) * VEC;
      if (VEC == 8) {
        float2 v = make_float2((float)t, (float)r);
        if (mode >= 1) { v = __ldcs(reinterpret_cast<const float2*>(p)); acc += v.x; }
        if (mode != 1) { v.x += 1.f; __stcs(reinterpret_cast<float2*>(p), v); }
      } else {
        float4 v = make_float4((float)t, (float)r, 0.f, 0.f);
        if (mode >= 1) { v = __ldcs(reinterpret_cast<const float4*>(p)); acc += v.x; }
        if (mode != 1) { v.x += 1.f; __stcs(reinterpret_cast<float4*>(p), v); }
      }
    }
  }
  if (acc == 12345.678f) *sink = acc;
}

int main(int argc, char** argv) {
  float* sink; cudaMalloc(&sink, 4);
  const long long pitch = 65536;
  printf("%-8s %-6s %-6s %-5s %-10s %s\n", "spanMB", "R", "run", "mode", "us", "GB/s (bytes touched / time)");
  for (long long H : {2048LL, 4096LL, 8192LL, 32768LL}) {
    char* buf; if (cudaMalloc(&buf, H * pitch) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, H * pitch);
    for (int R : {64, 128}) for (int run : {128, 256, 512}) for (int mode : {0, 1, 2}) {
      const long long ntiles = (H / R) * (pitch / run);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      const int grid = 148 * 4, threads = 256;
      auto launch = [&] {
        if (run == 512) k<16><<<grid, threads>>>(buf, H, pitch, R, run, mode, ntiles, sink);
        else k<8><<<grid, threads>>>(buf, H, pitch, R, run, mode, ntiles, sink);
      };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(a); for (int i = 0; i < 3; i++) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
      const double bytes = (double)H * pitch * (mode == 2 ? 2 : 1);
      printf("%-8lld %-6d %-6d %-5d %-10.1f %.0f\n", H * pitch >> 20, R, run, mode, ms * 1e3, bytes / ms / 1e6);
    }
    cudaFree(buf);
  }
  return 0;
}
