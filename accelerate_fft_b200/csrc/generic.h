// Arbitrary-length axis transforms (non power-of-two): mixed-radix shared-memory kernel for
// smooth lengths and Bluestein (chirp-z on the power-of-two engine) for everything else.
// cuFFT accepts every length and the reference's own property suite draws n in [1,1024]
// (/root/reference/test/Test/Base.hs:44-58), so the drop-in has to as well.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <functional>

struct b200fft_plan_s;

namespace b200fft {

struct GenericPass {
  int bluestein = 0;
  int tiny = 0;                   // N <= 32, not a power of two: one thread per line, direct sum
  long long O = 1, N = 1, I = 1;  // array viewed as [O][N][I], transform along N
  long long lines = 0;            // O*I
  // Bluestein
  long long M = 0;                // power of two >= 2N-1
  long long chunk_lines = 0;
  void* chirp = nullptr;          // device, N entries: exp(-i pi n^2/N)
  void* filt = nullptr;           // device, M entries: FFT_M(wrapped conj chirp) / M
  b200fft_plan_s* sub = nullptr;  // rows of length M, batch chunk_lines
  const void* fused = nullptr;    // contiguous lines, M <= 4096: the whole chain in one launch (bluestein_kernel.cuh)
  void* fused_tw = nullptr;       // device: stage twiddles of the M-point engine
  int fused_tl = 0, fused_threads = 0;
  size_t fused_smem = 0;
  // mixed radix
  int nstages = 0;
  int radix[40] = {0};
  void* tw = nullptr;             // device, N entries: exp(-2 pi i m/N)
  int TL = 1, threads = 0, tpl_log2 = 0;
  size_t smem = 0;
  size_t workspace_bytes = 0;
  char desc[256] = {0};
};

using Uploader = std::function<void*(const void* host, size_t bytes)>;

int plan_generic_axis(int is_double, long long O, long long N, long long I, GenericPass* gp, const Uploader& up);
void destroy_generic(GenericPass* gp);
cudaError_t launch_generic(int is_double, const GenericPass& gp, const void* src, void* dst, void* workspace, int inverse,
                           double scale, cudaStream_t stream, long long* nlaunches);
cudaError_t launch_copy_scale(int is_double, const void* src, void* dst, long long count, double scale, cudaStream_t stream,
                              long long* nlaunches);
cudaError_t launch_slab_pack(int is_double, bool pack, const void* src, void* dst, long long dl, long long h, long long w, int P,
                             cudaStream_t stream);
// DFT/Centre.hs as stand-alone elementwise passes (rank <= 3 viewed as [d][h][w])
cudaError_t launch_shift(int is_double, const void* src, void* dst, long long d, long long h, long long w, long long sd, long long sh,
                         long long sw, cudaStream_t stream);
cudaError_t launch_centre(int is_double, const void* src, void* dst, long long d, long long h, long long w, cudaStream_t stream);
int generic_set_attrs();
// stream-ordered allocation from the library-owned scratch pool (plan.cu); release with cudaFreeAsync
cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t stream);

}  // namespace b200fft
