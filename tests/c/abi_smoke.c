/* The boundary is a C ABI: this file is plain C (compiled with gcc -std=c99 -pedantic, no CUDA headers), includes
 * include/b200fft.h as a C header and drives the library through the host-buffer entry point the way a non-C++, non-Python
 * host would (the reference's host is Haskell: foreign import ccall needs exactly this).
 *   exit 0: transform computed and correct;   exit 3: no device (B200FFT_NO_DEVICE) -- the expected outcome on a CPU-only box;
 *   anything else: failure. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200fft.h"

int main(void) {
  const int64_t n = 1000, batch = 3; /* 1000 = 2^3 5^3: the mixed-radix path; rank-2 `fft` over the innermost axis */
  const int64_t shape[2] = {3, 1000};
  double *in = (double *)calloc((size_t)(2 * n * batch), sizeof(double));
  double *out = (double *)calloc((size_t)(2 * n * batch), sizeof(double));
  int64_t b, k;
  int status;
  double worst = 0.0;
  if (!in || !out) return 2;
  for (b = 0; b < batch; b++) in[2 * (b * n + (b + 1))] = 1.0; /* row b: delta at m = b + 1  ->  exp(-2 pi i m k / n) */
  status = accfft_run_host(0 /* fft */, 0 /* Forward */, 2, shape, B200FFT_Z2Z, in, out);
  if (status == B200FFT_NO_DEVICE) {
    printf("no device: %s\n", b200fftErrorString(status));
    return 3;
  }
  if (status != B200FFT_SUCCESS) {
    printf("accfft_run_host failed: %s\n", b200fftErrorString(status));
    return 1;
  }
  for (b = 0; b < batch; b++)
    for (k = 0; k < n; k++) {
      const double a = -2.0 * 3.14159265358979323846 * (double)(((b + 1) * k) % n) / (double)n;
      const double er = out[2 * (b * n + k)] - cos(a), ei = out[2 * (b * n + k) + 1] - sin(a);
      const double e = sqrt(er * er + ei * ei);
      if (e > worst) worst = e;
    }
  printf("max abs error %.3e over %ld x %ld outputs, %ld kernel launches\n", worst, (long)batch, (long)n, (long)b200fftKernelLaunches());
  return worst < 1e-12 ? 0 : 1;
}
