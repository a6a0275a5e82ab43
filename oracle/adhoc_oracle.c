/*
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, and there only as the checker or the CPU baseline.
 *
 * What this is: a plain-C restatement of the reference's pure-Accelerate FFT path
 * ("oracle #2": src/Data/Array/Accelerate/Math/FFT/Adhoc.hs + the 2D/3D compositions in
 * src/Data/Array/Accelerate/Math/FFT.hs), in binary32 and binary64, plus an exact-math
 * (x87 80-bit long double) DFT used to pin it.
 *
 * PARITY PINNING STATUS: the reference holds NO golden vectors or known-answer tests for
 * this path (only seven algebraic properties at ~5e-3 tolerance), and
 * the reference itself (Haskell; needs GHC + FFTW/cuFFT) cannot be built or run in this
 * image.  The restatement is therefore pinned against (a) the mathematical definition of the
 * DFT evaluated in long double (exact_dft_* below), (b) independent library FFTs (scipy
 * pocketfft, torch/MKL) standing in for the reference's FFTW path, (c) closed-form
 * known-answer vectors committed under tests/golden/, and (d) the reference's own seven
 * properties re-expressed in tests/.  By the task's definition that is "parity unpinned by
 * reference-held vectors"; DESIGN.md says the same.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF f32
#define COS cosf
#define SIN sinf
#define SQRT sqrtf
#include "adhoc_impl.inc"
#undef REAL
#undef SUF
#undef COS
#undef SIN
#undef SQRT
#undef FN

#define REAL double
#define SUF f64
#define COS cos
#define SIN sin
#define SQRT sqrt
#include "adhoc_impl.inc"
#undef REAL
#undef SUF
#undef COS
#undef SIN
#undef SQRT
#undef FN

/* ---------------------------------------------------------------------------------------
 * Exact-math definition:  X[k] = sum_j x[j] * exp(sign * 2*pi*i * j*k / n)
 * (Mode.hs:21-26 sign convention; DFT.hs:42-45 / DFT/Roots.hs:26-36 definition), evaluated
 * in long double with the exponent reduced exactly (j*k mod n in integers).
 * Input is double (float inputs are widened exactly by the caller); output long double.
 * O(n^2): for pinning at small n only.
 * ------------------------------------------------------------------------------------- */
void exact_dft(int sign, size_t batch, size_t n, const double *in, long double *out) {
  const long double two_pi = 6.283185307179586476925286766559005768L;
  long double *c = (long double *)malloc(sizeof(long double) * (n ? n : 1));
  long double *s = (long double *)malloc(sizeof(long double) * (n ? n : 1));
  for (size_t m = 0; m < n; m++) {
    long double a = two_pi * (long double)m / (long double)n;
    c[m] = cosl(a); s[m] = (long double)sign * sinl(a);
  }
#pragma omp parallel for schedule(static)
  for (long long b = 0; b < (long long)batch; b++)
    for (size_t k = 0; k < n; k++) {
      long double re = 0, im = 0;
      for (size_t j = 0; j < n; j++) {
        size_t m = (size_t)(((unsigned __int128)j * k) % n);
        long double xr = in[2 * (b * n + j)], xi = in[2 * (b * n + j) + 1];
        re += xr * c[m] - xi * s[m];
        im += xr * s[m] + xi * c[m];
      }
      out[2 * (b * n + k)] = re; out[2 * (b * n + k) + 1] = im;
    }
  free(c); free(s);
}

/* One output bin of a long strided transform, for sampled checks at sizes where a full
 * exact transform is too slow (2^28, 1024^3): X[k] over x[j*stride], j < n.  Kahan-free:
 * long double accumulation of <= 2^30 terms of magnitude <= 2 keeps ~1e-15 relative. */
void exact_bin(int sign, size_t n, size_t stride, size_t k, const double *in, long double *out2) {
  const long double two_pi = 6.283185307179586476925286766559005768L;
  long double re = 0, im = 0;
  for (size_t j = 0; j < n; j++) {
    size_t m = (size_t)(((unsigned __int128)j * k) % n);
    long double a = two_pi * (long double)m / (long double)n;
    long double c = cosl(a), s = (long double)sign * sinl(a);
    long double xr = in[2 * j * stride], xi = in[2 * j * stride + 1];
    re += xr * c - xi * s;
    im += xr * s + xi * c;
  }
  out2[0] = re; out2[1] = im;
}
