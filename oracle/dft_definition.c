/*
 * TEST INFRASTRUCTURE ONLY (see adhoc_oracle.c for the rules on who may use oracle/).
 *
 * Second, independent oracle: the reference's DEFINITION module restated in the working precision,
 *   src/Data/Array/Accelerate/Math/DFT/Roots.hs:26-36  rootsOfUnity:        roots[i] = cos k :+ (-sin k),  k = 2*pi*i/n
 *   src/Data/Array/Accelerate/Math/DFT/Roots.hs:41-51  inverseRootsOfUnity: roots[i] = cos k :+   sin k
 *   src/Data/Array/Accelerate/Math/DFT.hs:68-92        dftG:  X[k] = fold (+) 0 (zipWith (*) arr roots[(k*n) `mod` l])
 *   src/Data/Array/Accelerate/Math/DFT.hs:42-45        dft  = dftG (rootsOfUnity sh)
 *   src/Data/Array/Accelerate/Math/DFT.hs:50-59        idft = map (/ n) . dftG (inverseRootsOfUnity sh)
 * every operation in binary32 for Complex Float and binary64 for Complex Double, the angle formed as the reference forms it
 * (((2*pi)*i)/n in the working type), the complex product as Data.Complex defines it, the sum taken left to right (Accelerate's
 * fold order is unspecified; the tests allow for the reordering).  It shares no code and no algorithm with adhoc_impl.inc
 * (split radix / mixed radix / Bluestein): tests/test_oracle.py checks the two against each other, so an error common to both
 * would have to be made twice, in two different algorithms.  O(n^2), for pinning at n <= a few thousand.
 *
 * PARITY PINNING STATUS: like the rest of oracle/, unpinned by reference-held vectors (the reference has none).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#define DEFINE_DFT(SUF, REAL, COS, SIN, PI)                                                              \
  void dft_definition_##SUF(int inverse, size_t batch, size_t n, const REAL *in, REAL *out) {           \
    REAL *roots = (REAL *)malloc(2 * n * sizeof(REAL));                                                  \
    const REAL nn = (REAL)n;                                                                             \
    for (size_t i = 0; i < n; i++) {                                                                     \
      const REAL k = (REAL)2 * PI * (REAL)i / nn; /* Roots.hs:33,48: 2 * pi * i / n */                   \
      roots[2 * i] = COS(k);                                                                             \
      roots[2 * i + 1] = inverse ? SIN(k) : -SIN(k);                                                     \
    }                                                                                                    \
    for (size_t b = 0; b < batch; b++) {                                                                 \
      const REAL *x = in + 2 * b * n;                                                                    \
      REAL *y = out + 2 * b * n;                                                                         \
      for (size_t k = 0; k < n; k++) {                                                                   \
        REAL sr = 0, si = 0; /* DFT.hs:76: A.fold (+) 0 */                                               \
        for (size_t j = 0; j < n; j++) {                                                                 \
          const size_t r = (k * j) % n; /* DFT.hs:92: (k*n) `mod` l */                                   \
          const REAL wr = roots[2 * r], wi = roots[2 * r + 1];                                           \
          const REAL xr = x[2 * j], xi = x[2 * j + 1];                                                   \
          sr += xr * wr - xi * wi; /* Data.Complex (*) */                                                \
          si += xr * wi + xi * wr;                                                                       \
        }                                                                                                \
        if (inverse) { /* DFT.hs:57-59: A.map (/scale), scale = n :+ 0 -- complex division by a real */ \
          sr = sr / nn;                                                                                  \
          si = si / nn;                                                                                  \
        }                                                                                                \
        y[2 * k] = sr;                                                                                   \
        y[2 * k + 1] = si;                                                                               \
      }                                                                                                  \
    }                                                                                                    \
    free(roots);                                                                                         \
  }

DEFINE_DFT(f32, float, cosf, sinf, 3.14159265358979323846f)
DEFINE_DFT(f64, double, cos, sin, 3.14159265358979323846)
