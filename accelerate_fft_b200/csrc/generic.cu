// Non power-of-two axis lengths.  See generic.h.
#include "generic.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <functional>
#include <vector>

#include "../../include/b200fft.h"
#include "cplx.cuh"
#include "bluestein_kernel.cuh"

namespace b200fft {

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
template <typename C>
__global__ void copy_scale_kernel(const C* __restrict__ in, C* __restrict__ out, long long n, real_of<C> scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    C v = in[i];
    v.x *= scale; v.y *= scale;
    out[i] = v;
  }
}

// Bluestein step 1: A[l][m] = x_l[m] * chirp[m] (m < N), 0 (N <= m < M); line l0+l = (o, i)
template <typename C>
__global__ void chirp_pack_kernel(const C* __restrict__ in, C* __restrict__ A, const C* __restrict__ chirp, long long N,
                                  long long I, long long M, long long l0, long long nl, int swap_in) {
  const long long total = nl * M;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // consecutive threads -> consecutive lines when the axis is strided (coalesced reads), else consecutive m
    long long l, m;
    if (I > 1) { l = idx % nl; m = idx / nl; } else { m = idx % M; l = idx / M; }
    C v = C{0, 0};
    if (m < N) {
      const long long line = l0 + l, o = line / I, i = line % I;
      v = in[o * N * I + m * I + i];
      if (swap_in) v.y = -v.y;
      v = cmul(v, chirp[m]);
    }
    A[l * M + m] = v;
  }
}

template <typename C>
__global__ void filt_mul_kernel(C* __restrict__ B, const C* __restrict__ filt, long long M, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    B[idx] = cmul(B[idx], filt[idx % M]);
}

// Bluestein step 5: out_l[k] = A[l][k] * chirp[k] * scale
template <typename C>
__global__ void chirp_unpack_kernel(const C* __restrict__ A, C* __restrict__ out, const C* __restrict__ chirp, long long N,
                                    long long I, long long M, long long l0, long long nl, int swap_out, real_of<C> scale) {
  const long long total = nl * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long l, k;
    if (I > 1) { l = idx % nl; k = idx / nl; } else { k = idx % N; l = idx / N; }
    C v = cmul(A[l * M + k], chirp[k]);
    v.x *= scale; v.y *= scale;
    if (swap_out) v.y = -v.y;
    const long long line = l0 + l, o = line / I, i = line % I;
    out[o * N * I + k * I + i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// mixed-radix Stockham in shared memory, runtime radices (2,3,4,5,7,8 and generic odd primes <= 13)
// tile = TL lines x N points, ping-pong buffers; twiddles from a global table w_N^m.
// ------------------------------------------------------------------------------------------
struct MixedParams {
  long long O, N, I;
  int nstages;
  int radix[12];
  unsigned magic_ns[12];   // ceil(2^32 / Ns) of stage s: j / Ns == __umulhi(j, magic) for j < 2^16
  int TL, tpl_log2;        // lines per tile; threads per line = 2^tpl_log2
  int line_fast;           // 1: adjacent threads load adjacent lines (I > 1), 0: adjacent points (I == 1)
  int swap_in, swap_out;
};

// compile-time cos / sin of 2 pi m / R (Taylor series about the nearest multiple of pi/2; |x| <= pi/4, 13 terms: < 1e-17)
constexpr double ct_poly_cos(double x) {
  double term = 1, sum = 1;
  for (int i = 1; i <= 13; i++) {
    if (term < 1e-40 && term > -1e-40) break;   // (an underflow is not a constant expression)
    term *= -x * x / ((2 * i - 1) * (2 * i)); sum += term;
  }
  return sum;
}
constexpr double ct_poly_sin(double x) {
  double term = x, sum = x;
  for (int i = 1; i <= 13; i++) {
    if (term < 1e-40 && term > -1e-40) break;
    term *= -x * x / ((2 * i) * (2 * i + 1)); sum += term;
  }
  return sum;
}
constexpr double ct_cos_turn(int m, int R) {   // cos(2 pi m / R), m in [0, R)
  const double pi = 3.141592653589793238462643383279502884;
  double x = 2 * pi * m / R;                                    // [0, 2 pi)
  int quad = 0;
  while (x > pi / 4) { x -= pi / 2; quad++; }
  switch (quad & 3) {
    case 0: return ct_poly_cos(x);
    case 1: return -ct_poly_sin(x);
    case 2: return -ct_poly_cos(x);
    default: return ct_poly_sin(x);
  }
}
constexpr double ct_sin_turn(int m, int R) { return ct_cos_turn((4 * m + 3 * R) % (4 * R), 4 * R); }   // sin(t) = cos(t - 1/4 turn)

__device__ __forceinline__ int mpad(int i) { return i + (i >> 5); }   // one pad element per 32: de-phases the Stockham strides

template <int R, typename C>
__device__ __forceinline__ void small_dft(C* a);

// composite radix R = RA*RB inside one thread (n = RB*n1 + n2, k = k1 + RA*k2): RB transforms of RA points, the
// twiddles w_R^(k1*n2) as compile-time constants, RA transforms of RB points -- two Stockham stages without the
// shared-memory round trip between them
template <int RA, int RB, typename C>
__device__ __forceinline__ void comp_dft(C* a) {
  using T = real_of<C>;
  constexpr int R = RA * RB;
  C y[R];
  static_for<0, RB>([&](auto nc) {
    constexpr int n2 = nc;
    C t[RA];
    static_for<0, RA>([&](auto ic) { constexpr int n1 = ic; t[n1] = a[RB * n1 + n2]; });
    small_dft<RA>(t);
    static_for<0, RA>([&](auto kc) {
      constexpr int k1 = kc;
      constexpr int m = (k1 * n2) % R;
      if constexpr (m == 0) y[k1 * RB + n2] = t[k1];
      else {
        constexpr T c = (T)ct_cos_turn(m, R), sn = (T)(-ct_sin_turn(m, R));
        y[k1 * RB + n2] = cmul(t[k1], C{c, sn});
      }
    });
  });
  static_for<0, RA>([&](auto kc) {
    constexpr int k1 = kc;
    C u[RB];
    static_for<0, RB>([&](auto nc) { constexpr int n2 = nc; u[n2] = y[k1 * RB + n2]; });
    small_dft<RB>(u);
    static_for<0, RB>([&](auto qc) { constexpr int k2 = qc; a[k1 + RA * k2] = u[k2]; });
  });
}

template <int R, typename C>
__device__ __forceinline__ void small_dft(C* a) {
  using T = real_of<C>;
  if constexpr (R == 6) comp_dft<2, 3>(a);
  else if constexpr (R == 9) comp_dft<3, 3>(a);
  else if constexpr (R == 10) comp_dft<2, 5>(a);
  else if constexpr (R == 12) comp_dft<4, 3>(a);
  else if constexpr (R == 14) comp_dft<2, 7>(a);
  else if constexpr (R == 15) comp_dft<3, 5>(a);
  else if constexpr (R == 18) comp_dft<2, 9>(a);
  else if constexpr (R == 20) comp_dft<4, 5>(a);
  else if constexpr (R == 21) comp_dft<3, 7>(a);
  else if constexpr (R == 24) comp_dft<8, 3>(a);
  else if constexpr (R == 25) comp_dft<5, 5>(a);
  else if constexpr (R == 27) comp_dft<3, 9>(a);
  else if constexpr (R == 28) comp_dft<4, 7>(a);
  else if constexpr (R == 30) comp_dft<5, 6>(a);
  else if constexpr (R == 2) {
    C x = a[0], y = a[1];
    a[0] = cadd(x, y); a[1] = csub(x, y);
  } else if constexpr (R == 4) {
    C t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = mul_mi(csub(a[1], a[3]));
    a[0] = cadd(t0, t2); a[1] = cadd(t1, t3); a[2] = csub(t0, t2); a[3] = csub(t1, t3);
  } else if constexpr (R == 3) {
    const T c = (T)-0.5, s = (T)-0.86602540378443864676372317075294;  // exp(-2 pi i/3) = c + i s
    C t = cadd(a[1], a[2]), d = csub(a[1], a[2]);
    C m = C{a[0].x + c * t.x, a[0].y + c * t.y};
    C r = C{-s * d.y, s * d.x};  // i*s*d
    a[0] = cadd(a[0], t); a[1] = cadd(m, r); a[2] = csub(m, r);
  } else if constexpr (R == 5) {
    const T c1 = (T)0.30901699437494742410229341718282, c2 = (T)-0.80901699437494742410229341718282;
    const T s1 = (T)-0.95105651629515357211643933337938, s2 = (T)-0.58778525229247312916870595463907;
    C t1 = cadd(a[1], a[4]), d1 = csub(a[1], a[4]), t2 = cadd(a[2], a[3]), d2 = csub(a[2], a[3]);
    C m1 = C{a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y};
    C m2 = C{a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y};
    C r1 = C{-(s1 * d1.y + s2 * d2.y), s1 * d1.x + s2 * d2.x};   // i*(s1 d1 + s2 d2)
    C r2 = C{-(s2 * d1.y - s1 * d2.y), s2 * d1.x - s1 * d2.x};   // i*(s2 d1 - s1 d2)
    a[0] = cadd(a[0], cadd(t1, t2));
    a[1] = cadd(m1, r1); a[4] = csub(m1, r1); a[2] = cadd(m2, r2); a[3] = csub(m2, r2);
  } else if constexpr (R == 8 || R == 16 || R == 32) {
    C v[R];
    static_for<0, R>([&](auto ic) { constexpr int i = ic; v[i] = a[i]; });
    dft<R>(v);                                             // cplx.cuh: radix-4 recursion, compile-time constants
    static_for<0, R>([&](auto ic) { constexpr int i = ic; a[i] = v[i]; });
  } else {
    // odd primes 7 ... 31: pairs (r, R-r) share cos / sin sums -- (R-1)^2 / 2 real multiplies, all compile-time roots
    constexpr int H = (R - 1) / 2;
    C t[H], d[H];
    static_for<0, H>([&](auto ic) { constexpr int i = ic; t[i] = cadd(a[i + 1], a[R - 1 - i]); d[i] = csub(a[i + 1], a[R - 1 - i]); });
    C sum = a[0];
    static_for<0, H>([&](auto ic) { constexpr int i = ic; sum = cadd(sum, t[i]); });
    static_for<1, H + 1>([&](auto qc) {
      constexpr int q = qc;
      T mx = a[0].x, my = a[0].y, rx = 0, ry = 0;
      static_for<0, H>([&](auto ic) {
        constexpr int i = ic;
        constexpr int m = (q * (i + 1)) % R;
        constexpr T c = (T)ct_cos_turn(m, R), sn = (T)(-ct_sin_turn(m, R));
        mx += c * t[i].x; my += c * t[i].y;
        rx -= sn * d[i].y; ry += sn * d[i].x;       // i * sn * d
      });
      a[q] = C{mx + rx, my + ry};
      a[R - q] = C{mx - rx, my - ry};
    });
    a[0] = sum;
  }
}

// one Stockham stage of radix R on this thread's line: butterflies j = tl, tl + TPL, ... < N/R.
// gin != nullptr: the first stage of a contiguous line reads global memory directly (unit stride across the threads of
// the line, conjugating for the inverse); gout != nullptr: the last stage writes global memory directly (scale,
// conjugate) -- two shared-memory round trips and two barriers fewer per transform.
template <int R, typename C>
__device__ __noinline__ void mixed_stage(const C* __restrict__ src, C* __restrict__ dst, const C* __restrict__ gin, C* __restrict__ gout,
                                         const C* __restrict__ tw, int N, int Ns, unsigned magic, int tl, int tpl, int swap_in,
                                         int swap_out, real_of<C> scale) {
  using T = real_of<C>;
  const int nb = N / R;
  const int tmul = nb / Ns;                       // w_{Ns R}^{r k} = w_N^{r k N/(Ns R)}; r*k*tmul < N: no reduction needed
  for (int j = tl; j < nb; j += tpl) {
    const int jb = Ns > 1 ? (int)__umulhi((unsigned)j, magic) : j, k = j - jb * Ns;
    C a[R];
    if (gin) {
      static_for<0, R>([&](auto rc) { constexpr int r = rc; a[r] = gin[j + r * nb]; });
      if (swap_in) static_for<0, R>([&](auto rc) { constexpr int r = rc; a[r].y = -a[r].y; });
    } else {
      static_for<0, R>([&](auto rc) { constexpr int r = rc; a[r] = src[mpad(j + r * nb)]; });
    }
    if (Ns > 1) {
      const int base = tmul * k;
      static_for<1, R>([&](auto rc) { constexpr int r = rc; a[r] = cmul(a[r], __ldg(tw + base * r)); });
    }
    small_dft<R>(a);
    const int o = jb * Ns * R + k;
    if (gout) {
      const T sy = swap_out ? -scale : scale;
      static_for<0, R>([&](auto qc) { constexpr int q = qc; gout[o + q * Ns] = C{a[q].x * scale, a[q].y * sy}; });
    } else {
      static_for<0, R>([&](auto qc) { constexpr int q = qc; dst[mpad(o + q * Ns)] = a[q]; });
    }
  }
}

// BIGP: the instance that also holds the prime radices 17 ... 31 (their butterflies need more registers: a separate kernel
// instance, so the register allocation -- and with it the occupancy -- of every other length stays what it was)
// BIGP = 2 (c64 only): also the primes 37 ... 61, in CTAs of at most 256 threads (255 registers per thread)
template <typename C, int BIGP>
__device__ __forceinline__ void mixed_stage_any(int R, const C* src, C* dst, const C* gin, C* gout, const C* tw, int N, int Ns,
                                                unsigned magic, int tl, int tpl, int swap_in, int swap_out, real_of<C> scale) {
#define MS_CASE(r) case r: mixed_stage<r>(src, dst, gin, gout, tw, N, Ns, magic, tl, tpl, swap_in, swap_out, scale); break;
  if constexpr (sizeof(C) == 8) {
    switch (R) {
      MS_CASE(2) MS_CASE(3) MS_CASE(4) MS_CASE(5) MS_CASE(6) MS_CASE(7) MS_CASE(8) MS_CASE(9) MS_CASE(10) MS_CASE(11) MS_CASE(12) MS_CASE(13)
      MS_CASE(14) MS_CASE(15) MS_CASE(16) MS_CASE(18) MS_CASE(20) MS_CASE(21) MS_CASE(24) MS_CASE(25) MS_CASE(27) MS_CASE(28) MS_CASE(30)
      MS_CASE(32)
      default:
        if constexpr (BIGP >= 1) {
          switch (R) {
            MS_CASE(17) MS_CASE(19) MS_CASE(23) MS_CASE(29) MS_CASE(31)      // prime radices: see factor_small
            default:
              if constexpr (BIGP >= 2) {
                switch (R) {
                  MS_CASE(37) MS_CASE(41) MS_CASE(43) MS_CASE(47) MS_CASE(53) MS_CASE(59) MS_CASE(61)
                  default: break;
                }
              }
              break;
          }
        }
        break;
    }
  } else {   // c128: at most 16 values per thread -- and the primes 17, 19, 23, whose pair sums fit the 255 registers of a 256-thread CTA
    switch (R) {
      MS_CASE(2) MS_CASE(3) MS_CASE(4) MS_CASE(5) MS_CASE(6) MS_CASE(7) MS_CASE(8) MS_CASE(9) MS_CASE(10) MS_CASE(11) MS_CASE(12) MS_CASE(13)
      MS_CASE(14) MS_CASE(15) MS_CASE(16)
      default:
        if constexpr (BIGP >= 1) {
          switch (R) {
            MS_CASE(17) MS_CASE(19) MS_CASE(23)
            default: break;
          }
        }
        break;
    }
  }
#undef MS_CASE
}

template <typename C, int BIGP = 0>
__global__ void __launch_bounds__((sizeof(C) == 8 && BIGP < 2) ? 512 : 256)
mixed_radix_kernel(const MixedParams p, const C* __restrict__ in, C* __restrict__ out, const C* __restrict__ tw, real_of<C> scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = (int)p.N, TL = p.TL, tpl = 1 << p.tpl_log2;
  const int pitch = mpad(N - 1) + 1;
  C* buf0 = reinterpret_cast<C*>(smem_raw);
  C* buf1 = buf0 + (size_t)pitch * TL;
  const long long nlines = p.O * p.I;
  const long long l0 = (long long)blockIdx.x * TL;
  const int tid = threadIdx.x;
  const int l = tid >> p.tpl_log2, tl = tid & (tpl - 1);     // compute mapping: one thread group per line

  if (!p.line_fast) {
    // contiguous lines: first stage straight from global memory, last stage straight to it
    const bool live = l0 + l < nlines;      // (a dead group still meets the barriers)
    const C* gin = in + (l0 + l) * p.N;
    C* gout = out + (l0 + l) * p.N;
    C* src = buf0 + l * pitch;
    C* dst = buf1 + l * pitch;
    int Ns = 1;
    for (int s = 0; s < p.nstages; s++) {
      const int R = p.radix[s];
      const bool first = s == 0, last = s + 1 == p.nstages;
      if (live) {
        mixed_stage_any<C, BIGP>(R, (const C*)src, dst, first ? gin : (const C*)nullptr, last ? gout : (C*)nullptr, tw, N, Ns, p.magic_ns[s], tl, tpl,
                        p.swap_in, p.swap_out, scale);
      }
      if (!last) __syncthreads();
      if (!first) { C* t = src; src = dst; dst = t; }   // the first stage wrote buf1 = `dst`; from then on ping-pong
      else { src = buf1 + l * pitch; dst = buf0 + l * pitch; }
      Ns *= R;
    }
    return;
  }

  // strided axis: adjacent threads = adjacent lines, the tile's TL lines give TL*sizeof(C) contiguous bytes per point
  {
    const int ll = tid % TL, n0 = tid / TL, nstep = (int)blockDim.x / TL;
    const long long line = l0 + ll;
    if (line < nlines) {
      const long long o = line / p.I, i = line - o * p.I;
      const C* ip = in + o * p.N * p.I + i;
      C* bp = buf0 + ll * pitch;
      for (int n = n0; n < N; n += nstep) {
        C v = ip[(long long)n * p.I];
        if (p.swap_in) v.y = -v.y;
        bp[mpad(n)] = v;
      }
    }
  }
  __syncthreads();
  C* src = buf0 + l * pitch;
  C* dst = buf1 + l * pitch;
  int Ns = 1;
  for (int s = 0; s < p.nstages; s++) {
    mixed_stage_any<C, BIGP>(p.radix[s], (const C*)src, dst, (const C*)nullptr, (C*)nullptr, tw, N, Ns, p.magic_ns[s], tl, tpl, 0, 0, scale);
    __syncthreads();
    C* t = src; src = dst; dst = t;
    Ns *= p.radix[s];
  }
  {
    const C* res = (p.nstages & 1) ? buf1 : buf0;
    const int ll = tid % TL, n0 = tid / TL, nstep = (int)blockDim.x / TL;
    const long long line = l0 + ll;
    if (line < nlines) {
      const long long o = line / p.I, i = line - o * p.I;
      C* op = out + o * p.N * p.I + i;
      const C* bp = res + ll * pitch;
      const real_of<C> sy = p.swap_out ? -scale : scale;
      for (int n = n0; n < N; n += nstep) {
        C v = bp[mpad(n)];
        v.x *= scale; v.y *= sy;
        op[(long long)n * p.I] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// lengths up to 32 that are not powers of two (3, 5, 6, 7, 9, ... 31, primes included): one THREAD per line, direct
// O(N^2) sum from shared memory ([point][line] layout: conflict-free; root table broadcast).  A thread group per
// 12-point line wasted most of the machine (11 % of HBM peak) and the primes 17..31 went through Bluestein.
// ------------------------------------------------------------------------------------------
struct TinyParams {
  long long O, N, I, lines;
  unsigned magic;          // ceil(2^20 / N)
  int swap_in, swap_out;
};

template <typename C, int LPB>
__global__ void __launch_bounds__(LPB) tiny_dft_kernel(const TinyParams p, const C* __restrict__ in, C* __restrict__ out,
                                                       const C* __restrict__ tw, real_of<C> scale) {
  using T = real_of<C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = (int)p.N, t = threadIdx.x;
  C* xs = reinterpret_cast<C*>(smem_raw);          // [n][LPB]
  C* ys = xs + (size_t)N * LPB;                     // [k][LPB]
  C* ws = ys + (size_t)N * LPB;                     // [N] roots w_N^m
  const long long l0 = (long long)blockIdx.x * LPB;
  const int nl = (int)((p.lines - l0 < LPB) ? p.lines - l0 : LPB);
  if (t < N) ws[t] = tw[t];
  if (p.I == 1) {   // rows: the tile is one contiguous chunk of nl*N elements
    const C* ip = in + l0 * N;
    for (int idx = t; idx < nl * N; idx += LPB) {
      const int line = (int)(((unsigned)idx * p.magic) >> 20), n = idx - line * N;
      C v = ip[idx];
      if (p.swap_in) v.y = -v.y;
      xs[n * LPB + line] = v;
    }
  } else if (t < nl) {   // strided axis: adjacent threads = adjacent lines
    const long long line = l0 + t, o = line / p.I, i = line - o * p.I;
    const C* ip = in + o * p.N * p.I + i;
    for (int n = 0; n < N; n++) {
      C v = ip[(long long)n * p.I];
      if (p.swap_in) v.y = -v.y;
      xs[n * LPB + t] = v;
    }
  }
  __syncthreads();
  if (t < nl) {
    const T sy = p.swap_out ? -scale : scale;
    for (int k = 0; k < N; k++) {
      C acc = xs[t];
      int m = 0;
      for (int n = 1; n < N; n++) {
        m += k;
        if (m >= N) m -= N;
        acc = cadd(acc, cmul(xs[n * LPB + t], ws[m]));
      }
      ys[k * LPB + t] = C{acc.x * scale, acc.y * sy};
    }
  }
  __syncthreads();
  if (p.I == 1) {
    C* op = out + l0 * N;
    for (int idx = t; idx < nl * N; idx += LPB) {
      const int line = (int)(((unsigned)idx * p.magic) >> 20), k = idx - line * N;
      op[idx] = ys[k * LPB + line];
    }
  } else if (t < nl) {
    const long long line = l0 + t, o = line / p.I, i = line - o * p.I;
    C* op = out + o * p.N * p.I + i;
    for (int k = 0; k < N; k++) op[(long long)k * p.I] = ys[k * LPB + t];
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static const size_t kMixedSmemCap = 160 * 1024;
static const size_t kBluesteinWorkspaceCap = 512ull << 20;  // per buffer

// Fewest Stockham stages over the radices a thread can hold (c64: up to 32 values, c128: up to 16); among equal counts
// the most balanced product; odd radices first (the first stage has no twiddles and an odd store stride).
// Prime factors 17 ... 61 (c128: 17, 19, 23) are radices of their own -- the pair-sum butterfly of small_dft costs ~R/2
// multiply-adds per point, far below what the line's HBM traffic leaves room for -- so a length like 34, 323 or 961 takes the
// one-pass mixed-radix kernel instead of Bluestein's two 4N-point transforms (191 of the 992 lengths in [33, 1024], the range
// the reference's suite draws from, test/Test/Base.hs:44-45).  B200FFT_MAX_PRIME=13 restores the round-1 behaviour.
static int max_prime_radix(int is_double) {
  int cap = is_double ? 23 : 61;
  if (const char* e = getenv("B200FFT_MAX_PRIME")) { const int v = atoi(e); if (v >= 13 && v < cap) cap = v; }
  return cap;
}
static bool factor_small(long long n, std::vector<int>* f, int rmax, int pmax) {
  static const int kRad[] = {61, 59, 53, 47, 43, 41, 37, 32, 31, 30, 29, 28, 27, 25, 24, 23, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
  auto big_prime = [](int r) { return r == 17 || r == 19 || r == 23 || r == 29 || r == 31 || r >= 37; };
  f->clear();
  {
    long long m = n;
    for (int pr : {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61}) {
      if (pr > 13 && pr > pmax) break;
      while (m % pr == 0) m /= pr;
    }
    if (m != 1) return false;
  }
  if (n >= (1LL << 20)) return false;
  // best[m] = (stages, largest radix, first radix): fewest stages, then the smallest largest radix (balanced stages
  // keep the butterflies per line high and the registers per thread low)
  struct Best { int stages, maxrad, rad; };
  std::vector<Best> memo((size_t)n + 1, Best{-1, 0, 0});
  std::function<Best(long long)> go = [&](long long m) -> Best {
    if (m == 1) return Best{0, 0, 0};
    if (memo[(size_t)m].stages >= 0) return memo[(size_t)m];
    Best best{1 << 20, 1 << 20, 0};
    for (int r : kRad) {
      if (big_prime(r) ? r > pmax : r > rmax) continue;
      if (m % r) continue;
      const Best c = go(m / r);
      if (c.stages >= (1 << 20)) continue;
      const int st = c.stages + 1, mx = c.maxrad > r ? c.maxrad : r;
      if (st < best.stages || (st == best.stages && mx < best.maxrad)) best = Best{st, mx, r};
    }
    memo[(size_t)m] = best;
    return best;
  };
  if (go(n).stages >= (1 << 20)) return false;
  for (long long m = n; m > 1; m /= memo[(size_t)m].rad) f->push_back(memo[(size_t)m].rad);
  std::stable_sort(f->begin(), f->end(), [](int a, int b) { return (a & 1) > (b & 1); });   // odd radices first
  return true;
}

template <typename T>
static void fill_roots(std::vector<T>& v, long long N) {
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (long long m = 0; m < N; m++) {
    long double a = two_pi * (long double)m / (long double)N;
    v[2 * m] = (T)cosl(a);
    v[2 * m + 1] = (T)(-sinl(a));
  }
}

int plan_generic_axis(int is_double, long long O, long long N, long long I, GenericPass* gp, const Uploader& up) {
  const size_t esz = is_double ? 16 : 8;
  gp->O = O; gp->N = N; gp->I = I; gp->lines = O * I;
  if (N <= 32) {   // (powers of two never get here)
    gp->bluestein = 0;
    gp->tiny = 1;
    gp->threads = is_double ? 64 : 128;
    gp->smem = ((size_t)2 * N * gp->threads + N) * esz;
    if (is_double) { std::vector<double> h(2 * (size_t)N); fill_roots(h, N); gp->tw = up(h.data(), h.size() * sizeof(double)); }
    else { std::vector<float> h(2 * (size_t)N); fill_roots(h, N); gp->tw = up(h.data(), h.size() * sizeof(float)); }
    if (!gp->tw) return B200FFT_ALLOC_FAILED;
    snprintf(gp->desc, sizeof gp->desc, "tiny direct N=%lld, one thread per line, %d lines per CTA smem=%zu (O=%lld I=%lld)", N, gp->threads,
             gp->smem, O, I);
    return 0;
  }
  std::vector<int> f;
  bool smooth = N < (1LL << 20) && factor_small(N, &f, is_double ? 16 : 32, max_prime_radix(is_double));
  if (smooth && f.size() > 2) {
    // a prime radix of 17 ... 31 pays in one- and two-stage plans (34: 21 -> 54 %, 323: 22 -> 45 %, 841: 26 -> 39 % of the HBM
    // roofline against Bluestein); with three stages it does not (986: 30 -> 30 %, 1023: 30 -> 25 %): those stay on Bluestein
    // (profiles/r02_non_pow2_prime_radix.txt)
    for (int r : f) if (r == 17 || r == 19 || r == 23 || r == 29 || r == 31 || r >= 37) smooth = false;
  }
  if (smooth && ((size_t)N + N / 32 + 1) * esz * 2 <= kMixedSmemCap && N < 65536 && f.size() <= 12) {
    gp->bluestein = 0;
    gp->nstages = (int)f.size();
    for (size_t i = 0; i < f.size(); i++) gp->radix[i] = f[i];
    // threads per line: the largest power of two not above the butterflies per line of the widest stage (so that no
    // stage leaves most of a line's threads idle); lines per CTA: enough for >= 128 threads, within ~64 KB of shared
    // memory for rows (3 CTAs per SM) and 8 lines (64-128 B runs) for strided axes
    const size_t line_bytes = ((size_t)N + N / 32 + 1) * esz * 2;
    int rbig = 2;
    for (int r : f) rbig = r > rbig ? r : rbig;
    int tl2 = 0;
    while ((2 << tl2) <= N / rbig && tl2 < 8) tl2++;
    const int tmax = (is_double || rbig > 32) ? 256 : 512;     // (the instance holding the primes 37 ... 61 is built for 256 threads)
    int want = I > 1 ? 8 : 4;
    while ((want << tl2) < 64) want *= 2;
    int TL = (int)((I > 1 ? kMixedSmemCap : (size_t)(64 * 1024)) / line_bytes);
    if (TL > want) TL = want;
    if (TL < 1) TL = 1;
    if ((long long)TL > gp->lines) TL = (int)gp->lines;
    while ((TL << tl2) > tmax && tl2 > 0) tl2--;
    while ((TL << tl2) > tmax && TL > 1) TL--;
    while ((TL << tl2) < 32) tl2++;
    gp->TL = TL;
    gp->smem = line_bytes * TL;
    gp->tpl_log2 = tl2;
    gp->threads = TL << tl2;
    if (is_double) { std::vector<double> h(2 * (size_t)N); fill_roots(h, N); gp->tw = up(h.data(), h.size() * sizeof(double)); }
    else { std::vector<float> h(2 * (size_t)N); fill_roots(h, N); gp->tw = up(h.data(), h.size() * sizeof(float)); }
    if (!gp->tw) return B200FFT_ALLOC_FAILED;
    snprintf(gp->desc, sizeof gp->desc, "mixed-radix N=%lld stages=%d TL=%d threads=%d smem=%zu (O=%lld I=%lld)", N, gp->nstages,
             gp->TL, gp->threads, gp->smem, O, I);
    return 0;
  }
  // ---- Bluestein -------------------------------------------------------------------------
  gp->bluestein = 1;
  long long M = 1;
  while (M < 2 * N - 1) M <<= 1;
  gp->M = M;
  long long chunk = (long long)(kBluesteinWorkspaceCap / ((size_t)M * esz));
  if (chunk < 1) chunk = 1;
  if (chunk > gp->lines) chunk = gp->lines;
  gp->chunk_lines = chunk;
  gp->workspace_bytes = 2 * (size_t)chunk * M * esz;
  // chirp[n] = exp(-i pi n^2 / N), n^2 reduced mod 2N exactly
  std::vector<double> hc(2 * (size_t)N), hb(2 * (size_t)M, 0.0);
  const long double pi = 3.141592653589793238462643383279502884L;
  for (long long n = 0; n < N; n++) {
    long long r = (long long)(((unsigned __int128)n * n) % (unsigned long long)(2 * N));
    long double a = pi * (long double)r / (long double)N;
    long double c = cosl(a), s = sinl(a);
    hc[2 * n] = (double)c; hc[2 * n + 1] = (double)(-s);
    // b = conj(chirp), wrapped
    hb[2 * n] = (double)c; hb[2 * n + 1] = (double)s;
    if (n > 0) { hb[2 * (M - n)] = (double)c; hb[2 * (M - n) + 1] = (double)s; }
  }
  // filter = FFT_M(b) / M, computed once on the device in double with our own engine
  b200fftHandle fp = nullptr;
  int e = b200fftPlanMany1d(&fp, M, 1, B200FFT_Z2Z);
  if (e) return e;
  void *db = nullptr, *df = nullptr;
  if (cudaMalloc(&db, (size_t)M * 16) != cudaSuccess || cudaMalloc(&df, (size_t)M * 16) != cudaSuccess) {
    if (db) cudaFree(db);
    b200fftDestroy(fp);
    return B200FFT_ALLOC_FAILED;
  }
  cudaMemcpy(db, hb.data(), (size_t)M * 16, cudaMemcpyHostToDevice);
  e = b200fftExecScaled(fp, db, df, B200FFT_FORWARD, 1.0 / (double)M, nullptr);
  std::vector<double> hf(2 * (size_t)M);
  cudaError_t ce = cudaMemcpy(hf.data(), df, (size_t)M * 16, cudaMemcpyDeviceToHost);
  cudaFree(db); cudaFree(df);
  b200fftDestroy(fp);
  if (e || ce != cudaSuccess) return e ? e : B200FFT_EXEC_FAILED;
  if (is_double) {
    gp->chirp = up(hc.data(), hc.size() * sizeof(double));
    gp->filt = up(hf.data(), hf.size() * sizeof(double));
  } else {
    std::vector<float> c32(hc.begin(), hc.end()), f32(hf.begin(), hf.end());
    gp->chirp = up(c32.data(), c32.size() * sizeof(float));
    gp->filt = up(f32.data(), f32.size() * sizeof(float));
  }
  if (!gp->chirp || !gp->filt) return B200FFT_ALLOC_FAILED;
  const BluesteinEntry* be = (I == 1 && !(getenv("B200FFT_BLUESTEIN_FUSED") && atoi(getenv("B200FFT_BLUESTEIN_FUSED")) == 0))
                                 ? find_bluestein(is_double, M) : nullptr;
  if (be) {
    // contiguous lines: pack, both transforms, filter and unpack in one launch, the lines never leave the SM
    std::vector<double> ht(2 * (size_t)(be->tw_len > 0 ? be->tw_len : 1), 0.0);
    {
      const long double two_pi = 6.283185307179586476925286766559005768L;
      size_t off = 0;
      long long Ns = be->rad[0];
      for (int s = 1; s < be->S; s++) {     // stage s >= 1: entry [(r-1)*Ns + k] = w_{Ns R}^(r k)  (as plan.cu make_stage_twiddles)
        const int R = be->rad[s];
        for (int r = 1; r < R; r++)
          for (long long kk = 0; kk < Ns; kk++) {
            const long double a = two_pi * (long double)((r * kk) % (Ns * R)) / (long double)(Ns * R);
            ht[2 * (off + (size_t)(r - 1) * Ns + kk)] = (double)cosl(a);
            ht[2 * (off + (size_t)(r - 1) * Ns + kk) + 1] = (double)(-sinl(a));
          }
        off += (size_t)(R - 1) * Ns;
        Ns *= R;
      }
    }
    if (is_double) gp->fused_tw = up(ht.data(), ht.size() * sizeof(double));
    else { std::vector<float> t32(ht.begin(), ht.end()); gp->fused_tw = up(t32.data(), t32.size() * sizeof(float)); }
    if (!gp->fused_tw) return B200FFT_ALLOC_FAILED;
    gp->fused = be->func; gp->fused_tl = be->TL; gp->fused_threads = be->threads; gp->fused_smem = be->smem;
    if (be->smem > 48 * 1024 && cudaFuncSetAttribute(be->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)be->smem) != cudaSuccess) {
      cudaGetLastError();
      return B200FFT_INTERNAL_ERROR;
    }
    gp->workspace_bytes = 0;
    snprintf(gp->desc, sizeof gp->desc, "bluestein N=%lld on M=%lld in one launch: chirp, FFT, filter, IFFT, chirp on chip; TL=%d threads=%d smem=%zu (O=%lld)",
             N, M, be->TL, be->threads, be->smem, O);
    return 0;
  }
  e = b200fftPlanMany1d(&gp->sub, M, chunk, is_double ? B200FFT_Z2Z : B200FFT_C2C);
  if (e) return e;
  snprintf(gp->desc, sizeof gp->desc, "bluestein N=%lld M=%lld chunk=%lld lines (%d sub-passes x2 + 3 pointwise) (O=%lld I=%lld)", N, M,
           chunk, b200fftNumPasses(gp->sub), O, I);
  return 0;
}

void destroy_generic(GenericPass* gp) {
  if (gp->sub) { b200fftDestroy(gp->sub); gp->sub = nullptr; }
}

template <typename C>
static cudaError_t launch_generic_t(const GenericPass& gp, const C* src, C* dst, void* workspace, int inverse, double scale,
                                    cudaStream_t stream, long long* nl) {
  using T = real_of<C>;
  if (gp.tiny) {
    TinyParams tp{};
    tp.O = gp.O; tp.N = gp.N; tp.I = gp.I; tp.lines = gp.lines;
    tp.magic = (unsigned)(((1u << 20) + gp.N - 1) / gp.N);
    tp.swap_in = inverse & 1; tp.swap_out = (inverse >> 1) & 1;
    constexpr int LPB = sizeof(C) == 8 ? 128 : 64;
    const long long tiles = (gp.lines + LPB - 1) / LPB;
    tiny_dft_kernel<C, LPB><<<(unsigned)tiles, LPB, gp.smem, stream>>>(tp, src, dst, (const C*)gp.tw, (T)scale);
    *nl += 1;
    return cudaGetLastError();
  }
  if (!gp.bluestein) {
    MixedParams mp{};
    mp.O = gp.O; mp.N = gp.N; mp.I = gp.I; mp.nstages = gp.nstages;
    for (int i = 0; i < gp.nstages; i++) mp.radix[i] = gp.radix[i];
    mp.TL = gp.TL; mp.tpl_log2 = gp.tpl_log2; mp.line_fast = gp.I > 1;
    {
      long long Ns = 1;
      for (int i = 0; i < gp.nstages; i++) { mp.magic_ns[i] = (unsigned)(((1ull << 32) + Ns - 1) / Ns); Ns *= gp.radix[i]; }
    }
    mp.swap_in = mp.swap_out = 0;
    // the caller folds first/last-pass information into `inverse`: see launch_generic
    mp.swap_in = inverse & 1; mp.swap_out = (inverse >> 1) & 1;
    const long long tiles = (gp.lines + gp.TL - 1) / gp.TL;
    int bigp = 0;
    for (int i = 0; i < gp.nstages; i++) {
      const int r = gp.radix[i];
      if (r > 32) bigp = 2;
      else if ((r == 17 || r == 19 || r == 23 || r == 29 || r == 31) && bigp < 1) bigp = 1;
    }
    if constexpr (sizeof(C) == 8) {
      if (bigp == 2) {
        mixed_radix_kernel<C, 2><<<(unsigned)tiles, gp.threads, gp.smem, stream>>>(mp, src, dst, (const C*)gp.tw, (T)scale);
        *nl += 1;
        return cudaGetLastError();
      }
    }
    if (bigp == 2) return cudaErrorInvalidValue;     // (the planner never gives c128 a radix above 23)
    if (bigp) mixed_radix_kernel<C, 1><<<(unsigned)tiles, gp.threads, gp.smem, stream>>>(mp, src, dst, (const C*)gp.tw, (T)scale);
    else mixed_radix_kernel<C, 0><<<(unsigned)tiles, gp.threads, gp.smem, stream>>>(mp, src, dst, (const C*)gp.tw, (T)scale);
    *nl += 1;
    return cudaGetLastError();
  }
  if (gp.fused) {
    const C* chirp = (const C*)gp.chirp; const C* filt = (const C*)gp.filt; const C* tws = (const C*)gp.fused_tw;
    int n = (int)gp.N, swap_in = inverse & 1, swap_out = (inverse >> 1) & 1;
    long long lines = gp.lines;
    T sc = (T)scale;
    void* args[] = {(void*)&src, (void*)&dst, (void*)&chirp, (void*)&filt, (void*)&tws, &n, &lines, &swap_in, &swap_out, &sc};
    const long long tiles = (gp.lines + gp.fused_tl - 1) / gp.fused_tl;
    *nl += 1;
    return cudaLaunchKernel(gp.fused, dim3((unsigned)tiles), dim3(gp.fused_threads), args, gp.fused_smem, stream);
  }
  const long long M = gp.M, N = gp.N;
  C* A = (C*)workspace;
  C* B = A + gp.chunk_lines * M;
  const int swap_in = inverse & 1, swap_out = (inverse >> 1) & 1;
  for (long long l0 = 0; l0 < gp.lines; l0 += gp.chunk_lines) {
    const long long nlc = (gp.lines - l0 < gp.chunk_lines) ? gp.lines - l0 : gp.chunk_lines;
    const int thr = 256;
    auto blocks = [&](long long total) { long long b = (total + thr - 1) / thr; return (unsigned)(b > 148 * 32 ? 148 * 32 : b); };
    chirp_pack_kernel<C><<<blocks(nlc * M), thr, 0, stream>>>(src, A, (const C*)gp.chirp, N, gp.I, M, l0, nlc, swap_in);
    int e = b200fftExec(gp.sub, A, B, B200FFT_FORWARD, stream);
    if (e) return cudaErrorUnknown;
    filt_mul_kernel<C><<<blocks(gp.chunk_lines * M), thr, 0, stream>>>(B, (const C*)gp.filt, M, gp.chunk_lines * M);
    e = b200fftExec(gp.sub, B, A, B200FFT_INVERSE, stream);
    if (e) return cudaErrorUnknown;
    chirp_unpack_kernel<C><<<blocks(nlc * N), thr, 0, stream>>>(A, dst, (const C*)gp.chirp, N, gp.I, M, l0, nlc, swap_out, (T)scale);
    *nl += 3;
  }
  return cudaGetLastError();
}

// `inverse` bit 0: conjugate on load (first pass of an inverse plan); bit 1: conjugate on store (last pass)
cudaError_t launch_generic(int is_double, const GenericPass& gp, const void* src, void* dst, void* workspace, int inverse,
                           double scale, cudaStream_t stream, long long* nl) {
  if (is_double) return launch_generic_t<double2>(gp, (const double2*)src, (double2*)dst, workspace, inverse, scale, stream, nl);
  return launch_generic_t<float2>(gp, (const float2*)src, (float2*)dst, workspace, inverse, scale, stream, nl);
}

cudaError_t launch_copy_scale(int is_double, const void* src, void* dst, long long count, double scale, cudaStream_t stream,
                              long long* nl) {
  long long b = (count + 255) / 256;
  unsigned blocks = (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
  if (is_double) copy_scale_kernel<double2><<<blocks, 256, 0, stream>>>((const double2*)src, (double2*)dst, count, scale);
  else copy_scale_kernel<float2><<<blocks, 256, 0, stream>>>((const float2*)src, (float2*)dst, count, (float)scale);
  *nl += 1;
  return cudaGetLastError();
}

// [dl][h][w] <-> [P][dl][h/P][w]; one 16-byte vector per thread-iteration
template <typename V, bool PACK>
__global__ void slab_pack_kernel(const V* __restrict__ src, V* __restrict__ dst, long long dl, long long h, long long wv, int P) {
  const long long hl = h / P, total = dl * h * wv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long x = idx % wv, y = (idx / wv) % h, z = idx / (wv * h);
    const long long r = y / hl, yl = y % hl;
    const long long packed = ((r * dl + z) * hl + yl) * wv + x;
    if (PACK) dst[packed] = src[idx]; else dst[idx] = src[packed];
  }
}

cudaError_t launch_slab_pack(int is_double, bool pack, const void* src, void* dst, long long dl, long long h, long long w, int P,
                             cudaStream_t stream) {
  // vector = 16 bytes = 2 c64 or 1 c128
  const long long wv = is_double ? w : w / 2;
  const bool vec_ok = is_double || (w % 2 == 0);
  const long long total = dl * h * (vec_ok ? wv : w);
  long long b = (total + 255) / 256;
  unsigned blocks = (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
  if (vec_ok) {
    if (pack) slab_pack_kernel<float4, true><<<blocks, 256, 0, stream>>>((const float4*)src, (float4*)dst, dl, h, wv, P);
    else slab_pack_kernel<float4, false><<<blocks, 256, 0, stream>>>((const float4*)src, (float4*)dst, dl, h, wv, P);
  } else {
    if (pack) slab_pack_kernel<float2, true><<<blocks, 256, 0, stream>>>((const float2*)src, (float2*)dst, dl, h, w, P);
    else slab_pack_kernel<float2, false><<<blocks, 256, 0, stream>>>((const float2*)src, (float2*)dst, dl, h, w, P);
  }
  return cudaGetLastError();
}

// DFT/Centre.hs:70-164 shift*/ishift*: backpermute, dst[z][y][x] = src[(z+sd)%d][(y+sh)%h][(x+sw)%w]
template <typename C>
__global__ void shift_kernel(const C* __restrict__ src, C* __restrict__ dst, long long d, long long h, long long w, long long sd,
                             long long sh, long long sw) {
  const long long total = d * h * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long x = idx % w, y = (idx / w) % h, z = idx / (w * h);
    dst[idx] = src[(((z + sd) % d) * h + (y + sh) % h) * w + (x + sw) % w];
  }
}
// DFT/Centre.hs:36-66 centre*: dst = (-1)^(z+y+x) * src
template <typename C>
__global__ void centre_kernel(const C* __restrict__ src, C* __restrict__ dst, long long d, long long h, long long w) {
  const long long total = d * h * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long x = idx % w, y = (idx / w) % h, z = idx / (w * h);
    C v = src[idx];
    if ((x + y + z) & 1) { v.x = -v.x; v.y = -v.y; }
    dst[idx] = v;
  }
}
static unsigned ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}
cudaError_t launch_shift(int is_double, const void* src, void* dst, long long d, long long h, long long w, long long sd, long long sh,
                         long long sw, cudaStream_t stream) {
  if (is_double) shift_kernel<double2><<<ew_blocks(d * h * w), 256, 0, stream>>>((const double2*)src, (double2*)dst, d, h, w, sd, sh, sw);
  else shift_kernel<float2><<<ew_blocks(d * h * w), 256, 0, stream>>>((const float2*)src, (float2*)dst, d, h, w, sd, sh, sw);
  return cudaGetLastError();
}
cudaError_t launch_centre(int is_double, const void* src, void* dst, long long d, long long h, long long w, cudaStream_t stream) {
  if (is_double) centre_kernel<double2><<<ew_blocks(d * h * w), 256, 0, stream>>>((const double2*)src, (double2*)dst, d, h, w);
  else centre_kernel<float2><<<ew_blocks(d * h * w), 256, 0, stream>>>((const float2*)src, (float2*)dst, d, h, w);
  return cudaGetLastError();
}

int generic_set_attrs() {
  cudaFuncSetAttribute(tiny_dft_kernel<float2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(tiny_dft_kernel<double2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (cudaFuncSetAttribute(mixed_radix_kernel<float2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixedSmemCap) != cudaSuccess ||
      cudaFuncSetAttribute(mixed_radix_kernel<float2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixedSmemCap) != cudaSuccess ||
      cudaFuncSetAttribute(mixed_radix_kernel<float2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixedSmemCap) != cudaSuccess)
    return B200FFT_INTERNAL_ERROR;
  if (cudaFuncSetAttribute(mixed_radix_kernel<double2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixedSmemCap) != cudaSuccess ||
      cudaFuncSetAttribute(mixed_radix_kernel<double2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixedSmemCap) != cudaSuccess)
    return B200FFT_INTERNAL_ERROR;
  return 0;
}

}  // namespace b200fft
