// Complex helpers and fully-unrolled in-register DFTs of size 2..32 (forward sign, e^{-2 pi i jk/R}).
// The inverse direction never needs its own butterflies: IDFT(x) = swap(DFT(swap(x))) where swap
// exchanges re and im, so kernels swap on load / store instead (see fft_kernel.cuh).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include <utility>

namespace b200fft {

template <typename T> struct cpx_of;
template <> struct cpx_of<float> { using type = float2; };
template <> struct cpx_of<double> { using type = double2; };
template <typename T> using cpx_t = typename cpx_of<T>::type;

template <typename C> using real_of = decltype(C::x);

// compile-time loop: static_for<0,N>([&](auto ic){ constexpr int i = ic; ... });
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(static_cast<F&&>(f));
  }
}

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { return C{a.x + b.x, a.y + b.y}; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { return C{a.x - b.x, a.y - b.y}; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  return C{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename C> __device__ __forceinline__ C mul_mi(C a) { return C{a.y, -a.x}; }  // a * (-i)
template <typename C> __device__ __forceinline__ C cswap(C a) { return C{a.y, a.x}; }

// Blackwell packed FP32: add/sub/mul/fma.f32x2 work on a (lo, hi) register pair -- one issue slot for both halves
// of a complex number -- and ptxas folds half swaps and per-half negations into the operand modifiers
// (FADD2 / FFMA2 Ra.F32x2.LO_HI.NP ...), so -i*a and conj cost nothing.  The c64 kernels are bound by
// instruction issue, not by the FP32 pipe (profiles/r01_sass_mix_*.txt), hence every complex add and twiddle
// multiply below goes through these; c128 keeps the scalar DADD/DFMA forms.
#ifndef B200FFT_NO_F32X2
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return upk2(add2(pk2(a.x, a.y), pk2(b.x, b.y))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return upk2(sub2(pk2(a.x, a.y), pk2(b.x, b.y))); }
// (a.x w.x - a.y w.y, a.x w.y + a.y w.x) = a * (w.x, w.x) + (a.y, a.x) * (-w.y, w.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return upk2(fma2(pk2(a.x, a.y), pk2(w.x, w.x), mul2(pk2(a.y, a.x), pk2(-w.y, w.y))));
}
// a + (-i) b and a - (-i) b: the two odd outputs of a radix-4 butterfly, one FFMA2 each
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return upk2(fma2(pk2(b.y, b.x), pk2(1.f, -1.f), pk2(a.x, a.y))); }
__device__ __forceinline__ float2 csub_mi(float2 a, float2 b) { return upk2(fma2(pk2(b.y, b.x), pk2(-1.f, 1.f), pk2(a.x, a.y))); }
// a * h, or (-i a) * h when rot
__device__ __forceinline__ float2 cscale(float2 a, float h, bool rot = false) {
  return rot ? upk2(mul2(pk2(a.y, a.x), pk2(h, -h))) : upk2(mul2(pk2(a.x, a.y), pk2(h, h)));
}
#endif
template <typename C> __device__ __forceinline__ C cadd_mi(C a, C b) { return C{a.x + b.y, a.y - b.x}; }
template <typename C> __device__ __forceinline__ C csub_mi(C a, C b) { return C{a.x - b.y, a.y + b.x}; }
template <typename C, typename T> __device__ __forceinline__ C cscale(C a, T h, bool rot = false) {
  return rot ? C{a.y * h, -a.x * h} : C{a.x * h, a.y * h};
}

// bulk data is touched once per launch: evict-first loads / stores keep the reused twiddle tables cached
__device__ __forceinline__ float2 ld_stream(const float2* p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double2* p, double2 v) { __stcs(p, v); }

// cos(2 pi m / 32), m = 0..8  (first quadrant incl. end points), correctly rounded doubles
__device__ __forceinline__ constexpr double cos32(int m) {
  constexpr double t[9] = {1.0,
                           0.98078528040323044912618223613424,
                           0.92387953251128675612818318939679,
                           0.83146961230254523707878837761791,
                           0.70710678118654752440084436210485,
                           0.55557023301960222474283081394853,
                           0.38268343236508977172845998403040,
                           0.19509032201612826784828486847702,
                           0.0};
  return t[m];
}
// real/imag part of exp(-2 pi i m / 32) for any integer m >= 0
__device__ __forceinline__ constexpr double w32_re(int m) {
  m &= 31;
  if (m <= 8) return cos32(m);
  if (m <= 16) return -cos32(16 - m);
  if (m <= 24) return -cos32(m - 16);
  return cos32(32 - m);
}
__device__ __forceinline__ constexpr double w32_im(int m) {  // -sin(2 pi m/32) = re(m + 8)
  return w32_re(m + 8);
}

// multiply by exp(-2 pi i M / 32) with M a compile-time constant; trivial cases cost 0-2 flops
template <int M, typename C>
__device__ __forceinline__ C mul_w32(C a) {
  using T = real_of<C>;
  constexpr int m = M & 31;
  if constexpr (m == 0) return a;
  else if constexpr (m == 8) return C{a.y, -a.x};
  else if constexpr (m == 16) return C{-a.x, -a.y};
  else if constexpr (m == 24) return C{-a.y, a.x};
  else if constexpr (m == 4) { constexpr T h = (T)0.70710678118654752440084436210485; return cscale(cadd_mi(a, a), h); }     // (x+y, y-x) h
  else if constexpr (m == 12) { constexpr T h = (T)0.70710678118654752440084436210485; return cscale(cadd_mi(a, a), h, true); }
  else if constexpr (m == 20) { constexpr T h = (T)0.70710678118654752440084436210485; return cscale(cadd_mi(a, a), -h); }
  else if constexpr (m == 28) { constexpr T h = (T)0.70710678118654752440084436210485; return cscale(cadd_mi(a, a), -h, true); }
  else {
    constexpr T wr = (T)w32_re(m), wi = (T)w32_im(m);
    return cmul(a, C{wr, wi});
  }
}

// In-place forward DFT of R register-resident values, natural order in and out.
template <int R, typename C>
__device__ __forceinline__ void dft(C (&x)[R]) {
  static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "unsupported register radix");
  if constexpr (R == 2) {
    C a = x[0], b = x[1];
    x[0] = cadd(a, b); x[1] = csub(a, b);
  } else if constexpr (R == 4) {
    C t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    C t2 = cadd(x[1], x[3]), t3 = csub(x[1], x[3]);
    x[0] = cadd(t0, t2); x[1] = cadd_mi(t1, t3); x[2] = csub(t0, t2); x[3] = csub_mi(t1, t3);
  } else if constexpr (R == 8) {
    C e[4] = {x[0], x[2], x[4], x[6]};
    C o[4] = {x[1], x[3], x[5], x[7]};
    dft<4>(e); dft<4>(o);
    static_for<0, 4>([&](auto kc) {
      constexpr int k = kc;
      C t = mul_w32<4 * k>(o[k]);
      x[k] = cadd(e[k], t); x[k + 4] = csub(e[k], t);
    });
  } else if constexpr (R >= 16) {
    constexpr int M = R / 4;
    C s0[M], s1[M], s2[M], s3[M];
    static_for<0, M>([&](auto mc) {
      constexpr int m = mc;
      s0[m] = x[4 * m]; s1[m] = x[4 * m + 1]; s2[m] = x[4 * m + 2]; s3[m] = x[4 * m + 3];
    });
    dft<M>(s0); dft<M>(s1); dft<M>(s2); dft<M>(s3);
    static_for<0, M>([&](auto kc) {
      constexpr int k = kc;
      constexpr int u = 32 / R;  // w_R^m = w_32^(m*u)
      C a0 = s0[k];
      C a1 = mul_w32<u * k>(s1[k]);
      C a2 = mul_w32<2 * u * k>(s2[k]);
      C a3 = mul_w32<3 * u * k>(s3[k]);
      C t0 = cadd(a0, a2), t1 = csub(a0, a2);
      C t2 = cadd(a1, a3), t3 = csub(a1, a3);
      x[k] = cadd(t0, t2); x[k + M] = cadd_mi(t1, t3); x[k + 2 * M] = csub(t0, t2); x[k + 3 * M] = csub_mi(t1, t3);
    });
  }
}

}  // namespace b200fft
