// Bluestein (chirp-z) for contiguous lines whose length N has a prime factor above 13, in ONE launch.
//
//   X[k] = conj(c[k]) ... in the usual form:  X[k] = w[k] * sum_n (x[n] w[n]) * conj(w)[k - n],   w[n] = exp(-i pi n^2 / N)
// i.e. a length-N linear convolution, done as a cyclic one of length M = 2^m >= 2N - 1.  The first version ran it as five
// passes over an M-padded workspace in HBM (pack * chirp, FFT_M, * filter, IFFT_M, * chirp + unpack): 18-34 x the algorithmic
// traffic, 4-8 % of the HBM roofline.  Here a CTA keeps its TL lines on chip for the whole chain: load N points (* chirp,
// zero padded to M in registers), the power-of-two Stockham stages of the line kernel (fft_kernel.cuh) forwards, the filter
// spectrum (M entries, L1/L2 resident, 1/M folded in), the same stages again for the inverse (conjugation trick), * chirp,
// store N points.  HBM traffic = one read + one write of the N-point lines: the algorithmic minimum.
//
// Replaces, for lengths cuFFT serves with its own Bluestein, cufftExecC2C / cufftExecZ2Z behind
// /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs:112-124 (the reference's suite draws n in [1,1024],
// /root/reference/test/Test/Base.hs:44-58).
#pragma once
#include "fft_kernel.cuh"

namespace b200fft {

// all register stages of K on lines held as v[e] = point (t + e*TPT) of line l, natural order in and out (row layout)
template <class K, typename C>
__device__ __forceinline__ void fft_stages_rows(C (&v)[K::E], C* sm, int l, int t, const C* __restrict__ tws) {
  run_stage<K, 0>(v, t, tws);
  if constexpr (K::S > 1) {
    scatter<K, 0, false>(v, sm, l, t);
    static_for<1, (K::S > 1 ? K::S - 1 : 1)>([&](auto sc) {
      constexpr int s = sc;
      __syncthreads();
      gather<K, false>(v, sm, l, t);
      run_stage<K, s>(v, t, tws);
      __syncthreads();
      scatter<K, s, false>(v, sm, l, t);
    });
    __syncthreads();
    gather<K, false>(v, sm, l, t);
    run_stage<K, K::S - 1>(v, t, tws);
  }
}

// in / out: [lines][N] contiguous; chirp: N entries exp(-i pi n^2 / N); filt: M entries FFT_M(wrapped conj chirp) / M
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MINB)
bluestein_rows_kernel(const cpx_t<typename K::real>* __restrict__ in, cpx_t<typename K::real>* __restrict__ out,
                      const cpx_t<typename K::real>* __restrict__ chirp, const cpx_t<typename K::real>* __restrict__ filt,
                      const cpx_t<typename K::real>* __restrict__ tws, int N, long long lines, int swap_in, int swap_out,
                      typename K::real scale) {
  using T = typename K::real;
  using C = cpx_t<T>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);
  const int tid = threadIdx.x;
  const int t = tid % K::TPT, l = tid / K::TPT;
  const long long line0 = (long long)blockIdx.x * K::TL + l;
  const bool valid = line0 < lines;
  const long long line = valid ? line0 : lines - 1;      // lines past the end re-read the last one; nothing is stored for them
  const C* ip = in + line * N;
  C v[K::E];
  static_for<0, K::E>([&](auto ec) {
    constexpr int e = ec;
    const int idx = t + e * K::TPT;
    C x = C{0, 0};
    if (idx < N) {
      x = ip[idx];
      if (swap_in) x.y = -x.y;
      x = cmul(x, __ldg(chirp + idx));
    }
    v[e] = x;
  });
  fft_stages_rows<K>(v, sm, l, t, tws);
  // spectrum * filter; conj so that the forward stages below compute the inverse transform (IDFT(y) = conj(DFT(conj(y))))
  static_for<0, K::E>([&](auto ec) {
    constexpr int e = ec;
    C y = cmul(v[e], __ldg(filt + t + e * K::TPT));
    y.y = -y.y;
    v[e] = y;
  });
  if constexpr (K::S > 1) __syncthreads();   // every gather of the forward transform is done before the exchange space is re-used
  fft_stages_rows<K>(v, sm, l, t, tws);
  C* op = out + line * N;
  const T sy = swap_out ? -scale : scale;
  static_for<0, K::E>([&](auto ec) {
    constexpr int e = ec;
    const int idx = t + e * K::TPT;
    if (valid && idx < N) {
      C y = v[e];
      y.y = -y.y;
      y = cmul(y, __ldg(chirp + idx));
      y.x *= scale; y.y *= sy;
      op[idx] = y;
    }
  });
}

struct BluesteinEntry {
  int is_double, M, TL, threads, S, rad[4], tw_len;
  size_t smem;
  const void* func;
};
const BluesteinEntry* find_bluestein(int is_double, long long M);

}  // namespace b200fft
