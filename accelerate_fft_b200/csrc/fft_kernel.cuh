// The power-of-two Stockham "line" kernel: one CTA transforms a tile of TL lines of length N that
// live anywhere in HBM (contiguous rows, strided columns, or rows-in / columns-out for the
// four-step transposition), keeping every butterfly in registers and exchanging between register
// stages through padded shared memory.  One HBM read + one HBM write per element per launch.
//
// Replaces, for its share of a plan, what cufftExecC2C / cufftExecZ2Z did behind
// /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs:112-124.
#pragma once
#include "cplx.cuh"

namespace b200fft {

// Where the lines of one launch live.  All strides in complex elements.
//   address(b, o, line, n) = b*bs + o*os + line*ls + n*ns
// b < nb (batch), o < no (outer), line < nl (tiled by TL), n < N (transform index).
struct Geom {
  long long ibs, ios, ils, ins;  // input
  long long obs, oos, ols, ons;  // output
  int nb, no, nl, ntl;           // ntl = ceil(nl / TL)
  // four-step twiddle  w_L^(k * (line / tw_div)),  L = 2^(tw_hi_bits_total): exponent split lo/hi
  int tw_div;
  int tw_lo_bits;
  int swap_in, swap_out;         // inverse = swap(fwd(swap(x)))
  int stream_hint;               // A/B switch: evict-first loads/stores for the bulk data
};

template <typename T_, int N_, int E_, int TL_, int MINB_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1>
struct Cfg {
  using real = T_;
  static constexpr int N = N_, E = E_, TL = TL_;
  static constexpr int TPT = N / E;            // threads per line
  static constexpr int THREADS = TPT * TL;
  // min resident CTAs per SM the register allocation must allow; 0 = derive from a register target of
  // 64 (c64 E<=16, c128 E<=8) or 128 (c64 E=32, c128 E=16) per thread
  static constexpr int TARGET_REGS = (sizeof(T_) == 4) ? (E_ <= 16 ? 64 : 128) : (E_ <= 8 ? 64 : 128);
  static constexpr int AUTO_MINB = 65536 / (TARGET_REGS * THREADS) < 1 ? 1 : (65536 / (TARGET_REGS * THREADS) > 16 ? 16 : 65536 / (TARGET_REGS * THREADS));
  static constexpr int MINB = MINB_ > 0 ? MINB_ : AUTO_MINB;
  static constexpr int rad[4] = {R0_, R1_, R2_, R3_};
  static constexpr int S = (R3_ > 1) ? 4 : (R2_ > 1) ? 3 : (R1_ > 1) ? 2 : 1;
  static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
  static_assert(N_ % E_ == 0 && E_ % R0_ == 0 && E_ % R1_ == 0 && E_ % R2_ == 0 && E_ % R3_ == 0, "bad E");
  static constexpr int ns(int s) { int p = 1; for (int i = 0; i < s; i++) p *= rad[i]; return p; }
  // stage twiddle table: for stage s >= 1, entries [(r-1)*Ns + k], r in 1..R-1, k < Ns
  static constexpr int tw_off(int s) { int o = 0; for (int i = 1; i < s; i++) o += (rad[i] - 1) * ns(i); return o; }
  static constexpr int TW_LEN = tw_off(S);
  static constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
  static constexpr int LOGQ = ilog2(R0_);      // pad one element every R0 elements (row layout)
  static constexpr int ESZ = 2 * (int)sizeof(T_);
  static constexpr int GROUP = 128 / ESZ;      // lanes served by one shared-memory wavefront
  static constexpr int padidx(int i) { return i + (i >> LOGQ); }
  // row layout: line pitch == 1 (mod GROUP) so that a column-mapped reader is conflict free too
  static constexpr int PITCH0 = padidx(N - 1) + 1;
  static constexpr int PITCH = (TL == 1) ? PITCH0 : ((PITCH0 + GROUP - 2) / GROUP) * GROUP + 1;
  // column layout [i][line] with an extra skew of XPAD per R0 block when a wavefront spans >1 i
  static constexpr int XPAD = (TL < GROUP) ? TL : 0;
  static constexpr int COL_ELEMS = N * TL + ((N - 1) >> LOGQ) * XPAD + TL;
  static constexpr int ROW_ELEMS = PITCH * TL;
  template <bool COL> static constexpr size_t smem_bytes() {
    return (S > 1) ? (size_t)(COL ? COL_ELEMS : ROW_ELEMS) * ESZ : 0;
  }
};

template <class K, bool COL>
__device__ __forceinline__ int smem_addr(int l, int i) {
  if constexpr (COL) return i * K::TL + (i >> K::LOGQ) * K::XPAD + l;
  else return l * K::PITCH + i + (i >> K::LOGQ);
}

// One register stage: E/R butterflies of radix R per thread.
template <class K, int s, typename C>
__device__ __forceinline__ void run_stage(C (&v)[K::E], int t, const C* __restrict__ tws) {
  constexpr int R = K::rad[s], B = K::E / R, Ns = K::ns(s);
  static_for<0, B>([&](auto bc) {
    constexpr int b = bc;
    C a[R];
    static_for<0, R>([&](auto rc) { constexpr int r = rc; a[r] = v[b + r * B]; });
    if constexpr (s > 0) {
      const int k = (t + b * K::TPT) & (Ns - 1);
      const C* tp = tws + K::tw_off(s) + k;
      static_for<1, R>([&](auto rc) {
        constexpr int r = rc;
        a[r] = cmul(a[r], __ldg(tp + (r - 1) * Ns));
      });
    }
    dft<R>(a);
    static_for<0, R>([&](auto rc) { constexpr int r = rc; v[b + r * B] = a[r]; });
  });
}

// Stockham scatter of stage s results into shared memory.
template <class K, int s, bool COL, typename C>
__device__ __forceinline__ void scatter(const C (&v)[K::E], C* sm, int l, int t) {
  constexpr int R = K::rad[s], B = K::E / R, Ns = K::ns(s);
  static_for<0, B>([&](auto bc) {
    constexpr int b = bc;
    const int j = t + b * K::TPT;
    const int base = (j & ~(Ns - 1)) * R + (j & (Ns - 1));
    static_for<0, R>([&](auto qc) {
      constexpr int q = qc;
      sm[smem_addr<K, COL>(l, base + q * Ns)] = v[b + q * B];
    });
  });
}

template <class K, bool COL, typename C>
__device__ __forceinline__ void gather(C (&v)[K::E], const C* sm, int l, int t) {
  static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = sm[smem_addr<K, COL>(l, t + e * K::TPT)]; });
}

template <class K, int s, bool LLF, bool SLF, typename C>
__device__ __forceinline__ void stages(C (&v)[K::E], C* sm, int& l, int& t, const C* __restrict__ tws) {
  // the shared-memory layout is columnar only when both ends are line-fastest
  constexpr bool COL = LLF && SLF;
  run_stage<K, s>(v, t, tws);
  if constexpr (s + 1 < K::S) {
    if constexpr (s > 0) __syncthreads();  // readers of the previous exchange are done
    scatter<K, s, COL>(v, sm, l, t);
    __syncthreads();
    if constexpr (s + 2 == K::S && LLF != SLF) {  // switch to the store mapping for the last stage
      const int tid = threadIdx.x;
      if constexpr (SLF) { l = tid % K::TL; t = tid / K::TL; } else { t = tid % K::TPT; l = tid / K::TPT; }
    }
    gather<K, COL>(v, sm, l, t);
    stages<K, s + 1, LLF, SLF>(v, sm, l, t, tws);
  }
}

// LLF: load mapping is line-fastest (adjacent threads = adjacent lines; use when ils == 1)
// SLF: same for the store side (ols == 1).   TW4: multiply the result by the four-step twiddle.
template <class K, bool LLF, bool SLF, bool TW4>
__global__ void __launch_bounds__(K::THREADS, K::MINB)
fft_lines_kernel(const Geom g, const cpx_t<typename K::real>* __restrict__ in, cpx_t<typename K::real>* __restrict__ out,
                 const cpx_t<typename K::real>* __restrict__ tws, const cpx_t<typename K::real>* __restrict__ tw_lo,
                 const cpx_t<typename K::real>* __restrict__ tw_hi, typename K::real scale) {
  using T = typename K::real;
  using C = cpx_t<T>;
  static_assert(LLF == SLF || K::S >= 2, "the transposing variant needs an exchange to re-map threads");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);

  const int tid = threadIdx.x;
  const unsigned tile = blockIdx.x;
  const int lt = tile % (unsigned)g.ntl;
  const unsigned rest = tile / (unsigned)g.ntl;
  const int o = rest % (unsigned)g.no;
  const int b = rest / (unsigned)g.no;

  int l, t;
  if constexpr (LLF) { l = tid % K::TL; t = tid / K::TL; } else { t = tid % K::TPT; l = tid / K::TPT; }

  C v[K::E];
  {
    const int line = lt * K::TL + l;
    const bool valid = line < g.nl;
    const C* ip = in + (long long)b * g.ibs + (long long)o * g.ios + (long long)line * g.ils + (long long)t * g.ins;
    const long long step = (long long)K::TPT * g.ins;
    static_for<0, K::E>([&](auto ec) {
      constexpr int e = ec;
      v[e] = valid ? (g.stream_hint ? ld_stream(ip + e * step) : ip[e * step]) : C{0, 0};
    });
    if (g.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = cswap(v[e]); });
  }

  stages<K, 0, LLF, SLF>(v, sm, l, t, tws);

  {
    const int line = lt * K::TL + l;
    const bool valid = line < g.nl;
    if constexpr (TW4) {
      const unsigned m = (unsigned)line / (unsigned)g.tw_div;
      const unsigned lomask = (1u << g.tw_lo_bits) - 1u;
      static_for<0, K::E>([&](auto ec) {
        constexpr int e = ec;
        const unsigned x = (unsigned)(t + e * K::TPT) * m;
        const C w = cmul(__ldg(tw_lo + (x & lomask)), __ldg(tw_hi + (x >> g.tw_lo_bits)));
        v[e] = cmul(v[e], w);
      });
    }
    if (scale != (T)1) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= scale; });
    if (g.swap_out) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = cswap(v[e]); });
    C* op = out + (long long)b * g.obs + (long long)o * g.oos + (long long)line * g.ols + (long long)t * g.ons;
    const long long step = (long long)K::TPT * g.ons;
    if (valid) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; if (g.stream_hint) st_stream(op + e * step, v[e]); else op[e * step] = v[e]; });
  }
}

}  // namespace b200fft
