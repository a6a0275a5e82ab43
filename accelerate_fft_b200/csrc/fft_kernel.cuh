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
#ifndef ROWROT_ENABLED
#define ROWROT_ENABLED true
#endif

// Where the lines of one launch live.  All strides in complex elements.
//   address(b, o, line, n) = b*bs + o*os + line*ls + n*ns
// b < nb (batch), o < no (outer), line < nl (tiled by TL), n < N (transform index).
struct Geom {
  long long ibs, ios, ils, ins;  // input
  long long obs, oos, ols, ons;  // output
  int nb, no, nl, ntl;           // ntl = ceil(nl / TL)
  // four-step twiddle  w_L^(k * (line / tw_div)),  L = 2^(tw_hi_bits_total): exponent split lo/hi
  int tw_div;
  int tw_from_o;                 // four-step multiplier is the outer index o instead of line / tw_div
  int tw_lo_bits;
  int swap_in, swap_out;         // conjugate on load / on store: inverse = conj(fwd(conj(x)))
  // PRE2 ("row pair") kernels only: the tile's input is x[n] + (-1)^line * x[n + pre2_off] -- the radix-2 first
  // stage of a strided axis of length 2*M folded into the row pass; the result is multiplied by w_{2M}^(line*o)
  long long pre2_off;
  // Scatter store: output index n of the transformed axis lands in buffer peer[n >> peer_shift] at index
  // n & (2^peer_shift - 1).  With the buffers of other GPUs (b200fftExecScatter) it is the all-to-all of the
  // slab-decomposed 3D transform folded into the pass's own stores (NVLink peer addresses are ordinary global
  // addresses); with peer[0] = out + N/2 and peer[1] = out it is the half rotation of DFT/Centre.hs's `shift`
  // (b200fftExecShifted) -- both free in the last butterfly pass.
  int npeers, peer_shift;
  void* peer[16];
};

// address of output index k of (b, o, line) under the scatter store
template <typename C>
__device__ __forceinline__ C* peer_addr(const Geom& g, long long off, int k) {
  return reinterpret_cast<C*>(g.peer[k >> g.peer_shift]) + off + (long long)(k & ((1 << g.peer_shift) - 1)) * g.ons;
}

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
  // single register stage, one whole line per thread: contiguous rows are staged through shared memory so that the
  // global loads / stores are coalesced (fft_small_rows_tile), pitch N + 1
  static constexpr bool SMALL_ROWS = (S == 1) && (TPT == 1) && (N >= 2);
  template <bool COL> static constexpr size_t smem_bytes() {
    return (S > 1) ? (size_t)(COL ? COL_ELEMS : ROW_ELEMS) * ESZ : (!COL && SMALL_ROWS) ? (size_t)TL * (N + 1) * ESZ : 0;
  }
};

// Shared-memory address of point i of line l.  Every index used below splits as i = it + ic with `it` a
// function of the thread and `ic` a compile-time constant chosen so that the low LOGQ bits never carry
// (all of N, E, TL, TPT, the radices and Q = 2^LOGQ are powers of two), hence
//     addr(l, it + ic) = addr_rt(l, it) + addr_ct(ic)
// and each access is one runtime base (computed once per stage) plus an immediate offset -- the
// integer/address instructions were 38 % of the issue slots before this split (profiles/r01_sass_mix_*.txt).
template <class K, bool COL>
__device__ __forceinline__ int addr_rt(int l, int it) {
  if constexpr (COL) return it * K::TL + (it >> K::LOGQ) * K::XPAD + l;
  else return l * K::PITCH + it + (it >> K::LOGQ);
}
template <class K, bool COL>
__device__ __forceinline__ constexpr int addr_ct(int ic) {
  if constexpr (COL) return ic * K::TL + (ic >> K::LOGQ) * K::XPAD;
  else return ic + (ic >> K::LOGQ);
}

// One register stage: E/R butterflies of radix R per thread.
// LDG: the twiddle table is in global memory (read through the non-coherent path); false = shared memory
template <class K, int s, typename C, bool LDG = true>
__device__ __forceinline__ void run_stage(C (&v)[K::E], int t, const C* __restrict__ tws) {
  constexpr int R = K::rad[s], B = K::E / R, Ns = K::ns(s);
  // twiddle index k = (t + b*TPT) & (Ns-1): thread part + compile-time part
  const C* tp0 = tws + K::tw_off(s) + ((Ns <= K::TPT) ? (t & (Ns - 1)) : t);
  static_for<0, B>([&](auto bc) {
    constexpr int b = bc;
    C a[R];
    static_for<0, R>([&](auto rc) { constexpr int r = rc; a[r] = v[b + r * B]; });
    if constexpr (s > 0) {
      constexpr int kc = (Ns <= K::TPT) ? 0 : ((b * K::TPT) & (Ns - 1));
      static_for<1, R>([&](auto rc) {
        constexpr int r = rc;
        if constexpr (LDG) a[r] = cmul(a[r], __ldg(tp0 + kc + (r - 1) * Ns));
        else a[r] = cmul(a[r], tp0[kc + (r - 1) * Ns]);
      });
    }
    dft<R>(a);
    static_for<0, R>([&](auto rc) { constexpr int r = rc; v[b + r * B] = a[r]; });
  });
}

// Stockham scatter of stage s results into shared memory: butterfly j = t + b*TPT writes its q-th output to
// point (j & ~(Ns-1))*R + (j & (Ns-1)) + q*Ns.
template <class K, int s, bool COL, typename C>
__device__ __forceinline__ void scatter(const C (&v)[K::E], C* sm, int l, int t) {
  constexpr int R = K::rad[s], B = K::E / R, Ns = K::ns(s);
  constexpr bool SMALL = Ns <= K::TPT;
  const int it = SMALL ? (t & ~(Ns - 1)) * R + (t & (Ns - 1)) : t;
  C* base = sm + addr_rt<K, COL>(l, it);
  static_for<0, B>([&](auto bc) {
    constexpr int b = bc;
    constexpr int jb = b * K::TPT;
    constexpr int ic0 = SMALL ? jb * R : (jb & ~(Ns - 1)) * R + (jb & (Ns - 1));
    static_for<0, R>([&](auto qc) {
      constexpr int q = qc;
      base[addr_ct<K, COL>(ic0 + q * Ns)] = v[b + q * B];
    });
  });
}

template <class K, bool COL, typename C>
__device__ __forceinline__ void gather(C (&v)[K::E], const C* sm, int l, int t) {
  const C* base = sm + addr_rt<K, COL>(l, t);
  static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = base[addr_ct<K, COL>(e * K::TPT)]; });
}

// The inverse direction re-uses the forward butterflies: IDFT(x) = conj(DFT(conj(x))).  The conjugations are
// applied by the first pass of a plan on load and by the last pass on store; each is compiled as a CLONE of
// the adjoining code (first stage / last stage) under a warp-uniform branch, so the negation folds into the
// operand modifiers of the butterflies' FADD/FFMA and costs nothing -- a runtime swap or sign multiply cost
// 2-4 register moves per point (profiles/r01_sass_mix_*.txt).

// LLF: load mapping is line-fastest (adjacent threads = adjacent lines; use when ils == 1); otherwise the
//      points of a line are contiguous (ins == 1, compile-time offsets).
// SLF: same for the store side (ols == 1, otherwise ons == 1).
// TW4: multiply the result by the four-step twiddle.
// PRE2: row kernels with TL == 1 only -- see Geom::pre2_off.  Two consecutive tiles (line = 0, 1) read the same
//      two rows (the second reader hits L2) and produce the two outputs of the radix-2 butterfly between them.
// CG:  read the input with ld.global.cg (L2 only) -- for tiles another CTA of the SAME launch produced.
// HINT: bit 0 evict-first loads (ld.global.cs), bit 1 evict-first stores (compile-time: no dead twins).
//
// Four-step twiddle w_L^(k*m), k = t + e*TPT the output index held by this thread, m the line's multiplier
// (line / tw_div, or the outer index o when g.tw_from_o).  Looking every factor up costs two dependent
// table reads per point (measured: the +tw passes of cfg4 ran at 2.8-4.7 TB/s against 5.5 without);
// instead one anchor per 8 points comes from the two-level table and the 7 in between from a running
// product with the step w_L^(TPT*m): 2 + E/4 table reads per thread instead of 2E, error <= 8 ulp.
template <class K, bool LLF, bool SLF, bool TW4, bool CG = false, int HINT = 0, bool PRE2 = false>
__device__ __forceinline__ void fft_lines_tile(const Geom& g, unsigned tile, const cpx_t<typename K::real>* __restrict__ in,
                                               cpx_t<typename K::real>* __restrict__ out, const cpx_t<typename K::real>* __restrict__ tws,
                                               const cpx_t<typename K::real>* __restrict__ tw_lo,
                                               const cpx_t<typename K::real>* __restrict__ tw_hi, typename K::real scale,
                                               cpx_t<typename K::real>* sm) {
  using T = typename K::real;
  using C = cpx_t<T>;
  static_assert(LLF == SLF || K::S >= 2, "the transposing variant needs an exchange to re-map threads");
  static_assert(!PRE2 || (!LLF && !SLF && !TW4 && K::TL == 1), "PRE2 is a row-kernel option");
  constexpr bool COL = LLF && SLF;   // the shared-memory layout is columnar only when both ends are line-fastest
  const int tid = threadIdx.x;
  int lt, o, b;
  if (g.no == 1 && g.nb == 1) { lt = (int)tile; o = 0; b = 0; }   // the common shape: no integer divisions
  else {
    lt = tile % (unsigned)g.ntl;
    const unsigned rest = tile / (unsigned)g.ntl;
    o = rest % (unsigned)g.no;
    b = rest / (unsigned)g.no;
  }

  int l, t;
  if constexpr (LLF) { l = tid % K::TL; t = tid / K::TL; } else { t = tid % K::TPT; l = tid / K::TPT; }

  C v[K::E];
  {
    // lines past the end of a ragged last tile re-read the last valid line (no predicates, no zero fill);
    // their results are simply not stored
    const int line = min(lt * K::TL + l, g.nl - 1);
    const C* ip = in + (long long)b * g.ibs + (long long)o * g.ios + (long long)line * g.ils;
    if constexpr (LLF) ip += (long long)t * g.ins; else ip += t;
    // byte stride between a thread's points, 32 bits (the planner rejects strides beyond 4 GiB): the address of
    // point e is ONE IMAD.WIDE.U32 (e * step + base) instead of a 64 x 64-bit multiply-add
    const unsigned step_b = (unsigned)((long long)K::TPT * g.ins * (long long)sizeof(C));
    auto head = [&](auto cj) {   // load + first register stage (+ its scatter), cloned on the conjugation
      constexpr bool CJ = decltype(cj)::value;
      const char* p = reinterpret_cast<const char*>(ip);
      static_for<0, K::E>([&](auto ec) {
        constexpr int e = ec;
        const C* q = LLF ? reinterpret_cast<const C*>(p + (unsigned long long)(unsigned)e * step_b) : ip + e * K::TPT;
        C x;
        if constexpr (CG) x = __ldcg(q);
        else if constexpr (HINT & 1) x = ld_stream(q);
        else x = *q;
        if constexpr (PRE2) {
          const C y = (ip + g.pre2_off)[e * K::TPT];
          const T sg = (lt & 1) ? (T)-1 : (T)1;
          x.x = fma(sg, y.x, x.x);
          x.y = fma(sg, y.y, x.y);
        }
        if constexpr (CJ) x.y = -x.y;
        v[e] = x;
      });
      run_stage<K, 0>(v, t, tws);
      if constexpr (K::S > 1) scatter<K, 0, COL>(v, sm, l, t);
    };
    if (g.swap_in) head(std::true_type{}); else head(std::false_type{});
  }

  // middle stages: gather, butterflies, scatter
  static_for<1, (K::S > 1 ? K::S - 1 : 1)>([&](auto sc) {
    constexpr int s = sc;
    __syncthreads();
    gather<K, COL>(v, sm, l, t);
    run_stage<K, s>(v, t, tws);
    __syncthreads();   // every gather of the previous exchange is done
    scatter<K, s, COL>(v, sm, l, t);
  });
  if constexpr (K::S > 1) {
    __syncthreads();
    if constexpr (LLF != SLF) {  // switch to the store mapping for the last stage
      if constexpr (SLF) { l = tid % K::TL; t = tid / K::TL; } else { t = tid % K::TPT; l = tid / K::TPT; }
    }
    gather<K, COL>(v, sm, l, t);
  }

  {
    const int line = lt * K::TL + l;
    const bool valid = line < g.nl;
    C* op = out + (long long)b * g.obs + (long long)o * g.oos + (long long)line * g.ols;
    if constexpr (SLF) op += (long long)t * g.ons; else op += t;
    const unsigned step_b = (unsigned)((long long)K::TPT * g.ons * (long long)sizeof(C));
    auto tail = [&](auto cj) {   // last register stage + twiddle + scale + store, cloned on the conjugation
      constexpr bool CJ = decltype(cj)::value;
      if constexpr (K::S > 1) run_stage<K, K::S - 1>(v, t, tws);
      if constexpr (TW4) {
        const unsigned m = g.tw_from_o ? (unsigned)o : (unsigned)line / (unsigned)g.tw_div;
        const unsigned lomask = (1u << g.tw_lo_bits) - 1u;
        auto root = [&](unsigned x) { return cmul(__ldg(tw_lo + (x & lomask)), __ldg(tw_hi + (x >> g.tw_lo_bits))); };
        constexpr int CH = (K::E < 8) ? K::E : 8;
        const C stepw = root((unsigned)K::TPT * m);
        static_for<0, K::E / CH>([&](auto qc) {
          constexpr int q = qc;
          C w = root((unsigned)(t + q * CH * K::TPT) * m);
          static_for<0, CH>([&](auto rc) {
            constexpr int e = q * CH + rc;
            v[e] = cmul(v[e], w);
            if constexpr (rc + 1 < CH) w = cmul(w, stepw);
          });
        });
      }
      if constexpr (PRE2) {
        if (lt & 1) {   // CTA-uniform: the odd output of the pair carries the twiddle w_{2M}^o
          const unsigned lomask = (1u << g.tw_lo_bits) - 1u;
          const C w = cmul(__ldg(tw_lo + ((unsigned)o & lomask)), __ldg(tw_hi + ((unsigned)o >> g.tw_lo_bits)));
          static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = cmul(v[e], w); });
        }
      }
      if (scale != (T)1) {
        const T sy = CJ ? -scale : scale;
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= sy; });
      } else if constexpr (CJ) {
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
      }
      if constexpr (SLF && !TW4 && !CG) {
        if (g.npeers) {   // CTA-uniform: scatter store (see Geom)
          if (valid) {
            const long long off = (long long)b * g.obs + (long long)o * g.oos + (long long)line * g.ols;
            static_for<0, K::E>([&](auto ec) { constexpr int e = ec; *peer_addr<C>(g, off, t + e * K::TPT) = v[e]; });
          }
          return;
        }
      }
      if constexpr (!SLF && !TW4 && !CG && !PRE2 && K::E >= 2 && ROWROT_ENABLED) {
        // contiguous lines: the half rotation of b200fftExecShifted is a rotation of the register index -- output
        // t + e*TPT lands at t + ((e + E/2) mod E)*TPT, compile-time offsets, no address arithmetic
        if (g.npeers) {
          if (valid) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; op[((e + K::E / 2) % K::E) * K::TPT] = v[e]; });
          return;
        }
      }
      if (valid) {
        char* p = reinterpret_cast<char*>(op);
        static_for<0, K::E>([&](auto ec) {
          constexpr int e = ec;
          C* q = SLF ? reinterpret_cast<C*>(p + (unsigned long long)(unsigned)e * step_b) : op + e * K::TPT;
          if constexpr (HINT & 2) st_stream(q, v[e]); else *q = v[e];
        });
      }
    };
    if (g.swap_out) tail(std::true_type{}); else tail(std::false_type{});
  }
}

// Rows of 2..16 points, one line per thread: a thread's points are N*sizeof(C) bytes apart from its neighbour's, so direct
// loads touch 32 different 128 B lines per instruction (c64 n=16: 30 % of HBM peak).  Contiguous rows go through shared
// memory instead: the CTA's TL*N-element chunk is loaded and stored with unit stride across threads.
template <class K>
__device__ __forceinline__ void fft_small_rows_tile(const Geom& g, unsigned tile, const cpx_t<typename K::real>* __restrict__ in,
                                                    cpx_t<typename K::real>* __restrict__ out, typename K::real scale,
                                                    cpx_t<typename K::real>* sm) {
  using T = typename K::real;
  using C = cpx_t<T>;
  constexpr int N = K::N, TL = K::TL, LG = K::ilog2(N);
  const int tid = threadIdx.x;
  int lt, o, b;
  if (g.no == 1 && g.nb == 1) { lt = (int)tile; o = 0; b = 0; }
  else {
    lt = tile % (unsigned)g.ntl;
    const unsigned rest = tile / (unsigned)g.ntl;
    o = rest % (unsigned)g.no;
    b = rest / (unsigned)g.no;
  }
  const long long line0 = (long long)lt * TL;
  const int nvalid = (int)(((long long)g.nl - line0 < TL ? (long long)g.nl - line0 : (long long)TL) * N);
  const C* ip = in + (long long)b * g.ibs + (long long)o * g.ios + line0 * N;
  C* op = out + (long long)b * g.obs + (long long)o * g.oos + line0 * N;
  static_for<0, N>([&](auto ic) {
    constexpr int i = ic;
    const int idx = tid + i * TL;
    if (idx < nvalid) sm[idx + (idx >> LG)] = ld_stream(ip + idx);
  });
  __syncthreads();
  C v[N];
  static_for<0, N>([&](auto ec) { constexpr int e = ec; v[e] = sm[tid * (N + 1) + e]; });
  if (g.swap_in) static_for<0, N>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
  dft<N>(v);
  const T sy = g.swap_out ? -scale : scale;
  if (scale != (T)1 || g.swap_out) static_for<0, N>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= sy; });
  static_for<0, N>([&](auto ec) { constexpr int e = ec; sm[tid * (N + 1) + e] = v[e]; });
  __syncthreads();
  static_for<0, N>([&](auto ic) {
    constexpr int i = ic;
    const int idx = tid + i * TL;
    if (idx < nvalid) st_stream(op + idx, sm[idx + (idx >> LG)]);
  });
}

// `in` and `out` carry __restrict__ although the planner runs some passes in place (src == dst: column passes of 2D / 3D
// plans, four-step first passes).  That is sound here for a reason the type system cannot see: a tile reads and writes exactly
// the same set of elements, all of its loads complete (they feed the first register stage, and at least one barrier follows)
// before its first store is issued, and no tile touches another tile's elements -- so no load can observe a store of the same
// launch, through the read-only path or otherwise.  tests/test_parity_gpu.py runs every in-place pass against the oracle.
template <class K, bool LLF, bool SLF, bool TW4, bool PRE2 = false>
__global__ void __launch_bounds__(K::THREADS, K::MINB)
fft_lines_kernel(const Geom g, const cpx_t<typename K::real>* __restrict__ in, cpx_t<typename K::real>* __restrict__ out,
                 const cpx_t<typename K::real>* __restrict__ tws, const cpx_t<typename K::real>* __restrict__ tw_lo,
                 const cpx_t<typename K::real>* __restrict__ tw_hi, typename K::real scale) {
  using C = cpx_t<typename K::real>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // programmatic dependent launch (plan.cu): let the next kernel on the stream be scheduled as this one drains, and do not
  // touch global memory before the previous kernel has completed and flushed (both are no-ops under an ordinary launch)
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if constexpr (!LLF && !SLF && !TW4 && !PRE2 && K::SMALL_ROWS) {
    if (g.ils == K::N && g.ols == K::N && g.ins == 1 && g.ons == 1 && g.npeers == 0) {   // CTA-uniform
      fft_small_rows_tile<K>(g, blockIdx.x, in, out, scale, reinterpret_cast<C*>(smem_raw));
      return;
    }
  }
#ifndef B200FFT_ROW_HINT
#define B200FFT_ROW_HINT 0
#endif
  // cache hints for long contiguous rows (bit 0 evict-first loads, bit 1 evict-first stores), compile-time experiment: measured on
  // B200, c64 8192-point rows: no hint 195.3 us, loads 216.6, stores 194.3, both 206.7 (cfg3 471 / 494 / 471 / 503 us) -- left off
  constexpr int HINT = (!LLF && !SLF && !TW4 && !PRE2 && K::N >= 2048) ? B200FFT_ROW_HINT : 0;
  fft_lines_tile<K, LLF, SLF, TW4, false, HINT, PRE2>(g, blockIdx.x, in, out, tws, tw_lo, tw_hi, scale, reinterpret_cast<C*>(smem_raw));
}

// The same tile function under a grid-stride loop: the launch decides how many CTAs (hence SMs) the pass occupies.  Used by
// the slab-decomposed 3D transform, whose NVLink-bound scatter pass needs only a fraction of the SMs and leaves the rest to
// the HBM-bound passes running beside it on other streams (slab.cu).
template <class K, bool LLF, bool SLF, bool TW4>
__global__ void __launch_bounds__(K::THREADS, K::MINB)
fft_lines_loop_kernel(const Geom g, const cpx_t<typename K::real>* __restrict__ in, cpx_t<typename K::real>* __restrict__ out,
                      const cpx_t<typename K::real>* __restrict__ tws, const cpx_t<typename K::real>* __restrict__ tw_lo,
                      const cpx_t<typename K::real>* __restrict__ tw_hi, typename K::real scale, unsigned ntiles) {
  using C = cpx_t<typename K::real>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    fft_lines_tile<K, LLF, SLF, TW4, false, 0, false>(g, tile, in, out, tws, tw_lo, tw_hi, scale, reinterpret_cast<C*>(smem_raw));
    __syncthreads();   // the exchange space is re-used by the next tile
  }
}

}  // namespace b200fft
