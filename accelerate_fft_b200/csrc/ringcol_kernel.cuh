// Strided (column) lines whose tile fills the shared memory of an SM -- c64 N=1024 x 16 columns = 128 KB: cfg5's z axis, the y / z
// scatter passes of the slab transform -- as a persistent TMA-fed kernel with ONE buffer.
//
// The lock-step kernel (fft_kernel.cuh) fits one such CTA per SM: load, butterflies and stores take turns and HBM idles in
// between (cfg5's z pass: 4.2 TB/s where 64 KB column tiles with two CTAs per SM reach 5.6).  Here the CTA stays resident: its
// next tile is fetched by the TMA engine (cp.async.bulk.tensor, a rank-3 map of the array: columns x rows x outer, the tile as
// boxes of 256 rows x TL columns) into the SAME buffer as soon as the last gather of the current tile has emptied it, i.e. while
// the last register stage and the stores of the current tile run.  The dense landing layout [row][TL] is the column exchange
// layout of the line kernel (no padding when TL fills a wavefront), so the landed tile is pulled straight into registers.
// (A second landing area for part of the next tile is a compile-time option, measured and left off: see B200FFT_RINGCOL_EARLY.)
// One HBM read + one HBM write per element; the stores can be scattered over peer buffers like the lock-step kernel's (Geom::peer).
//
// Replaces, for its share of a plan, cufftExecC2C / cufftExecZ2Z behind
// /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs:112-124 (plan3D :155).
#pragma once
#include <cuda.h>

#include "ring_kernel.cuh"

namespace b200fft {

__device__ __forceinline__ void mbar_arrive_plain(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

// EARLY boxes of the next tile can land ahead of time in a second area (requested right after the pull).  Measured on B200, cfg5
// (1024^3 c64, whole transform): EARLY = 0: 9461 us, 1: 9446, 2: 9606, 3: 9830 (lock-step kernel: 9896) -- the load is not what
// the pass waits for, and the extra shared memory shrinks the L1 that serves the twiddle tables.  Default 0: one buffer.
#ifndef B200FFT_RINGCOL_EARLY
#define B200FFT_RINGCOL_EARLY 0
#endif
template <class K_, bool TW4_ = false, int EARLY_ = B200FFT_RINGCOL_EARLY>
struct RingColCfg {
  using K = K_;
  static constexpr bool TW4 = TW4_;     // multiply by the four-step twiddle w_L^(k * line) before the store (first pass of a big 1D)
  static constexpr int THREADS = K::THREADS;
  static constexpr int BOX_ROWS = K::N < 256 ? K::N : 256;          // a TMA box holds at most 256 rows
  static constexpr int NBOX = K::N / BOX_ROWS;
  static constexpr size_t TILE_BYTES = (size_t)K::N * K::TL * K::ESZ;
  static constexpr size_t BUF_BYTES = (((size_t)K::COL_ELEMS * K::ESZ + 127) / 128) * 128;
  static constexpr size_t BOX_BYTES = (size_t)BOX_ROWS * K::TL * K::ESZ;
  static constexpr int EARLY = EARLY_ < NBOX ? EARLY_ : NBOX - 1;    // boxes that land in the second area, ahead of time
  static constexpr size_t EARLY_BYTES = BOX_BYTES * EARLY;
  static constexpr int EARLY_ROWS = BOX_ROWS * EARLY;
  static constexpr size_t SMEM = BUF_BYTES + EARLY_BYTES + 32;
  static_assert(NBOX >= 2 && EARLY >= 0 && SMEM <= 232448, "tile + early landing area must fit one SM");
  static_assert(EARLY_ROWS % K::TPT == 0, "a thread's points split between the two areas at a compile-time index");
  static_assert(K::XPAD == 0, "the landed tile must already be in the exchange layout");
  static_assert(K::S >= 2, "at least one exchange");
};

// g: the pass's geometry (nb == 1, ils == ols == 1, nl a multiple of TL); tm: the INPUT as a tensor {2 I, N, O}
template <class R>
__global__ void __launch_bounds__(R::THREADS, 1)
fft_ringcol_kernel(const __grid_constant__ CUtensorMap tm, const Geom g, cpx_t<typename R::K::real>* __restrict__ out,
                   const cpx_t<typename R::K::real>* __restrict__ tws, const cpx_t<typename R::K::real>* __restrict__ tw_lo,
                   const cpx_t<typename R::K::real>* __restrict__ tw_hi, typename R::K::real scale) {
  using K = typename R::K;
  using T = typename K::real;
  using C = cpx_t<T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);                                   // exchange space; its tail also lands the last box
  const C* early = reinterpret_cast<const C*>(smem_raw + R::BUF_BYTES);      // rows [0, EARLY_ROWS) of the next tile
  uint64_t* full_early = reinterpret_cast<uint64_t*>(smem_raw + R::BUF_BYTES + R::EARLY_BYTES);
  uint64_t* full_last = full_early + 1;
  const int tid = threadIdx.x;
  const int l = tid % K::TL, t = tid / K::TL;
  const long long ntiles = (long long)g.no * g.ntl;
  const int nk = ntiles > blockIdx.x ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  // one thread: this CTA's k-th tile = NBOX boxes of BOX_ROWS rows x TL columns; all but the last go to the early area
  auto issue_early = [&](int k) {
    const long long tile = blockIdx.x + (long long)k * gridDim.x;
    const int lt = (int)(tile % g.ntl), o = (int)(tile / g.ntl);
    if (R::EARLY == 0) { mbar_arrive_plain(full_early); return; }
    mbar_expect_tx(full_early, (uint32_t)R::EARLY_BYTES);
    for (int b = 0; b < R::EARLY; b++)
      tma_load_3d(smem_raw + R::BUF_BYTES + (size_t)b * R::BOX_BYTES, &tm, 2 * lt * K::TL, b * R::BOX_ROWS, o, full_early);
  };
  auto issue_last = [&](int k) {   // the last box lands at its own rows inside the exchange space
    const long long tile = blockIdx.x + (long long)k * gridDim.x;
    const int lt = (int)(tile % g.ntl), o = (int)(tile / g.ntl);
    mbar_expect_tx(full_last, (uint32_t)(R::BOX_BYTES * (R::NBOX - R::EARLY)));
    for (int b = R::EARLY; b < R::NBOX; b++)
      tma_load_3d(smem_raw + (size_t)b * R::BOX_BYTES, &tm, 2 * lt * K::TL, b * R::BOX_ROWS, o, full_last);
  };

  // programmatic dependent launch (plan.cu): set up while the previous kernel drains, touch its results only afterwards
  asm volatile("griddepcontrol.launch_dependents;");
  if (tid == 0) { mbar_init(full_early, 1); mbar_init(full_last, 1); fence_mbar_init(); }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid == 0 && nk > 0) { issue_early(0); issue_last(0); }

  for (int k = 0; k < nk; k++) {
    const long long tile = blockIdx.x + (long long)k * gridDim.x;
    const int lt = (int)(tile % g.ntl), o = (int)(tile / g.ntl);
    mbar_wait(full_early, (uint32_t)(k & 1));
    mbar_wait(full_last, (uint32_t)(k & 1));
    C v[K::E];
    static_for<0, K::E>([&](auto ec) {
      constexpr int e = ec;
      if constexpr (e * K::TPT < R::EARLY_ROWS) v[e] = early[(t + e * K::TPT) * K::TL + l];
      else v[e] = sm[(t + e * K::TPT) * K::TL + l];
    });
    if (g.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
    fence_proxy_async();
    __syncthreads();   // the landed tile is in registers: the early area is free for the next tile, the rest is the exchange space
    if (tid == 0 && k + 1 < nk) issue_early(k + 1);
    run_stage<K, 0>(v, t, tws);
    scatter<K, 0, true>(v, sm, l, t);
    static_for<1, K::S - 1>([&](auto sc) {
      constexpr int s = sc;
      __syncthreads();
      gather<K, true>(v, sm, l, t);
      run_stage<K, s>(v, t, tws);
      __syncthreads();
      scatter<K, s, true>(v, sm, l, t);
    });
    __syncthreads();
    gather<K, true>(v, sm, l, t);
    // the buffer is empty: hand it to the TMA engine for the next tile, then finish this one from registers
    fence_proxy_async();
    __syncthreads();
    if (tid == 0 && k + 1 < nk) issue_last(k + 1);

    run_stage<K, K::S - 1>(v, t, tws);
    if constexpr (R::TW4) {
      // four-step twiddle w_L^(k * m), k = t + e*TPT, m = the line's multiplier: an anchor per 8 points from the two-level
      // table and a running product in between (as fft_kernel.cuh)
      const unsigned m = g.tw_from_o ? (unsigned)o : (unsigned)(lt * K::TL + l) / (unsigned)g.tw_div;
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
    const T sy = g.swap_out ? -scale : scale;
    if (scale != (T)1 || g.swap_out) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= sy; });
    const int line = lt * K::TL + l;
    const long long off = (long long)o * g.oos + line;
    if (g.npeers) {   // scatter store over peer buffers (see Geom)
      static_for<0, K::E>([&](auto ec) { constexpr int e = ec; *peer_addr<C>(g, off, t + e * K::TPT) = v[e]; });
    } else {
      char* pb = reinterpret_cast<char*>(out + off + (long long)t * g.ons);
      const unsigned step_b = (unsigned)((long long)K::TPT * g.ons * (long long)sizeof(C));
      static_for<0, K::E>([&](auto ec) {
        constexpr int e = ec;
        *reinterpret_cast<C*>(pb + (unsigned long long)(unsigned)e * step_b) = v[e];
      });
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// The transposing last pass of a big four-step (rows in, line-fastest out) the same way: TL contiguous lines of N points land by
// bulk copies (one per line), are pulled into registers, exchanged in the padded row layout inside the same buffer, and after the
// last gather -- now with adjacent threads on adjacent LINES -- the buffer goes back to the TMA engine for the next tile while the
// last stage runs and the results are stored as runs of TL consecutive lines (32 lines = 256 B runs where the lock-step kernel's
// 64 KB tile has 16 = 128 B; page-scattered 128 B stores reach 4.2 TB/s, 256 B ones 5.6: profiles/r02_scatter_bw.txt).
template <class K_>
struct RingTransCfg {
  using K = K_;
  static constexpr int THREADS = K::THREADS;
  static constexpr size_t LINE_BYTES = (size_t)K::N * K::ESZ;
  static constexpr size_t TILE_BYTES = LINE_BYTES * K::TL;
  static constexpr size_t BUF_ELEMS = (size_t)K::ROW_ELEMS > (size_t)K::N * K::TL ? (size_t)K::ROW_ELEMS : (size_t)K::N * K::TL;
  static constexpr size_t BUF_BYTES = ((BUF_ELEMS * K::ESZ + 127) / 128) * 128;
  static constexpr size_t SMEM = BUF_BYTES + 16;
  static_assert(K::S >= 2, "the transposing variant needs an exchange to re-map threads");
  static_assert(SMEM <= 232448, "one SM");
};

// g: the pass's geometry (ins == 1, ols == 1, nl a multiple of TL); tile = (b, o, lt)
template <class R>
__global__ void __launch_bounds__(R::THREADS, 1)
fft_ringtrans_kernel(const Geom g, const cpx_t<typename R::K::real>* __restrict__ in, cpx_t<typename R::K::real>* __restrict__ out,
                     const cpx_t<typename R::K::real>* __restrict__ tws, typename R::K::real scale) {
  using K = typename R::K;
  using T = typename K::real;
  using C = cpx_t<T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + R::BUF_BYTES);
  const int tid = threadIdx.x;
  const long long ntiles = (long long)g.nb * g.no * g.ntl;
  const int nk = ntiles > blockIdx.x ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  auto decode = [&](int k, int& lt, int& o, int& b) {
    const long long tile = blockIdx.x + (long long)k * gridDim.x;
    lt = (int)(tile % g.ntl);
    const long long rest = tile / g.ntl;
    o = (int)(rest % g.no);
    b = (int)(rest / g.no);
  };
  auto issue = [&](int k) {   // one thread: TL bulk copies, one per line
    int lt, o, b;
    decode(k, lt, o, b);
    const C* base = in + (long long)b * g.ibs + (long long)o * g.ios + (long long)lt * K::TL * g.ils;
    mbar_expect_tx(full, (uint32_t)R::TILE_BYTES);
    for (int r = 0; r < K::TL; r++) bulk_g2s(smem_raw + (size_t)r * R::LINE_BYTES, base + (long long)r * g.ils, (uint32_t)R::LINE_BYTES, full);
  };

  asm volatile("griddepcontrol.launch_dependents;");
  if (tid == 0) { mbar_init(full, 1); fence_mbar_init(); }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid == 0 && nk > 0) issue(0);

  for (int k = 0; k < nk; k++) {
    int lt, o, b;
    decode(k, lt, o, b);
    int t = tid % K::TPT, l = tid / K::TPT;          // load mapping: adjacent threads = adjacent points of a line
    mbar_wait(full, (uint32_t)(k & 1));
    C v[K::E];
    static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = sm[l * K::N + t + e * K::TPT]; });
    if (g.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
    __syncthreads();   // the dense tile is in registers: the buffer becomes the (padded) exchange space
    run_stage<K, 0>(v, t, tws);
    scatter<K, 0, false>(v, sm, l, t);
    static_for<1, K::S - 1>([&](auto sc) {
      constexpr int s = sc;
      __syncthreads();
      gather<K, false>(v, sm, l, t);
      run_stage<K, s>(v, t, tws);
      __syncthreads();
      scatter<K, s, false>(v, sm, l, t);
    });
    __syncthreads();
    l = tid % K::TL; t = tid / K::TL;                // store mapping: adjacent threads = adjacent lines
    gather<K, false>(v, sm, l, t);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0 && k + 1 < nk) issue(k + 1);

    run_stage<K, K::S - 1>(v, t, tws);
    const T sy = g.swap_out ? -scale : scale;
    if (scale != (T)1 || g.swap_out) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= sy; });
    char* pb = reinterpret_cast<char*>(out + (long long)b * g.obs + (long long)o * g.oos + (long long)(lt * K::TL + l) + (long long)t * g.ons);
    const unsigned step_b = (unsigned)((long long)K::TPT * g.ons * (long long)sizeof(C));
    static_for<0, K::E>([&](auto ec) {
      constexpr int e = ec;
      *reinterpret_cast<C*>(pb + (unsigned long long)(unsigned)e * step_b) = v[e];
    });
  }
}

}  // namespace b200fft
