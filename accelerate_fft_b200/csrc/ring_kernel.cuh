// Persistent, TMA-fed variant of the Stockham line kernel for CONTIGUOUS rows whose tile is too big to
// hide HBM latency by occupancy alone (tiles of 32-68 KB: c128 N=2048/4096, c64 N=4096/8192).
//
// One CTA per SM (or two) stays resident and walks over its tiles.  A ring of NS shared-memory stage
// buffers is filled by bulk asynchronous copies (cp.async.bulk = TMA 1D, completion on an mbarrier);
// G independent thread groups each take every G-th tile, pull their points from the stage buffer into
// registers, run the register-radix stages with the exchanges done IN the same stage buffer (named
// barriers per group, not __syncthreads), and as soon as a group has gathered for its last stage the
// buffer is handed back to the TMA engine for tile k+NS while the group finishes the last butterflies
// and streams the result to HBM straight from registers.  HBM reads are therefore always in flight
// behind the arithmetic: one HBM read + one HBM write per element, like the plain kernel, but without
// the exposed load latency (ncu on the plain c128 N=4096 kernel: long_scoreboard was the top stall).
//
// Replaces, for its share of a plan, cufftExecC2C / cufftExecZ2Z behind PTX.hs:112-124.
#pragma once
#include <cstdint>

#include "fft_kernel.cuh"

namespace b200fft {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// `mbarrier.try_wait` (unlike `test_wait`) is itself a blocking instruction: the hardware suspends the thread until the phase
// completes or a system-defined time limit passes, so this loop is not a busy spin on the issue slots.  An explicit
// suspend-time hint (try_wait with a 10 ms hint) was measured on the band kernel, where a quarter of all issued warp
// instructions were these re-polls: 473 us against 472 us (profiles/r02_band_experiments.txt) -- no difference, not taken.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// TMA 1D: global -> shared, completion (bytes) signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <class K, int G_, int NS_>
struct RingCfg {
  using Base = K;
  static constexpr int G = G_, NS = NS_;
  static constexpr int THREADS = G * K::THREADS;
  static constexpr int MINB = (THREADS >= 512) ? 1 : 2;   // 128 registers per thread either way
  // a stage holds either the linear tile (as TMA delivers it) or the padded exchange layout
  static constexpr int STAGE_ELEMS = (K::ROW_ELEMS > K::N * K::TL) ? K::ROW_ELEMS : K::N * K::TL;
  static constexpr size_t STAGE_BYTES = (((size_t)STAGE_ELEMS * K::ESZ + 127) / 128) * 128;
  static constexpr size_t SMEM = STAGE_BYTES * NS + 8 * NS + 16;
};

// Contiguous rows only: line `line` occupies [line*N, (line+1)*N) of both `in` and `out`.
// g.nl = number of lines, g.ntl = number of tiles (TL lines each); in/out 16-byte aligned.
template <class R>
__global__ void __launch_bounds__(R::THREADS, R::MINB)
fft_ring_rows_kernel(const Geom g, const cpx_t<typename R::Base::real>* __restrict__ in, cpx_t<typename R::Base::real>* __restrict__ out,
                     const cpx_t<typename R::Base::real>* __restrict__ tws, const cpx_t<typename R::Base::real>* __restrict__,
                     const cpx_t<typename R::Base::real>* __restrict__, typename R::Base::real scale) {
  using K = typename R::Base;
  using T = typename K::real;
  using C = cpx_t<T>;
  static_assert(K::S >= 2, "ring kernel needs at least one exchange");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + R::STAGE_BYTES * R::NS);

  const int grp = threadIdx.x / K::THREADS;
  const int tid = threadIdx.x % K::THREADS;
  const int t = tid % K::TPT, l = tid / K::TPT;
  const long long ntiles = g.ntl;
  // this CTA's tiles: blockIdx.x + k * gridDim.x, k = 0 .. nk-1
  const int nk = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  auto issue = [&](int k) {  // one thread: start the TMA copy of this CTA's k-th tile into stage k % NS
    const long long tile = blockIdx.x + (long long)k * gridDim.x;
    const long long line0 = tile * K::TL;
    const int lines = (g.nl - line0 < K::TL) ? (int)(g.nl - line0) : K::TL;
    const uint32_t bytes = (uint32_t)lines * K::N * K::ESZ;
    const int s = k % R::NS;
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(smem_raw + R::STAGE_BYTES * s, in + line0 * K::N, bytes, &full[s]);
  };

  // programmatic dependent launch (plan.cu): the next kernel on the stream may be scheduled while this one drains; this one
  // sets up its barriers while its predecessor drains and touches global memory only after the predecessor has completed
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0) {
    for (int s = 0; s < R::NS; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0)
    for (int k = 0; k < R::NS && k < nk; k++) issue(k);

  const int bar_id = 1 + grp;
  for (int k = grp; k < nk; k += R::G) {
    const int s = k % R::NS;
    C* sm = reinterpret_cast<C*>(smem_raw + R::STAGE_BYTES * s);
    const long long tile = blockIdx.x + (long long)k * gridDim.x;
    const long long line = tile * K::TL + l;
    const bool valid = line < g.nl;

    mbar_wait(&full[s], (uint32_t)((k / R::NS) & 1));
    C v[K::E];
    static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = sm[l * K::N + t + e * K::TPT]; });
    if (g.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
    group_bar(bar_id, K::THREADS);  // the linear tile is consumed; the buffer becomes the exchange space

    // all stages but the last, exchanging through the stage buffer
    run_stage<K, 0>(v, t, tws);
    scatter<K, 0, false>(v, sm, l, t);
    group_bar(bar_id, K::THREADS);
    gather<K, false>(v, sm, l, t);
    if constexpr (K::S >= 3) {
      run_stage<K, 1>(v, t, tws);
      group_bar(bar_id, K::THREADS);
      scatter<K, 1, false>(v, sm, l, t);
      group_bar(bar_id, K::THREADS);
      gather<K, false>(v, sm, l, t);
    }
    if constexpr (K::S >= 4) {
      run_stage<K, 2>(v, t, tws);
      group_bar(bar_id, K::THREADS);
      scatter<K, 2, false>(v, sm, l, t);
      group_bar(bar_id, K::THREADS);
      gather<K, false>(v, sm, l, t);
    }
    // hand the buffer back to the TMA engine for tile k + NS, then finish from registers
    fence_proxy_async();
    group_bar(bar_id, K::THREADS);
    if (tid == 0 && k + R::NS < nk) issue(k + R::NS);

    run_stage<K, K::S - 1>(v, t, tws);
    if (scale != (T)1) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= scale; });
    if (g.swap_out) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
    if (valid) {
      C* op = out + line * K::N + t;
      static_for<0, K::E>([&](auto ec) { constexpr int e = ec; st_stream(op + e * K::TPT, v[e]); });
    }
  }
}

}  // namespace b200fft
