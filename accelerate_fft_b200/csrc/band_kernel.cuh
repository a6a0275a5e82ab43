// Two four-step phases in ONE launch with the intermediate kept in L2 -- persistent, warp-specialised, TMA-fed.
//
// Problem (cfg3's 8192-point column axis, cfg4's 2^28 points): a transform of N = N1*N2 points whose lines are strided
// needs two Stockham passes, and two launches send the intermediate out to HBM and back (3 HBM passes for a 2D FFT
// where 2 is the minimum).  The first fused kernel (fused_kernel.cuh) already proved that the 126 MB L2 can hold the
// intermediate (ncu: one HBM read + one HBM write for both phases) but ran its 16 KB tiles in lock step --
// load, wait, butterflies, store, fence, signal, barrier -- with eight 4-warp CTAs per SM: 37 % issue utilisation,
// `barrier` and `long_scoreboard` the top stalls, dependency spins in every CTA (profiles/r02_fused_l2_ncu.txt).
//
// Here ONE CTA per SM stays resident and its warps have roles:
//   * warp 0, one lane = LOADER.  Walks this CTA's tickets (same A-runs-`la`-bands-ahead order as the first kernel),
//     waits for a free landing stage, checks the ticket's dependency (phase B: all phase-A tiles of the band have
//     signalled; phase A: the band that used the scratch slot before has been drained) and starts the tile's copy:
//     phase A tiles come from the user's array through a TMA tensor map (cp.async.bulk.tensor, one instruction per
//     64 x 32-element tile, L2 evict-first), phase B tiles from the L2-resident scratch slot with one bulk copy
//     (cp.async.bulk).  Completion is an mbarrier transaction count; nobody else ever spins on global memory.
//   * warp 1 = SIGNALLER: publishing a stored phase-A tile needs a gpu-scope fence (~2000 cycles); one fence per round
//     covers whatever the groups have finished meanwhile, off every compute warp's path.
//   * G compute groups of 128 threads.  Whichever group is free claims the next landed tile (a ring of NST landing
//     stages), pulls it into registers and hands the stage straight back to the loader; the Stockham exchange runs in
//     the group's own buffer (two named barriers per tile), and the group finishes from registers: phase A multiplies
//     by the four-step twiddle and stores into the scratch slot, phase B stores to the output (evict-first).
// The compute warps execute no global loads, no address arithmetic for them, no fences and no dependency checks.
//
// One HBM read + one HBM write per element for BOTH phases.
//
// Replaces, for its share of a plan, cufftExecC2C / cufftExecZ2Z behind
// /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs:112-124 (plan2D :148, plan1D :141).
#pragma once
#include <cuda.h>
#include <cstdint>

#include "ring_kernel.cuh"

namespace b200fft {

// MODE_STRIDED: a strided axis [O][N1*N2][I], band = Wb adjacent columns of one o; phase A = N1-point columns over n1
//               (+ twiddle w_N^(k1 n2)), phase B = N2-point columns over n2, output row k1 + N1 k2 of the same columns.
//               OUTER: phase B multiplies by w_L^((k1 + N1 k2) * column) before the store (first pass of a big four-step).
// MODE_ROWS:    contiguous rows [R][N1*N2], band = TLB adjacent rows; phase A = N1-point columns over n1 (stride N2
//               inside the row, + twiddle w_M^(k1 n2)), phase B = N2-point rows with the transposed store
//               X[row + R*(k1 + N1*k2)] (the last pass of a big four-step).
enum BandMode { MODE_STRIDED = 0, MODE_ROWS = 1 };

struct BandParams {
  int nbands, nA, nB;            // bands; tiles per band in phase A / B
  int la, nslots;                // phase A runs `la` bands ahead of phase B; scratch slots
  int nbi;                       // band = bo * nbi + bi
  int a_ncg;                     // phase A: column groups (TLA wide) per band (STRIDED) / per row (ROWS)
  long long slot_elems;          // elements per scratch slot
  long long out_bo, out_bi;      // element offset of band (bo, bi) in the output
  long long out_ks;              // output stride of the transformed index k = k1 + N1*k2
  int tw_lo_bits;                // inner four-step twiddle table split (L = N1*N2)
  int tw_lo_n, tw_hi_n;          // its two parts' lengths (tw_lo_n + tw_hi_n <= TW4_MAX)
  int otw_lo_bits;               // OUTER: outer table split
  long long otw_col0;            // OUTER: column index of band (bo = 0, bi = 0), columns advance by Wb per bi
  int swap_in, swap_out;         // conjugate on load (phase A) / on store (phase B)
  int wb;                        // STRIDED: columns per band
  int debug;                     // developer timing experiments (results invalid): low 3 bits 1 ignore dependencies, 2 phase A only,
                                 // 3 phase B only; +8 no butterflies, +16 no global stores, +32 no loads (stages 'land' at once)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// TMA: one box of a rank-4 tensor, global -> shared, completion (bytes) on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_release_gpu(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_relaxed_gpu(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class KA_, class KB_, int MODE_, bool OUTER_, int G_, int NST_>
struct BandCfg {
  using KA = KA_;
  using KB = KB_;
  using real = typename KA_::real;
  static constexpr int MODE = MODE_;
  static constexpr bool OUTER = OUTER_;
  static constexpr int G = G_, NST = NST_;
  static constexpr int A_RUN = NST_ / 2;             // phase-A tiles started per loader round at most (phase B goes first)
  static constexpr int GT = KA::THREADS;             // threads of one compute group
  static constexpr int CTRL = 64;                    // warp 0 = loader, warp 1 = signaller
  static constexpr int THREADS = CTRL + G_ * GT;
  static constexpr int N1 = KA::N, N2 = KB::N, TLA = KA::TL, TLB = KB::TL;
  static constexpr int ESZ = KA::ESZ;
  static constexpr bool B_ROWS = (MODE_ == MODE_ROWS);
  static constexpr int A_BYTES = N1 * TLA * ESZ, B_BYTES = N2 * TLB * ESZ;
  static constexpr size_t STAGE_BYTES = (((size_t)(A_BYTES > B_BYTES ? A_BYTES : B_BYTES) + 127) / 128) * 128;   // a landed tile, dense
  static constexpr int B_LAY = B_ROWS ? KB::ROW_ELEMS : KB::COL_ELEMS;
  static constexpr size_t XCH_BYTES = (((size_t)(KA::COL_ELEMS > B_LAY ? KA::COL_ELEMS : B_LAY) * ESZ + 127) / 128) * 128;   // a group's exchange space
  static constexpr int TW4_MAX = 512;                // inner four-step table (lo + hi parts) kept in shared memory
  static constexpr int TW_ELEMS = KA::TW_LEN + KB::TW_LEN + TW4_MAX;
  static constexpr size_t XCH_OFF = STAGE_BYTES * NST;
  static constexpr size_t TW_OFF = XCH_OFF + XCH_BYTES * G;
  static constexpr size_t DESC_OFF = TW_OFF + (((size_t)TW_ELEMS * ESZ + 127) / 128) * 128;
  static constexpr size_t BAR_OFF = DESC_OFF + 16 * NST;
  static constexpr size_t SIG_OFF = BAR_OFF + 16 * NST;            // per group: [0] A tiles stored, [1] finished, [2] claimed seq, [8..15] band ring
  static constexpr size_t SMEM = SIG_OFF + 64 * G_ + 16 + 16;      // + the claim counter
  static_assert(KA::THREADS == KB::THREADS, "both phases run in the same group shape");
  static_assert(KA::S >= 2 && KB::S >= 2, "both phases exchange through shared memory");
  static_assert(G_ <= 14, "one named barrier per group");
  static_assert(SMEM <= 232448, "shared memory of one SM");
};

// counters: [0 .. nbands) phase-A tiles stored per band, [nbands .. 2 nbands) phase-B tiles pulled per band
template <class P>
__global__ void __launch_bounds__(P::THREADS, 1)
fft_band_kernel(const __grid_constant__ CUtensorMap tmA, const BandParams p, cpx_t<typename P::real>* __restrict__ out,
                cpx_t<typename P::real>* __restrict__ slots, const cpx_t<typename P::real>* __restrict__ twsA,
                const cpx_t<typename P::real>* __restrict__ twsB, const cpx_t<typename P::real>* __restrict__ tw_lo,
                const cpx_t<typename P::real>* __restrict__ tw_hi, const cpx_t<typename P::real>* __restrict__ otw_lo,
                const cpx_t<typename P::real>* __restrict__ otw_hi, typename P::real scale, unsigned* __restrict__ counters) {
  using KA = typename P::KA;
  using KB = typename P::KB;
  using T = typename P::real;
  using C = cpx_t<T>;
  constexpr int G = P::G, NST = P::NST, GT = P::GT;
  constexpr int N1 = P::N1, N2 = P::N2, TLA = P::TLA, TLB = P::TLB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* stw = reinterpret_cast<C*>(smem_raw + P::TW_OFF);           // stage twiddles of KA, then of KB
  int4* desc = reinterpret_cast<int4*>(smem_raw + P::DESC_OFF);  // per stage: {phase (-1 = stop), band, tile, 0}
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + P::BAR_OFF);
  uint64_t* empty = full + NST;
  volatile int* sig = reinterpret_cast<volatile int*>(smem_raw + P::SIG_OFF);

  for (int i = threadIdx.x; i < KA::TW_LEN; i += P::THREADS) stw[i] = twsA[i];
  for (int i = threadIdx.x; i < KB::TW_LEN; i += P::THREADS) stw[KA::TW_LEN + i] = twsB[i];
  // (the loader's acquire polls invalidate L1 every round: tables read with __ldg went back to L2 each time)
  for (int i = threadIdx.x; i < p.tw_lo_n; i += P::THREADS) stw[KA::TW_LEN + KB::TW_LEN + i] = tw_lo[i];
  for (int i = threadIdx.x; i < p.tw_hi_n; i += P::THREADS) stw[KA::TW_LEN + KB::TW_LEN + p.tw_lo_n + i] = tw_hi[i];
  if (threadIdx.x < G * 16 + 1) sig[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], GT); }
    fence_mbar_init();
  }
  __syncthreads();

  const unsigned nA = (unsigned)p.nA, nB = (unsigned)p.nB;
  auto stage_ptr = [&](int st) -> unsigned char* { return smem_raw + P::STAGE_BYTES * st; };

  if (threadIdx.x < 32) {
    // ================================ LOADER (one warp) ================================
    // This CTA owns every gridDim-th phase-A tile and every gridDim-th phase-B tile (two lists, each in band order).
    // Per round, lanes 16..31 look at the next 16 phase-A tiles and lanes 0..15 at the next 16 phase-B tiles: decode,
    // poll the tile's dependency counter (ONE L2 round trip for all 32 -- a single polling thread spent ~1 us per tile
    // and starved the compute groups, profiles/r02_band_kernel.txt), then the ready prefix of each list is started, B
    // first, each copy by its own lane into the next stage of the ring.  The lists advance independently: a phase-B
    // tile that is not ready yet does not hold back the HBM prefetch of phase A, and phase A runs ahead only as far as
    // the scratch slots allow.
    const int lane = threadIdx.x;
    const uint64_t pol = policy_evict_first();
    const unsigned totA = (unsigned)p.nbands * nA, totB = (unsigned)p.nbands * nB;
    unsigned nTA = (totA > blockIdx.x) ? (totA - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    unsigned nTB = (totB > blockIdx.x) ? (totB - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    if ((p.debug & 7) == 2) nTB = 0;
    if ((p.debug & 7) == 3) nTA = 0;
    const bool isA = lane >= 16;
    unsigned ia = 0, ib = 0, iss = 0;
    while (ia < nTA || ib < nTB) {
      const unsigned idx = (isA ? ia : ib) + (unsigned)(lane & 15);
      const bool valid = idx < (isA ? nTA : nTB);
      const unsigned g = blockIdx.x + idx * gridDim.x;
      const unsigned per = isA ? nA : nB;
      const int band = valid ? (int)(g / per) : 0;
      const unsigned tile = valid ? g % per : 0u;
      // what the tile waits for: phase B -- every phase-A tile of its band has been stored; phase A -- the band that
      // used its scratch slot before has been pulled completely
      bool ready = false;
      if (valid) {
        if (p.debug & 7) ready = true;
        else if (!isA) ready = ld_acquire_gpu(counters + band) >= nA;
        else ready = band < p.nslots || ld_acquire_gpu(counters + p.nbands + (band - p.nslots)) >= nB;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, ready);
      unsigned nbr = (unsigned)__ffs((int)(~mask & 0xffffu | 0x10000u)) - 1u;          // ready prefix of the B window
      unsigned nar = (unsigned)__ffs((int)(~(mask >> 16) & 0xffffu | 0x10000u)) - 1u;  // ... of the A window
      // never more than NST copies per round: two lanes of one round must not meet on the same stage (the later one's parity
      // wait on `empty` could be satisfied by the phase BEFORE the one the earlier lane is still waiting for)
      if (nar > (unsigned)P::A_RUN) nar = (unsigned)P::A_RUN;
      if (nbr > (unsigned)(NST - P::A_RUN)) nbr = (unsigned)(NST - P::A_RUN);
      if (nbr == 0 && nar == 0) { __nanosleep(40); continue; }
      const unsigned mine = isA ? (unsigned)(lane - 16) : (unsigned)lane;
      if (mine < (isA ? nar : nbr)) {
        const unsigned i = iss + (isA ? nbr + mine : mine);    // issue sequence number -> landing stage
        const int st = (int)(i % NST);
        if (i >= (unsigned)NST) mbar_wait(&empty[st], ((i / NST) - 1) & 1);
        const int bo = band / p.nbi, bi = band % p.nbi;
        desc[st] = make_int4(isA ? 0 : 1, band, (int)tile, (int)i);
        if (p.debug & 32) {
          mbar_arrive(&full[st]);     // experiment: no copy at all
        } else if (isA) {
          mbar_expect_tx(&full[st], (uint32_t)P::A_BYTES);
          const int u = (int)(tile / (unsigned)p.a_ncg), cg = (int)(tile % (unsigned)p.a_ncg);
          if constexpr (P::MODE == MODE_STRIDED)   // dims {2 I, N2, N1, O}: box {2 TLA, 1, N1, 1}
            tma_load_4d(stage_ptr(st), &tmA, 2 * (bi * p.wb + cg * TLA), u, 0, bo, &full[st], pol);
          else                                     // dims {2 N2, N1, R, 1}: box {2 TLA, N1, 1, 1}; u = row inside the band
            tma_load_4d(stage_ptr(st), &tmA, 2 * cg * TLA, 0, band * TLB + u, 0, &full[st], pol);
        } else {
          mbar_expect_tx(&full[st], (uint32_t)P::B_BYTES);
          const C* sl = slots + (long long)(band % p.nslots) * p.slot_elems;
          if constexpr (P::MODE == MODE_STRIDED) {
            // slot layout [cB][k1][n2][TLB]; tile = cB * N1 + k1: one contiguous block
            bulk_g2s(stage_ptr(st), sl + (long long)tile * (N2 * TLB), (uint32_t)P::B_BYTES, &full[st]);
          } else {
            // slot layout [row][k1][n2]; tile = k1: TLB segments of N2 elements -> landing [row][n2]
            for (int r = 0; r < TLB; r++)
              bulk_g2s(stage_ptr(st) + (size_t)r * N2 * P::ESZ, sl + ((long long)r * N1 + tile) * N2, (uint32_t)(N2 * P::ESZ), &full[st]);
          }
        }
      }
      __syncwarp();
      iss += nbr + nar; ib += nbr; ia += nar;
    }
    if (lane == 0) {
      unsigned i = iss;
      for (int gq = 0; gq < G; gq++, i++) {   // one stop marker per group
        const int st = (int)(i % NST);
        if (i >= (unsigned)NST) mbar_wait(&empty[st], ((i / NST) - 1) & 1);
        desc[st] = make_int4(-1, 0, 0, (int)i);
        mbar_arrive(&full[st]);
      }
    }
    return;
  }

  if (threadIdx.x < P::CTRL) {
    // ================================ SIGNALLER (one warp) ================================
    // Publishing a stored phase-A tile needs a gpu-scope fence (~2000 cycles: the stores must have reached L2) before the
    // band's counter is bumped.  Done by a compute thread it sat on the group's critical path at every tile; here lane g
    // watches group g's `stored` count, ONE fence per round covers whatever the groups have finished meanwhile.  The
    // proxy fence orders the generic-proxy stores before the async-proxy (bulk copy) reads that the counter releases.
    const int lane = threadIdx.x - 32;
    int seen = 0;
    bool fin = lane >= G;
    while (true) {
      int cnt = seen;
      if (!fin) {
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(cnt) : "r"(smem_u32((const void*)&sig[lane * 16])) : "memory");
      }
      const bool pending = cnt > seen;
      if (__any_sync(0xffffffffu, pending)) {
        __threadfence();
        fence_proxy_async_all();
        for (; seen < cnt; seen++) red_add_relaxed_gpu(counters + sig[lane * 16 + 8 + (seen & 7)], 1u);
        // the group may re-use ring entries below this (release: the reads of the entries above are ordered before it)
        if (!fin) asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32((const void*)&sig[lane * 16 + 3])), "r"(seen) : "memory");
      } else {
        int left = 0;
        if (!fin) asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(left) : "r"(smem_u32((const void*)&sig[lane * 16 + 1])) : "memory");
        if (!fin && left != 0) {   // the group has left: one more look at its count, then done
          asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(cnt) : "r"(smem_u32((const void*)&sig[lane * 16])) : "memory");
          if (cnt == seen) fin = true;
        }
        if (__all_sync(0xffffffffu, fin)) break;
        __nanosleep(100);
      }
    }
    return;
  }

  // ================================ COMPUTE GROUPS ================================
  const int grp = (threadIdx.x - P::CTRL) / GT;
  const int tid = (threadIdx.x - P::CTRL) % GT;
  const int gbar = 1 + grp;
  const C* stwA = stw;
  const C* stwB = stw + KA::TW_LEN;
  C* xch = reinterpret_cast<C*>(smem_raw + P::XCH_OFF + P::XCH_BYTES * grp);
  const unsigned lomask = (1u << p.tw_lo_bits) - 1u;
  const C* s_lo = stw + KA::TW_LEN + KB::TW_LEN;
  const C* s_hi = s_lo + p.tw_lo_n;
  auto root = [&](unsigned x) { return cmul(s_lo[x & lomask], s_hi[x >> p.tw_lo_bits]); };

  int na_done = 0;   // phase-A tiles this group has stored
  int* claim = const_cast<int*>(sig) + G * 16;     // next unclaimed issue sequence number
  // whichever group is free takes the next tile in issue order (a fixed tile -> group map stalled the loader on the
  // busiest group's stages while the others starved); the claim for the NEXT tile rides on a barrier of the current one
  if (tid == 0) sig[grp * 16 + 2] = atomicAdd(claim, 1);
  group_bar(gbar, GT);
  unsigned seq = (unsigned)sig[grp * 16 + 2];
  while (true) {
    const int st = (int)(seq % NST);
    // Up to 2 G claims are outstanding (every group holds its current tile and the pre-claimed next one), which can be
    // more than NST: this tile's stage may still be waiting for its PREVIOUS use to land, and a parity wait cannot tell
    // the phase before that one from the phase it wants.  The loader therefore stamps every stage with the sequence
    // number it carries, and a group that woke up one use too early simply waits again (rare).
    int4 d;
    while (true) {
      mbar_wait(&full[st], (seq / NST) & 1);
      d = desc[st];
      if ((unsigned)d.w == seq) break;
      __nanosleep(64);
    }
    if (d.x < 0) {
      if (tid == 0) asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32((const void*)&sig[grp * 16 + 1])), "r"(1) : "memory");
      break;
    }
    const int band = d.y;
    const unsigned tile = (unsigned)d.z;
    const C* land = reinterpret_cast<const C*>(stage_ptr(st));
    const int bo = band / p.nbi, bi = band % p.nbi;

    if (d.x == 0) {
      // ------------------------------- phase A: N1-point columns, TLA lines -------------------------------
      using K = KA;
      const int l = tid % K::TL, t = tid / K::TL;
      C v[K::E];
      static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = land[(t + e * K::TPT) * K::TL + l]; });
      mbar_arrive(&empty[st]);                  // pulled: the loader may refill the stage
      if (p.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
      const bool nomath = (p.debug & 8) != 0, nostore = (p.debug & 16) != 0;   // developer timing experiments (results invalid)
      if (!nomath) {
        run_stage<K, 0, C, false>(v, t, stwA);
        scatter<K, 0, true>(v, xch, l, t);
        static_for<1, K::S - 1>([&](auto sc) {
          constexpr int s = sc;
          group_bar(gbar, GT);
          gather<K, true>(v, xch, l, t);
          run_stage<K, s, C, false>(v, t, stwA);
          group_bar(gbar, GT);
          scatter<K, s, true>(v, xch, l, t);
        });
      }
      if (tid == 0) sig[grp * 16 + 2] = atomicAdd(claim, 1);
      group_bar(gbar, GT);
      if (!nomath) gather<K, true>(v, xch, l, t);
      seq = (unsigned)sig[grp * 16 + 2];
      if (!nomath) run_stage<K, K::S - 1, C, false>(v, t, stwA);
      const int u = (int)(tile / (unsigned)p.a_ncg), cg = (int)(tile % (unsigned)p.a_ncg);
      // inner four-step twiddle w_{N1 N2}^(k1 * n2), k1 = t + e*TPT: an anchor per 8 points + a running product
      const unsigned m = (P::MODE == MODE_STRIDED) ? (unsigned)u : (unsigned)(cg * TLA + l);
      if (!nomath) {
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
      C* sl = slots + (long long)(band % p.nslots) * p.slot_elems;
      if (nostore) {
        if (v[0].x == (T)123.456f && v[1].y == (T)654.321f) sl[tid] = v[2];   // (keeps the values alive)
      } else if constexpr (P::MODE == MODE_STRIDED) {
        // slot [cB][k1][n2][TLB], column c = cg*TLA + l of the band, n2 = u
        const int c = cg * TLA + l;
        C* op = sl + ((long long)(c / TLB) * N1 + t) * (N2 * TLB) + u * TLB + (c % TLB);
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; op[(long long)e * K::TPT * (N2 * TLB)] = v[e]; });
      } else {
        // slot [row][k1][n2], row = u, n2 = cg*TLA + l
        C* op = sl + ((long long)u * N1 + t) * N2 + cg * TLA + l;
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; op[(long long)e * K::TPT * N2] = v[e]; });
      }
      group_bar(gbar, GT);                      // all of the tile's stores have been issued; the exchange space is free again
      if (tid == 0) {                           // hand the tile to the signaller warp (fence + counter bump off this path)
        // single-producer / single-consumer ring of 8 band numbers with the signaller warp (it is normally at most 1-2 behind):
        // entries are published by the release store of the count below and recycled through the signaller's release store of
        // its read position -- ordered by acquire / release on the two counters, not by a barrier (compute-sanitizer's racecheck
        // only knows barriers and reports these four accesses, profiles/r02_compute_sanitizer.txt)
        int consumed;
        do {
          asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(consumed) : "r"(smem_u32((const void*)&sig[grp * 16 + 3])) : "memory");
          if (na_done - consumed >= 8) __nanosleep(50);
        } while (na_done - consumed >= 8);
        sig[grp * 16 + 8 + (na_done & 7)] = band;
        asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32((const void*)&sig[grp * 16])), "r"(na_done + 1) : "memory");
      }
      na_done++;
    } else {
      // ------------------------------- phase B: N2 points, TLB lines -------------------------------
      using K = KB;
      constexpr bool ROWS = P::B_ROWS;
      int l, t;
      if constexpr (ROWS) { t = tid % K::TPT; l = tid / K::TPT; } else { l = tid % K::TL; t = tid / K::TL; }
      C v[K::E];
      if constexpr (ROWS) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = land[l * K::N + t + e * K::TPT]; });
      else static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = land[(t + e * K::TPT) * K::TL + l]; });
      mbar_arrive(&empty[st]);
      if (tid == 0) red_add_relaxed_gpu(counters + p.nbands + band, 1u);   // the slot's data for this tile has been copied out
      const bool nomath = (p.debug & 8) != 0, nostore = (p.debug & 16) != 0;   // developer timing experiments (results invalid)
      if (!nomath) {
        run_stage<K, 0, C, false>(v, t, stwB);
        scatter<K, 0, !ROWS>(v, xch, l, t);
        static_for<1, K::S - 1>([&](auto sc) {
          constexpr int s = sc;
          group_bar(gbar, GT);
          gather<K, !ROWS>(v, xch, l, t);
          run_stage<K, s, C, false>(v, t, stwB);
          group_bar(gbar, GT);
          scatter<K, s, !ROWS>(v, xch, l, t);
        });
      }
      if (tid == 0) sig[grp * 16 + 2] = atomicAdd(claim, 1);
      group_bar(gbar, GT);
      if constexpr (ROWS) { l = tid % K::TL; t = tid / K::TL; }   // store mapping: line-fastest
      if (!nomath) gather<K, !ROWS>(v, xch, l, t);
      seq = (unsigned)sig[grp * 16 + 2];
      group_bar(gbar, GT);                      // every thread has gathered: the exchange space is free for the next tile
      if (!nomath) run_stage<K, K::S - 1, C, false>(v, t, stwB);
      // output index k = k1 + N1*k2, k2 = t + e*TPT
      int k1, cb;          // STRIDED: tile = cB*N1 + k1; ROWS: tile = k1
      if constexpr (P::MODE == MODE_STRIDED) { cb = (int)(tile / (unsigned)N1); k1 = (int)(tile % (unsigned)N1); } else { cb = 0; k1 = (int)tile; }
      if constexpr (P::OUTER) {
        // outer four-step twiddle w_L^(k * col), col = the band's first column + cB*TLB + l
        const unsigned omask = (1u << p.otw_lo_bits) - 1u;
        auto oroot = [&](unsigned long long x) {
          return cmul(__ldg(otw_lo + ((unsigned)x & omask)), __ldg(otw_hi + (unsigned)(x >> p.otw_lo_bits)));
        };
        const unsigned long long col = (unsigned long long)(p.otw_col0 + (long long)bi * p.wb + cb * TLB + l);
        constexpr int CH = (K::E < 8) ? K::E : 8;
        const C stepw = oroot((unsigned long long)(N1 * K::TPT) * col);
        static_for<0, K::E / CH>([&](auto qc) {
          constexpr int q = qc;
          C w = oroot((unsigned long long)(k1 + N1 * (t + q * CH * K::TPT)) * col);
          static_for<0, CH>([&](auto rc) {
            constexpr int e = q * CH + rc;
            v[e] = cmul(v[e], w);
            if constexpr (rc + 1 < CH) w = cmul(w, stepw);
          });
        });
      }
      const T sy = p.swap_out ? -scale : scale;
      if (scale != (T)1 || p.swap_out) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= sy; });
      C* op = out + (long long)bo * p.out_bo + (long long)bi * p.out_bi + (long long)(k1 + N1 * t) * p.out_ks + cb * TLB + l;
      const unsigned step_b = (unsigned)((long long)N1 * K::TPT * p.out_ks * (long long)sizeof(C));
      char* pb = reinterpret_cast<char*>(op);
      if (nostore) {
        if (v[0].x == (T)123.456f && v[1].y == (T)654.321f) *op = v[2];   // (keeps the values alive)
      } else {
        static_for<0, K::E>([&](auto ec) {
          constexpr int e = ec;
          st_stream(reinterpret_cast<C*>(pb + (unsigned long long)(unsigned)e * step_b), v[e]);
        });
      }
    }
  }
}

}  // namespace b200fft
