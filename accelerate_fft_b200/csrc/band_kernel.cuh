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
//   * G compute groups of 128 threads.  A group owns NSTG landing stages; it pulls a landed tile into registers,
//     runs the register-radix stages with the exchange done IN the stage buffer (named barriers per group), hands
//     the stage back to the loader right after its last gather, and finishes from registers: phase A multiplies by
//     the four-step twiddle and stores into the scratch slot, then one thread fences and bumps the band's counter;
//     phase B stores to the output (evict-first).
// With NST = G*NSTG stages (12 x 16 KB) in flight per SM the HBM reads never wait for the arithmetic, and the compute
// warps execute no global loads, no address arithmetic for them and no dependency checks.
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
  int otw_lo_bits;               // OUTER: outer table split
  long long otw_col0;            // OUTER: column index of band (bo = 0, bi = 0), columns advance by Wb per bi
  int swap_in, swap_out;         // conjugate on load (phase A) / on store (phase B)
  int wb;                        // STRIDED: columns per band
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

template <class KA_, class KB_, int MODE_, bool OUTER_, int G_, int NSTG_>
struct BandCfg {
  using KA = KA_;
  using KB = KB_;
  using real = typename KA_::real;
  static constexpr int MODE = MODE_;
  static constexpr bool OUTER = OUTER_;
  static constexpr int G = G_, NSTG = NSTG_, NST = G_ * NSTG_;
  static constexpr int GT = KA::THREADS;             // threads of one compute group
  static constexpr int THREADS = 32 + G_ * GT;       // warp 0 = loader
  static constexpr int N1 = KA::N, N2 = KB::N, TLA = KA::TL, TLB = KB::TL;
  static constexpr int ESZ = KA::ESZ;
  static constexpr bool B_ROWS = (MODE_ == MODE_ROWS);
  static constexpr int A_ELEMS = KA::COL_ELEMS > N1 * TLA ? KA::COL_ELEMS : N1 * TLA;
  static constexpr int B_LAY = B_ROWS ? KB::ROW_ELEMS : KB::COL_ELEMS;
  static constexpr int B_ELEMS = B_LAY > N2 * TLB ? B_LAY : N2 * TLB;
  static constexpr size_t STAGE_BYTES = (((size_t)(A_ELEMS > B_ELEMS ? A_ELEMS : B_ELEMS) * ESZ + 127) / 128) * 128;
  static constexpr int A_BYTES = N1 * TLA * ESZ, B_BYTES = N2 * TLB * ESZ;
  static constexpr int TW_ELEMS = KA::TW_LEN + KB::TW_LEN;
  static constexpr size_t TW_OFF = STAGE_BYTES * NST;
  static constexpr size_t DESC_OFF = TW_OFF + (((size_t)TW_ELEMS * ESZ + 127) / 128) * 128;
  static constexpr size_t BAR_OFF = DESC_OFF + 16 * NST;
  static constexpr size_t SMEM = BAR_OFF + 16 * NST + 16;
  static_assert(KA::THREADS == KB::THREADS, "both phases run in the same group shape");
  static_assert(KA::S >= 2 && KB::S >= 2, "both phases exchange through the stage buffer");
  static_assert(G_ <= 14, "one named barrier per group");
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
  constexpr int G = P::G, NSTG = P::NSTG, NST = P::NST, GT = P::GT;
  constexpr int N1 = P::N1, N2 = P::N2, TLA = P::TLA, TLB = P::TLB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* stw = reinterpret_cast<C*>(smem_raw + P::TW_OFF);           // stage twiddles of KA, then of KB
  int4* desc = reinterpret_cast<int4*>(smem_raw + P::DESC_OFF);  // per stage: {phase (-1 = stop), band, tile, 0}
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + P::BAR_OFF);
  uint64_t* empty = full + NST;

  for (int i = threadIdx.x; i < KA::TW_LEN; i += P::THREADS) stw[i] = twsA[i];
  for (int i = threadIdx.x; i < KB::TW_LEN; i += P::THREADS) stw[KA::TW_LEN + i] = twsB[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], GT); }
    fence_mbar_init();
  }
  __syncthreads();

  const unsigned nA = (unsigned)p.nA, nB = (unsigned)p.nB, seg = nA + nB;
  const unsigned total = (unsigned)p.nbands * seg;
  auto stage_of = [](unsigned i) -> int { return (int)(i % G) + G * (int)((i / G) % NSTG); };
  auto stage_ptr = [&](int st) -> unsigned char* { return smem_raw + P::STAGE_BYTES * st; };

  if (threadIdx.x < 32) {
    // ================================ LOADER ================================
    if (threadIdx.x != 0) return;
    const uint64_t pol = policy_evict_first();
    const unsigned head = (unsigned)p.la * nA;
    unsigned i = 0;
    for (unsigned ticket = blockIdx.x; ticket < total; ticket += gridDim.x, i++) {
      // decode: A(0..la-1) | [A(j) B(j-la)] for j = la..nb-1 | B(nb-la..nb-1)
      int phase, band;
      unsigned tile;
      if (ticket < head) { phase = 0; band = (int)(ticket / nA); tile = ticket % nA; }
      else {
        const unsigned r0 = ticket - head;
        const int j = (int)(r0 / seg) + p.la;
        if (j < p.nbands) {
          const unsigned r = r0 % seg;
          if (r < nA) { phase = 0; band = j; tile = r; } else { phase = 1; band = j - p.la; tile = r - nA; }
        } else {
          const unsigned r1 = r0 - (unsigned)(p.nbands - p.la) * seg;
          phase = 1; band = p.nbands - p.la + (int)(r1 / nB); tile = r1 % nB;
        }
      }
      const int st = stage_of(i);
      if (i >= (unsigned)NST) mbar_wait(&empty[st], ((i / NST) - 1) & 1);
      const int bo = band / p.nbi, bi = band % p.nbi;
      if (phase == 0) {
        if (band >= p.nslots) {   // the band that used this scratch slot before must have been pulled completely
          const unsigned* f = counters + p.nbands + (band - p.nslots);
          while (ld_acquire_gpu(f) < nB) __nanosleep(40);
        }
        desc[st] = make_int4(0, band, (int)tile, 0);
        mbar_expect_tx(&full[st], (uint32_t)P::A_BYTES);
        const int u = (int)(tile / (unsigned)p.a_ncg), cg = (int)(tile % (unsigned)p.a_ncg);
        if constexpr (P::MODE == MODE_STRIDED)   // dims {2 I, N2, N1, O}: box {2 TLA, 1, N1, 1}
          tma_load_4d(stage_ptr(st), &tmA, 2 * (bi * p.wb + cg * TLA), u, 0, bo, &full[st], pol);
        else                                     // dims {2 N2, N1, R, 1}: box {2 TLA, N1, 1, 1}; u = row inside the band
          tma_load_4d(stage_ptr(st), &tmA, 2 * cg * TLA, 0, band * TLB + u, 0, &full[st], pol);
      } else {
        const unsigned* f = counters + band;
        while (ld_acquire_gpu(f) < nA) __nanosleep(40);
        fence_proxy_async_all();
        desc[st] = make_int4(1, band, (int)tile, 0);
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
    for (int gq = 0; gq < G; gq++, i++) {   // one stop marker per group
      const int st = stage_of(i);
      if (i >= (unsigned)NST) mbar_wait(&empty[st], ((i / NST) - 1) & 1);
      desc[st] = make_int4(-1, 0, 0, 0);
      mbar_arrive(&full[st]);
    }
    return;
  }

  // ================================ COMPUTE GROUPS ================================
  const int grp = (threadIdx.x - 32) / GT;
  const int tid = (threadIdx.x - 32) % GT;
  const int gbar = 1 + grp;
  const C* stwA = stw;
  const C* stwB = stw + KA::TW_LEN;
  const unsigned lomask = (1u << p.tw_lo_bits) - 1u;
  auto root = [&](unsigned x) { return cmul(__ldg(tw_lo + (x & lomask)), __ldg(tw_hi + (x >> p.tw_lo_bits))); };

  for (unsigned k = 0;; k++) {
    const int st = grp + G * (int)(k % NSTG);
    mbar_wait(&full[st], (k / NSTG) & 1);
    const int4 d = desc[st];
    if (d.x < 0) break;
    const int band = d.y;
    const unsigned tile = (unsigned)d.z;
    C* sm = reinterpret_cast<C*>(stage_ptr(st));
    const int bo = band / p.nbi, bi = band % p.nbi;

    if (d.x == 0) {
      // ------------------------------- phase A: N1-point columns, TLA lines -------------------------------
      using K = KA;
      const int l = tid % K::TL, t = tid / K::TL;
      C v[K::E];
      static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = sm[(t + e * K::TPT) * K::TL + l]; });
      if (p.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
      run_stage<K, 0, C, false>(v, t, stwA);
      group_bar(gbar, GT);                      // every thread has pulled its points: the stage becomes the exchange space
      scatter<K, 0, true>(v, sm, l, t);
      static_for<1, K::S - 1>([&](auto sc) {
        constexpr int s = sc;
        group_bar(gbar, GT);
        gather<K, true>(v, sm, l, t);
        run_stage<K, s, C, false>(v, t, stwA);
        group_bar(gbar, GT);
        scatter<K, s, true>(v, sm, l, t);
      });
      group_bar(gbar, GT);
      gather<K, true>(v, sm, l, t);
      fence_proxy_async();                      // generic-proxy accesses of the stage are ordered before the next TMA write
      mbar_arrive(&empty[st]);                  // the loader may refill the stage
      run_stage<K, K::S - 1, C, false>(v, t, stwA);
      const int u = (int)(tile / (unsigned)p.a_ncg), cg = (int)(tile % (unsigned)p.a_ncg);
      // inner four-step twiddle w_{N1 N2}^(k1 * n2), k1 = t + e*TPT: an anchor per 8 points + a running product
      const unsigned m = (P::MODE == MODE_STRIDED) ? (unsigned)u : (unsigned)(cg * TLA + l);
      {
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
      if constexpr (P::MODE == MODE_STRIDED) {
        // slot [cB][k1][n2][TLB], column c = cg*TLA + l of the band, n2 = u
        const int c = cg * TLA + l;
        C* op = sl + ((long long)(c / TLB) * N1 + t) * (N2 * TLB) + u * TLB + (c % TLB);
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; op[(long long)e * K::TPT * (N2 * TLB)] = v[e]; });
      } else {
        // slot [row][k1][n2], row = u, n2 = cg*TLA + l
        C* op = sl + ((long long)u * N1 + t) * N2 + cg * TLA + l;
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; op[(long long)e * K::TPT * N2] = v[e]; });
      }
      group_bar(gbar, GT);                      // all of the tile's stores have been issued
      if (tid == 0) { __threadfence(); red_add_relaxed_gpu(counters + band, 1u); }
    } else {
      // ------------------------------- phase B: N2 points, TLB lines -------------------------------
      using K = KB;
      constexpr bool ROWS = P::B_ROWS;
      if (tid == 0) red_add_relaxed_gpu(counters + p.nbands + band, 1u);   // the slot's data for this tile is in shared memory
      int l, t;
      if constexpr (ROWS) { t = tid % K::TPT; l = tid / K::TPT; } else { l = tid % K::TL; t = tid / K::TL; }
      C v[K::E];
      if constexpr (ROWS) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = sm[l * K::N + t + e * K::TPT]; });
      else static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e] = sm[(t + e * K::TPT) * K::TL + l]; });
      run_stage<K, 0, C, false>(v, t, stwB);
      group_bar(gbar, GT);
      scatter<K, 0, !ROWS>(v, sm, l, t);
      static_for<1, K::S - 1>([&](auto sc) {
        constexpr int s = sc;
        group_bar(gbar, GT);
        gather<K, !ROWS>(v, sm, l, t);
        run_stage<K, s, C, false>(v, t, stwB);
        group_bar(gbar, GT);
        scatter<K, s, !ROWS>(v, sm, l, t);
      });
      group_bar(gbar, GT);
      if constexpr (ROWS) { l = tid % K::TL; t = tid / K::TL; }   // store mapping: line-fastest
      gather<K, !ROWS>(v, sm, l, t);
      fence_proxy_async();
      mbar_arrive(&empty[st]);
      run_stage<K, K::S - 1, C, false>(v, t, stwB);
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
      static_for<0, K::E>([&](auto ec) {
        constexpr int e = ec;
        st_stream(reinterpret_cast<C*>(pb + (unsigned long long)(unsigned)e * step_b), v[e]);
      });
    }
  }
}

}  // namespace b200fft
