// Persistent, software-pipelined column kernel -- optionally spanning a thread-block cluster.
//
// The plain column kernel (fft_kernel.cuh) and the first cluster kernel (cluster_kernel.cuh) run load -> compute ->
// store in lock step; with 64-128 KB tiles only one or two CTAs fit on an SM, so HBM sees loads for a fraction of
// the time (ncu: long_scoreboard + lg_throttle top stalls, 44-67 % of HBM peak; profiles/r01_cluster_ablation.txt).
// Here ONE CTA per SM stays resident and walks over its tiles:
//   * a LANDING buffer (one tile, NQ parts with an mbarrier each) is filled by cp.async (LDGSTS.128, L1 bypassed);
//     the moment a part has been pulled into registers it is re-armed with the same part of the CTA's NEXT tile,
//     so the HBM reads never stop;
//   * G = 2 thread groups alternate over the tiles (named barriers, a turn token keeps the landing buffer FIFO);
//     each group owns an exchange buffer for the Stockham stages, so one group's butterflies overlap the other's
//     waits, and every store to HBM goes straight from registers;
//   * CS > 1: the CTA is one of a cluster of CS that transforms columns of length N = N1*CS together (four-step
//     inside the cluster as in cluster_kernel.cuh).  The exchange uses st.async: the data AND its completion
//     count travel to the destination CTA's mbarrier, so there is no cluster-wide barrier in steady state -- a
//     group only waits for "all CS peers have drained their exchange buffer" (remote mbarrier arrives) and for
//     "my N1*TL values have landed" (transaction bytes).
// One HBM read + one HBM write per element.
//
// Replaces, for its share of a plan, what cufftExecC2C / cufftExecZ2Z did behind
// /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs:112-124.
#pragma once
#include <cstdint>

#include "cluster_kernel.cuh"

namespace b200fft {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier gets one (pre-counted) arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {}
}
// store into a peer CTA's shared memory; the peer's mbarrier is credited with the bytes when they have landed
__device__ __forceinline__ void st_async(uint32_t addr, float2 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(v.x),
               "f"(v.y), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async(uint32_t addr, double2 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr), "d"(v.x),
               "d"(v.y), "r"(mbar)
               : "memory");
}

// ROWS: contiguous lines (one line per tile, row layout in the exchange buffer) instead of strided columns
template <class K_, int CS_, int NQ_ = 4, bool ROWS_ = false>
struct PipeCfg {
  using K = K_;
  static constexpr int CS = CS_, G = 2, NQ = NQ_;
  static constexpr bool ROWS = ROWS_;
  static_assert(!ROWS_ || (CS_ == 1 && K_::TL == 1), "row mode: one contiguous line per tile, no cluster");
  static constexpr int N1 = K::N, N = K::N * CS_;
  static constexpr int EP = K::E / CS_;          // CS > 1: phase-2 butterflies per thread
  static constexpr int KL = K::N / CS_;          // CS > 1: values of k1 owned by one CTA
  static constexpr int THREADS = G * K::THREADS;
  static constexpr int TILE_ELEMS = K::N * K::TL;
  static constexpr int TILE_BYTES = TILE_ELEMS * K::ESZ;
  static constexpr int ROW_CHUNKS = ROWS_ ? TILE_ELEMS * K::ESZ / 16 : K::TL * K::ESZ / 16;   // 16-byte chunks per contiguous run
  static constexpr int PART_CHUNKS = TILE_BYTES / 16 / NQ_;              // chunks per landing part
  static constexpr int CPT = PART_CHUNKS / K::THREADS;                   // cp.async per thread and part
  static constexpr int EQ = K::E / NQ_;                                  // registers fed by one part
  static constexpr int LAY_ELEMS = ROWS_ ? K::ROW_ELEMS : K::COL_ELEMS;
  static constexpr int XCH_ELEMS = (LAY_ELEMS > TILE_ELEMS) ? LAY_ELEMS : TILE_ELEMS;
  static constexpr size_t XCH_BYTES = (((size_t)XCH_ELEMS * K::ESZ + 127) / 128) * 128;
  // the stage twiddles and this rank's row of inner twiddles live in shared memory for the life of the CTA: their
  // global loads queued behind the cp.async stream in the LSU (lg_throttle 22 % of the stall samples)
  // (when the table fits: the three-stage row kernels keep theirs, 64 KB, in global memory)
  static constexpr bool TW_SMEM = (size_t)K::TW_LEN * K::ESZ <= 16384;
  static constexpr int TW_ELEMS = (TW_SMEM ? K::TW_LEN : 0) + (CS_ > 1 ? K::N : 0);
  static constexpr size_t TW_BYTES = (((size_t)TW_ELEMS * K::ESZ + 127) / 128) * 128;
  static constexpr size_t BAR_OFF = (size_t)TILE_BYTES + G * XCH_BYTES + TW_BYTES;
  static constexpr size_t SMEM = BAR_OFF + 8 * (NQ_ + 2 * G) + 16;
  static_assert(K::S >= 2, "needs at least one shared-memory exchange");
  static_assert(K::E % CS_ == 0 && K::E % NQ_ == 0, "E must split over the cluster and over the landing parts");
  static_assert((ROWS_ || (K::TL * K::ESZ) % 16 == 0) && PART_CHUNKS % K::THREADS == 0 && (ROWS_ || K::THREADS % ROW_CHUNKS == 0), "cp.async tiling");
  static_assert(CS_ == 1 || CS_ == 2 || CS_ == 4 || CS_ == 8 || CS_ == 16, "cluster size = register radix of phase 2");
};

// Geom: address(b, o, line, n) = b*bs + o*os + line + n*ns (ils == ols == 1), n < N = N1*CS; nl % TL == 0; the tile
// rows must be 16-byte aligned (planner).  Grid = (resident clusters) * CS, cluster dimension CS.
template <class P, bool TW4>
__global__ void __launch_bounds__(P::THREADS, 1)
fft_pipe_cols_kernel(const Geom g, const cpx_t<typename P::K::real>* __restrict__ in, cpx_t<typename P::K::real>* __restrict__ out,
                     const cpx_t<typename P::K::real>* __restrict__ tws, const cpx_t<typename P::K::real>* __restrict__ tw_lo,
                     const cpx_t<typename P::K::real>* __restrict__ tw_hi, typename P::K::real scale,
                     const cpx_t<typename P::K::real>* __restrict__ ctw) {
  using K = typename P::K;
  using T = typename K::real;
  using C = cpx_t<T>;
  constexpr int CS = P::CS, NQ = P::NQ, EQ = P::EQ;
  constexpr bool ROWS = P::ROWS, COL = !P::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* land = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P::BAR_OFF);
  C* stw = reinterpret_cast<C*>(smem_raw + P::TILE_BYTES + P::G * P::XCH_BYTES);   // stage twiddles, then inner twiddles
  uint64_t* full = bars;                 // [NQ]  landing part q holds the current tile
  uint64_t* ready = bars + NQ;           // [G]   CS > 1: every CTA of the cluster has drained its exchange buffer g
  uint64_t* landed = bars + NQ + P::G;   // [G]   CS > 1: the exchanged values for group g have arrived here

  const int grp = threadIdx.x / K::THREADS;
  const int tid = threadIdx.x % K::THREADS;
  const int l = ROWS ? 0 : tid % K::TL, t = ROWS ? tid : tid / K::TL;
  const unsigned rank = (CS > 1) ? cluster_ctarank() : 0u;
  const unsigned q0 = (CS > 1) ? cluster_id_x() : blockIdx.x;
  const unsigned nclus = gridDim.x / CS;
  const unsigned ntiles = (unsigned)g.nb * (unsigned)g.no * (unsigned)g.ntl;
  const int nk = (q0 < ntiles) ? (int)((ntiles - q0 + nclus - 1) / nclus) : 0;   // tiles of this CTA: q0 + j*nclus
  C* xch = reinterpret_cast<C*>(smem_raw + P::TILE_BYTES + (size_t)grp * P::XCH_BYTES);

  constexpr int TWL = P::TW_SMEM ? K::TW_LEN : 0;   // inner twiddles follow the stage table
  if constexpr (P::TW_SMEM)
    for (int i = threadIdx.x; i < K::TW_LEN; i += P::THREADS) stw[i] = tws[i];
  const C* stage_tw = P::TW_SMEM ? stw : tws;
  if constexpr (CS > 1)
    if (rank != 0)
      for (int i = threadIdx.x; i < K::N; i += P::THREADS) stw[TWL + i] = ctw[(size_t)(rank - 1) * K::N + i];
  if (threadIdx.x == 0) {
    for (int q = 0; q < NQ; q++) mbar_init(&full[q], K::THREADS);
    for (int i = 0; i < P::G; i++) { mbar_init(&ready[i], CS); mbar_init(&landed[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  if constexpr (CS > 1) { cluster_arrive(); cluster_wait(); }   // peers' barriers exist before anything is sent to them

  auto tile_coords = [&](int j, int& lt, int& o, int& b) {
    const unsigned tile = q0 + (unsigned)j * nclus;
    if (g.no == 1 && g.nb == 1) { lt = (int)tile; o = 0; b = 0; }
    else {
      lt = tile % (unsigned)g.ntl;
      const unsigned rest = tile / (unsigned)g.ntl;
      o = rest % (unsigned)g.no;
      b = rest / (unsigned)g.no;
    }
  };
  // part q of tile j: rows [q*N1/NQ, (q+1)*N1/NQ) of this CTA's share (row n1 of the CTA = row n1*CS + rank of the
  // column), TL contiguous elements each; the calling group's threads copy CPT 16-byte chunks each
  const uint32_t land_u32 = smem_u32(land);
  const unsigned long long row_b = (unsigned long long)((long long)CS * g.ins * (long long)sizeof(C));
  auto issue_part = [&](int j, int q) {
    int lt, o, b;
    tile_coords(j, lt, o, b);
    const char* base = reinterpret_cast<const char*>(in + (long long)b * g.ibs + (long long)o * g.ios +
                                                     (ROWS ? (long long)lt * g.ils : (long long)lt * K::TL + (long long)rank * g.ins));
    static_for<0, P::CPT>([&](auto ic) {
      constexpr int i = ic;
      const int c = q * P::PART_CHUNKS + i * K::THREADS + tid;
      if constexpr (ROWS) cp_async16(land_u32 + (uint32_t)c * 16u, base + (unsigned)c * 16u);
      else {
        const int row = c / P::ROW_CHUNKS, sub = c % P::ROW_CHUNKS;
        cp_async16(land_u32 + (uint32_t)c * 16u, base + (unsigned long long)row * row_b + (unsigned)sub * 16u);
      }
    });
    cp_async_arrive_noinc(&full[q]);
  };

  if (grp == 0 && nk > 0)
    for (int q = 0; q < NQ; q++) issue_part(0, q);

  const int gbar = 1 + grp;                 // named barrier of this group
  const int my_turn = 3 + grp, other_turn = 3 + (1 - grp);

  for (int j = grp; j < nk; j += P::G) {
    int lt, o, b;
    tile_coords(j, lt, o, b);
    const int line = lt * K::TL + l;
    const uint32_t par = (uint32_t)((j / P::G) & 1);   // phase of this group's ready / landed barriers
    if constexpr (CS > 1)
      if (tid == 0) mbar_expect_tx(&landed[grp], (uint32_t)P::TILE_BYTES);

    // ---- pull the tile out of the landing buffer, part by part, re-arming each part with the next tile ----
    C v[K::E];
    if (j > 0) group_bar(my_turn, P::THREADS);         // the other group has taken tile j-1
    {
      const C* lp = reinterpret_cast<const C*>(land) + t * K::TL + l;
      static_for<0, NQ>([&](auto qc) {
        constexpr int q = qc;
        mbar_wait(&full[q], (uint32_t)(j & 1));
        static_for<0, EQ>([&](auto ec) { constexpr int e = q * EQ + ec; v[e] = lp[e * K::TPT * K::TL]; });
        group_bar(gbar, K::THREADS);
        if (j + 1 < nk) issue_part(j + 1, q);
      });
      if (j + 1 < nk) bar_arrive(other_turn, P::THREADS);
    }

    // ---- phase 1: N1-point transform in this group's exchange buffer -----------------------------------
    if (g.swap_in) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].y = -v[e].y; });
    run_stage<K, 0, C, !P::TW_SMEM>(v, t, stage_tw);
    scatter<K, 0, COL>(v, xch, l, t);
    static_for<1, K::S - 1>([&](auto sc) {
      constexpr int s = sc;
      group_bar(gbar, K::THREADS);
      gather<K, COL>(v, xch, l, t);
      run_stage<K, s, C, !P::TW_SMEM>(v, t, stage_tw);
      group_bar(gbar, K::THREADS);
      scatter<K, s, COL>(v, xch, l, t);
    });
    group_bar(gbar, K::THREADS);
    gather<K, COL>(v, xch, l, t);
    run_stage<K, K::S - 1, C, !P::TW_SMEM>(v, t, stage_tw);    // v[e] = output k1 = t + e*TPT of this CTA's N1-point transform

    if constexpr (CS == 1) {
      const unsigned step_b = ROWS ? (unsigned)(K::TPT * sizeof(C)) : (unsigned)((long long)K::TPT * g.ons * (long long)sizeof(C));
      C* op = out + (long long)b * g.obs + (long long)o * g.oos + (ROWS ? (long long)lt * g.ols + t : line + (long long)t * g.ons);
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
      const T sy = g.swap_out ? -scale : scale;
      if (scale != (T)1 || g.swap_out) static_for<0, K::E>([&](auto ec) { constexpr int e = ec; v[e].x *= scale; v[e].y *= sy; });
      if (!ROWS && !TW4 && g.npeers) {   // scatter store over peer memory (see Geom)
        const long long off = (long long)b * g.obs + (long long)o * g.oos + line;
        static_for<0, K::E>([&](auto ec) { constexpr int e = ec; *peer_addr<C>(g, off, t + e * K::TPT) = v[e]; });
      } else {
        char* p = reinterpret_cast<char*>(op);
        static_for<0, K::E>([&](auto ec) {
          constexpr int e = ec;
          st_stream(reinterpret_cast<C*>(p + (unsigned long long)(unsigned)e * step_b), v[e]);
        });
      }
    } else {
      constexpr int EP = P::EP, KL = P::KL;
      // this group has drained its exchange buffer: tell every CTA of the cluster (lanes 0..CS-1 of the first warp)
      group_bar(gbar, K::THREADS);
      if (tid < CS) mbar_arrive_remote(map_to_rank(smem_u32(&ready[grp]), (unsigned)tid));
      // inner four-step twiddle w_N^(k1 * rank): anchors from ctw[rank-1][k1] + running product (<= 8 ulp)
      if (rank != 0) {
        const C* wp = stw + TWL;
        constexpr int CH = (K::E < 8) ? K::E : 8;
        const C stepw = wp[K::TPT];
        static_for<0, K::E / CH>([&](auto qc) {
          constexpr int q = qc;
          C w = wp[t + q * CH * K::TPT];
          static_for<0, CH>([&](auto rc) {
            constexpr int e = q * CH + rc;
            v[e] = cmul(v[e], w);
            if constexpr (rc + 1 < CH) w = cmul(w, stepw);
          });
        });
      }
      // ---- exchange: register e goes to CTA e / EP, slot [rank][(e % EP)*TPT + t][l] of ITS buffer `grp` ----
      mbar_wait(&ready[grp], par);
      {
        const uint32_t base = smem_u32(xch) + (uint32_t)(((int)rank * KL + t) * K::TL + l) * (uint32_t)sizeof(C);
        const uint32_t lbar = smem_u32(&landed[grp]);
        static_for<0, CS>([&](auto dc) {
          constexpr int d = dc;
          const uint32_t ra = map_to_rank(base, (unsigned)d);
          const uint32_t rb = map_to_rank(lbar, (unsigned)d);
          static_for<0, EP>([&](auto jc) {
            constexpr int jj = jc;
            st_async(ra + (uint32_t)(jj * K::TPT * K::TL) * (uint32_t)sizeof(C), v[d * EP + jj], rb);
          });
        });
      }
      mbar_wait(&landed[grp], par);

      // ---- phase 2: radix-CS butterflies over n2 for k1 = rank*KL + t + jj*TPT, rows k1 + N1*k2 ------------
      const C* gp = xch + t * K::TL + l;
      C* op = out + (long long)b * g.obs + (long long)o * g.oos + line + ((long long)rank * KL + t) * g.ons;
      const unsigned jstep_b = (unsigned)((long long)K::TPT * g.ons * (long long)sizeof(C));
      const unsigned long long kstep_b = (unsigned long long)((long long)K::N * g.ons * (long long)sizeof(C));
      const T sy = g.swap_out ? -scale : scale;
      const bool do_scale = (scale != (T)1) || g.swap_out;
      static_for<0, EP>([&](auto jc) {
        constexpr int jj = jc;
        C a[CS];
        static_for<0, CS>([&](auto nc) { constexpr int n2 = nc; a[n2] = gp[(n2 * KL + jj * K::TPT) * K::TL]; });
        dft<CS>(a);
        if constexpr (TW4) {
          const unsigned m = g.tw_from_o ? (unsigned)o : (unsigned)line / (unsigned)g.tw_div;
          const unsigned lomask = (1u << g.tw_lo_bits) - 1u;
          auto root = [&](unsigned x) { return cmul(__ldg(tw_lo + (x & lomask)), __ldg(tw_hi + (x >> g.tw_lo_bits))); };
          C w = root((rank * KL + t + jj * K::TPT) * m);
          const C stepw = root((unsigned)K::N * m);
          static_for<0, CS>([&](auto kc) {
            constexpr int k2 = kc;
            a[k2] = cmul(a[k2], w);
            if constexpr (k2 + 1 < CS) w = cmul(w, stepw);
          });
        }
        if (do_scale) static_for<0, CS>([&](auto kc) { constexpr int k2 = kc; a[k2].x *= scale; a[k2].y *= sy; });
        char* p = reinterpret_cast<char*>(op) + (unsigned long long)(unsigned)jj * jstep_b;
        static_for<0, CS>([&](auto kc) {
          constexpr int k2 = kc;
          st_stream(reinterpret_cast<C*>(p + (unsigned long long)k2 * kstep_b), a[k2]);
        });
      });
    }
  }
  if constexpr (CS > 1) { cluster_arrive(); cluster_wait(); }   // no CTA leaves while a peer may still address it
}

}  // namespace b200fft
