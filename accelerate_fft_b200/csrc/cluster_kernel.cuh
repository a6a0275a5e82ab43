// Strided ("column") transforms too long for one CTA: a thread-block CLUSTER of CS CTAs transforms a tile of TL
// adjacent columns of length N = N1 * CS in ONE HBM pass, exchanging once through distributed shared memory.
//
// A column pass needs >= 64 B contiguous per row to use whole DRAM sectors, i.e. TL >= 8 (c64) columns per tile;
// at N = 8192 that is a 512 KB tile -- more than the 227 KB of one CTA, which is why such axes took two passes
// (four-step through HBM) before.  Here the four-step runs inside the cluster (n = n1*CS + n2, k = k1 + N1*k2):
//   phase 1  CTA r (= n2) loads rows n1*CS + r, transforms them over n1 with the ordinary register-radix /
//            shared-memory Stockham stages of Cfg K (N1 points, TL columns) and multiplies by w_N^(k1*r);
//   exchange every thread stores its results straight into the shared memory of the CTA that owns that k1
//            (st.shared::cluster; owner = k1 / (N1/CS) is a compile-time function of the register index);
//   phase 2  CTA r' does the radix-CS butterflies over n2 for its N1/CS values of k1 in registers and stores rows
//            k1 + N1*k2 -- TL contiguous elements each, like any column kernel.
// One HBM read + one HBM write per element; (CS-1)/CS of the elements cross the SM-to-SM network once.
//
// Replaces, for its share of a plan, what cufftExecC2C / cufftExecZ2Z did behind
// /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs:112-124.
#pragma once
#include <cstdint>

#include "ring_kernel.cuh"   // smem_u32

namespace b200fft {

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared::cta address -> the same offset in the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, unsigned rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double2 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <class K_, int CS_>
struct ClusterCfg {
  using K = K_;
  static constexpr int CS = CS_;
  static constexpr int N1 = K::N, N = K::N * CS_;
  static constexpr int EP = K::E / CS_;          // phase-2 butterflies per thread = registers per destination CTA
  static constexpr int KL = K::N / CS_;          // values of k1 owned by one CTA
  static_assert(K::E % CS_ == 0 && K::E >= CS_, "every thread must hold a whole number of registers per destination CTA");
  static_assert(CS_ == 2 || CS_ == 4 || CS_ == 8 || CS_ == 16, "cluster size = register radix of phase 2");
  // the receive buffer [n2][k1 local][line] aliases the phase-1 exchange space
  static constexpr int RECV_ELEMS = K::N * K::TL;
  static constexpr int SM_ELEMS = (K::S > 1 && K::COL_ELEMS > RECV_ELEMS) ? K::COL_ELEMS : RECV_ELEMS;
  static constexpr size_t SMEM = (size_t)SM_ELEMS * K::ESZ;
};

// Geom: address(b, o, line, n) = b*bs + o*os + line + n*ns  (ils == ols == 1), n < N = N1*CS, line < nl tiled by
// TL; the grid is ntiles * CS CTAs launched with cluster dimension CS.  TW4: the result is further multiplied by
// the four-step twiddle w_L^(k * m), m = line / tw_div (or o) as in fft_lines_tile.
template <class CC, bool TW4>
__global__ void __launch_bounds__(CC::K::THREADS, CC::K::MINB)
fft_cluster_cols_kernel(const Geom g, const cpx_t<typename CC::K::real>* __restrict__ in, cpx_t<typename CC::K::real>* __restrict__ out,
                        const cpx_t<typename CC::K::real>* __restrict__ tws, const cpx_t<typename CC::K::real>* __restrict__ tw_lo,
                        const cpx_t<typename CC::K::real>* __restrict__ tw_hi, typename CC::K::real scale,
                        const cpx_t<typename CC::K::real>* __restrict__ ctw) {
  using K = typename CC::K;
  using T = typename K::real;
  using C = cpx_t<T>;
  constexpr int CS = CC::CS, EP = CC::EP, KL = CC::KL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);

  const int tid = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  const unsigned tile = cluster_id_x();
  int lt, o, b;
  if (g.no == 1 && g.nb == 1) { lt = (int)tile; o = 0; b = 0; }
  else {
    lt = tile % (unsigned)g.ntl;
    const unsigned rest = tile / (unsigned)g.ntl;
    o = rest % (unsigned)g.no;
    b = rest / (unsigned)g.no;
  }
  const int l = tid % K::TL, t = tid / K::TL;
  const int line = lt * K::TL + l;

  // ---- phase 1: N1-point transforms over n1 of the rows n1*CS + rank ---------------------------------
  C v[K::E];
  {
    const int lline = min(line, g.nl - 1);   // ragged last tile: re-read the last valid column, never store it
    const C* ip = in + (long long)b * g.ibs + (long long)o * g.ios + lline + ((long long)t * CS + rank) * g.ins;
    const unsigned step_b = (unsigned)((long long)K::TPT * CS * g.ins * (long long)sizeof(C));
    auto head = [&](auto cj) {
      constexpr bool CJ = decltype(cj)::value;
      const char* p = reinterpret_cast<const char*>(ip);
      static_for<0, K::E>([&](auto ec) {
        constexpr int e = ec;
        C x = ld_stream(reinterpret_cast<const C*>(p + (unsigned long long)(unsigned)e * step_b));
        if constexpr (CJ) x.y = -x.y;
        v[e] = x;
      });
      run_stage<K, 0>(v, t, tws);
      if constexpr (K::S > 1) scatter<K, 0, true>(v, sm, l, t);
    };
    if (g.swap_in) head(std::true_type{}); else head(std::false_type{});
  }
  static_for<1, (K::S > 1 ? K::S - 1 : 1)>([&](auto sc) {
    constexpr int s = sc;
    __syncthreads();
    gather<K, true>(v, sm, l, t);
    run_stage<K, s>(v, t, tws);
    __syncthreads();
    scatter<K, s, true>(v, sm, l, t);
  });
  if constexpr (K::S > 1) {
    __syncthreads();
    gather<K, true>(v, sm, l, t);
  }
  if constexpr (K::S > 1) run_stage<K, K::S - 1>(v, t, tws);
  // this CTA no longer reads its exchange space (the butterflies above consumed every gathered value, and a warp
  // issues in order): from the cluster's point of view it may now be overwritten.  Nothing is published here, so
  // the arrive is relaxed (the release form drains the memory pipe: ERRBAR, 5 % of the stall samples).
  cluster_arrive_relaxed();
  // v[e] = local output k1 = t + e*TPT; inner four-step twiddle w_N^(k1 * rank) from the table ctw[rank-1][k1]:
  // one anchor per 8 registers and a running product with the step w_N^(TPT * rank) in between (<= 8 ulp), so a
  // thread reads 1 + E/8 table entries instead of E.  Rank 0 multiplies by one: CTA-uniform skip.
  if (rank != 0) {
    const C* wp = ctw + (size_t)(rank - 1) * K::N;
    constexpr int CH = (K::E < 8) ? K::E : 8;
    const C stepw = __ldg(wp + K::TPT);
    static_for<0, K::E / CH>([&](auto qc) {
      constexpr int q = qc;
      C w = __ldg(wp + t + q * CH * K::TPT);
      static_for<0, CH>([&](auto rc) {
        constexpr int e = q * CH + rc;
        v[e] = cmul(v[e], w);
        if constexpr (rc + 1 < CH) w = cmul(w, stepw);
      });
    });
  }

  // ---- exchange: register e goes to CTA e / EP, slot [rank][(e % EP)*TPT + t][l] ----------------------
  cluster_wait();   // every CTA of the cluster has finished reading its own exchange space
  {
    const uint32_t base = smem_u32(sm) + (uint32_t)(((int)rank * KL + t) * K::TL + l) * (uint32_t)sizeof(C);
    static_for<0, CS>([&](auto dc) {
      constexpr int d = dc;
      const uint32_t ra = map_to_rank(base, (unsigned)d);
      static_for<0, EP>([&](auto jc) {
        constexpr int j = jc;
        st_cluster(ra + (uint32_t)(j * K::TPT * K::TL) * (uint32_t)sizeof(C), v[d * EP + j]);
      });
    });
  }
  cluster_arrive();
  cluster_wait();   // all remote stores have landed

  // ---- phase 2: radix-CS butterflies over n2 for k1 = rank*KL + t + j*TPT, store rows k1 + N1*k2 ------
  {
    const bool valid = line < g.nl;
    const C* base = sm + t * K::TL + l;
    C* op = out + (long long)b * g.obs + (long long)o * g.oos + line + ((long long)rank * KL + t) * g.ons;
    const unsigned jstep_b = (unsigned)((long long)K::TPT * g.ons * (long long)sizeof(C));
    const unsigned long long kstep_b = (unsigned long long)((long long)K::N * g.ons * (long long)sizeof(C));
    auto tail = [&](auto cj) {
      constexpr bool CJ = decltype(cj)::value;
      static_for<0, EP>([&](auto jc) {
        constexpr int j = jc;
        C a[CS];
        static_for<0, CS>([&](auto nc) { constexpr int n2 = nc; a[n2] = base[(n2 * KL + j * K::TPT) * K::TL]; });
        dft<CS>(a);
        if constexpr (TW4) {
          const unsigned m = g.tw_from_o ? (unsigned)o : (unsigned)line / (unsigned)g.tw_div;
          const unsigned lomask = (1u << g.tw_lo_bits) - 1u;
          auto root = [&](unsigned x) { return cmul(__ldg(tw_lo + (x & lomask)), __ldg(tw_hi + (x >> g.tw_lo_bits))); };
          const unsigned k0 = rank * KL + t + j * K::TPT;
          C w = root(k0 * m);
          const C stepw = root((unsigned)K::N * m);
          static_for<0, CS>([&](auto kc) {
            constexpr int k2 = kc;
            a[k2] = cmul(a[k2], w);
            if constexpr (k2 + 1 < CS) w = cmul(w, stepw);
          });
        }
        if (scale != (T)1) {
          const T sy = CJ ? -scale : scale;
          static_for<0, CS>([&](auto kc) { constexpr int k2 = kc; a[k2].x *= scale; a[k2].y *= sy; });
        } else if constexpr (CJ) {
          static_for<0, CS>([&](auto kc) { constexpr int k2 = kc; a[k2].y = -a[k2].y; });
        }
        if (valid) {
          char* p = reinterpret_cast<char*>(op) + (unsigned long long)(unsigned)j * jstep_b;
          static_for<0, CS>([&](auto kc) {
            constexpr int k2 = kc;
            st_stream(reinterpret_cast<C*>(p + (unsigned long long)k2 * kstep_b), a[k2]);
          });
        }
      });
    };
    if (g.swap_out) tail(std::true_type{}); else tail(std::false_type{});
  }
}

}  // namespace b200fft

namespace b200fft {

// ---- rows + the first radix-CS stage of a strided axis, in one pass -------------------------------------------------
// A 2D transform [H][W] whose column axis is too long for one pass (H = CS*M) normally costs rows + two column passes.
// Here a cluster of CS CTAs takes the CS rows {n2 + M*n1} (n1 = CTA rank): every CTA transforms its row along W with
// the ordinary row stages, then the cluster does the radix-CS butterflies ACROSS the rows (the first four-step stage of
// the column axis, n = n1*M + n2) -- rank c collects columns [c*W/CS, (c+1)*W/CS) of all CS rows through distributed
// shared memory (register e goes to CTA e/(E/CS): compile-time), butterflies over n1, multiplies by w_H^(k1*n2) (one
// constant per output row) and stores W/CS contiguous elements of each of the CS rows k1*M + n2.  What is left of the
// column axis is ONE M-point pass over consecutive rows (stride W, index-reversed store k1 + CS*k2).
template <class K_, int CS_>
struct ClusterRowCfg {
  using K = K_;
  static constexpr int CS = CS_;
  static constexpr int EP = K::E / CS_;          // phase-2 butterflies per thread
  static constexpr int CH = K::N / CS_;          // columns owned by one CTA in phase 2
  static_assert(K::TL == 1 && K::S >= 2, "one contiguous row per CTA");
  static_assert(K::E % CS_ == 0 && (CS_ == 2 || CS_ == 4 || CS_ == 8 || CS_ == 16), "cluster size = register radix across the rows");
  static constexpr int SM_ELEMS = K::ROW_ELEMS > K::N ? K::ROW_ELEMS : K::N;
  static constexpr size_t SMEM = (size_t)SM_ELEMS * K::ESZ;
};

// Geom: row (b, n1, n2) of the input at b*ibs + n1*ios + n2*ils (W contiguous elements, ins == 1); same for the output
// with obs / oos / ols.  g.nl = M (clusters per batch entry), g.nb = batch.  ctw[(k1-1)*M + n2] = w_H^(k1*n2).
template <class CC>
__global__ void __launch_bounds__(CC::K::THREADS, CC::K::MINB)
fft_cluster_rows_kernel(const Geom g, const cpx_t<typename CC::K::real>* __restrict__ in, cpx_t<typename CC::K::real>* __restrict__ out,
                        const cpx_t<typename CC::K::real>* __restrict__ tws, const cpx_t<typename CC::K::real>* __restrict__,
                        const cpx_t<typename CC::K::real>* __restrict__, typename CC::K::real scale,
                        const cpx_t<typename CC::K::real>* __restrict__ ctw) {
  using K = typename CC::K;
  using T = typename K::real;
  using C = cpx_t<T>;
  constexpr int CS = CC::CS, EP = CC::EP, CH = CC::CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);

  const int t = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  const unsigned tile = cluster_id_x();
  const int n2 = (int)(tile % (unsigned)g.nl), b = (int)(tile / (unsigned)g.nl);

  // ---- phase 1: this CTA's row along W ---------------------------------------------------------------
  C v[K::E];
  {
    const C* ip = in + (long long)b * g.ibs + (long long)rank * g.ios + (long long)n2 * g.ils + t;
    auto head = [&](auto cj) {
      constexpr bool CJ = decltype(cj)::value;
      static_for<0, K::E>([&](auto ec) {
        constexpr int e = ec;
        C x = ld_stream(ip + e * K::TPT);
        if constexpr (CJ) x.y = -x.y;
        v[e] = x;
      });
      run_stage<K, 0>(v, t, tws);
      scatter<K, 0, false>(v, sm, 0, t);
    };
    if (g.swap_in) head(std::true_type{}); else head(std::false_type{});
  }
  static_for<1, K::S - 1>([&](auto sc) {
    constexpr int s = sc;
    __syncthreads();
    gather<K, false>(v, sm, 0, t);
    run_stage<K, s>(v, t, tws);
    __syncthreads();
    scatter<K, s, false>(v, sm, 0, t);
  });
  __syncthreads();
  gather<K, false>(v, sm, 0, t);
  run_stage<K, K::S - 1>(v, t, tws);      // v[e] = X[kx = t + e*TPT] of row n1 = rank
  cluster_arrive_relaxed();               // this CTA no longer reads its exchange space

  // ---- exchange: columns [c*CH, (c+1)*CH) = registers [c*EP, (c+1)*EP) go to CTA c, slot [rank][(e % EP)*TPT + t] ----
  cluster_wait();
  {
    const uint32_t base = smem_u32(sm) + (uint32_t)((int)rank * CH + t) * (uint32_t)sizeof(C);
    static_for<0, CS>([&](auto dc) {
      constexpr int d = dc;
      const uint32_t ra = map_to_rank(base, (unsigned)d);
      static_for<0, EP>([&](auto jc) {
        constexpr int j = jc;
        st_cluster(ra + (uint32_t)(j * K::TPT) * (uint32_t)sizeof(C), v[d * EP + j]);
      });
    });
  }
  cluster_arrive();
  cluster_wait();

  // ---- phase 2: radix-CS over n1 for columns rank*CH + t + j*TPT, twiddle w_H^(k1*n2), rows k1*M + n2 -----------
  {
    C w[CS];
    w[0] = C{(T)1, (T)0};
    static_for<1, CS>([&](auto kc) { constexpr int k1 = kc; w[k1] = __ldg(ctw + (size_t)(k1 - 1) * g.nl + n2); });
    C* op = out + (long long)b * g.obs + (long long)n2 * g.ols + (long long)rank * CH + t;
    const T sy = g.swap_out ? -scale : scale;
    const bool post = (scale != (T)1) || g.swap_out;
    static_for<0, EP>([&](auto jc) {
      constexpr int j = jc;
      C a[CS];
      static_for<0, CS>([&](auto nc) { constexpr int n1 = nc; a[n1] = sm[n1 * CH + j * K::TPT + t]; });
      dft<CS>(a);
      static_for<1, CS>([&](auto kc) { constexpr int k1 = kc; a[k1] = cmul(a[k1], w[k1]); });
      if (post) static_for<0, CS>([&](auto kc) { constexpr int k1 = kc; a[k1].x *= scale; a[k1].y *= sy; });
      static_for<0, CS>([&](auto kc) { constexpr int k1 = kc; op[(long long)k1 * g.oos + j * K::TPT] = a[k1]; });
    });
  }
}

}  // namespace b200fft
