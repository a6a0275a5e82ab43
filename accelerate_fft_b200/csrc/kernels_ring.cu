// Persistent TMA-fed row kernels (ring_kernel.cuh) for the tile sizes where occupancy alone cannot
// hide HBM latency.          RingCfg<Cfg<T, N, E, TL, minb, R0, R1, R2, R3>, G groups, NS stages>
#include "kernel_inst.cuh"
#include "ring_kernel.cuh"
namespace b200fft {

template <class R>
KernelEntry make_ring_entry() {
  using K = typename R::Base;
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = K::N; e.E = K::E; e.TL = K::TL;
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.flavor = FL_RING;
  e.threads = R::THREADS;
  e.smem = R::SMEM;
  e.G = R::G; e.NS = R::NS;
  e.minb = 1;
  e.func = reinterpret_cast<const void*>(&fft_ring_rows_kernel<R>);
  return e;
}

#define REG_RING(G, NS, ...) add(make_ring_entry<RingCfg<Cfg<__VA_ARGS__>, G, NS>>())

void register_ring(void (*add)(const KernelEntry&)) {
  // cfg2 (c128 N=4096, 64 KB lines).  v0: two single-stage 256-thread CTAs per SM -- the next line lands while the last register
  // stage and the stores of the current one run: 1424 us per direction; v1 (B200FFT_VARIANTS=g4096d=1): one CTA of two groups
  // alternating over three stages: 1442-1450 us (round 1's default); plain kernel: 1.58 ms
  REG_RING(1, 1, double, 4096, 16, 1, 2, 16, 16, 16);
  REG_RING(2, 3, double, 4096, 16, 1, 1, 16, 16, 16);
  // small batches of short rows (cfg1: 4096 rows of c64 N=1024 = 14 tiles per SM): the whole share of an SM is
  // requested from HBM at kernel start instead of wave by wave
  REG_RING(4, 12, float, 1024, 16, 2, 1, 16, 16, 4);     // v0: 4 groups x 128 thr, 12 x 17 KB stages
  REG_RING(8, 24, float, 1024, 16, 1, 1, 16, 16, 4);     // v1: 8 groups x 64 thr, 24 x 8.5 KB stages
  // 128 KB lines (c64 N=16384, c128 N=8192): one CTA per SM and no room for a second buffer -- a "ring" of ONE stage still lets the
  // TMA engine fetch the next line while the last register stage and the stores of the current one run
  REG_RING(1, 1, float, 16384, 32, 1, 1, 32, 32, 16);
  REG_RING(1, 1, float, 8192, 32, 1, 2, 32, 16, 16);      // 64 KB lines, two single-stage CTAs per SM (cfg3's rows)
  REG_RING(1, 1, double, 8192, 16, 1, 1, 16, 16, 16, 2);
  // measured on B200 (profiles/r01_ring_vs_plain.txt): the ring wins only where FP64 + 68 KB tiles leave the plain
  // kernel latency-bound (c128 N=4096: 81.7 % -> 88.4 % of measured HBM peak).  It loses for c128 N=2048
  // (93.0 % vs 96.3 %), c64 N=8192 (73.5 % vs 80.4 %) and c64 N=4096 (81.3 % vs 91.6 %), where the extra
  // shared-memory traffic of the staged tile costs more than the hidden latency gains -- not instantiated.
}
}  // namespace b200fft
