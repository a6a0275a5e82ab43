// Persistent TMA-fed two-phase kernels with an L2-resident intermediate (band_kernel.cuh).
#include "band_kernel.cuh"
#include "kernel_inst.cuh"
namespace b200fft {

template <class P>
static BandEntry make_band() {
  BandEntry e{};
  e.is_double = sizeof(typename P::real) == 8;
  e.mode = P::MODE;
  e.outer = P::OUTER ? 1 : 0;
  e.N1 = P::N1; e.N2 = P::N2; e.TLA = P::TLA; e.TLB = P::TLB;
  e.a = describe_cfg<typename P::KA, true, true, true>();
  e.b = P::B_ROWS ? describe_cfg<typename P::KB, false, true, false>() : describe_cfg<typename P::KB, true, true, false>();
  e.threads = P::THREADS;
  e.smem = P::SMEM;
  e.func = reinterpret_cast<const void*>(&fft_band_kernel<P>);
  return e;
}

#ifndef B200FFT_BAND_G
#define B200FFT_BAND_G 6
#endif
#ifndef B200FFT_BAND_NSTG
#define B200FFT_BAND_NSTG 2
#endif

void register_band(void (*add)(const BandEntry&)) {
  constexpr int G = B200FFT_BAND_G, NS = B200FFT_BAND_NSTG;
  using F64 = Cfg<float, 64, 16, 32, 1, 16, 4>;      // 128 threads, 32 lines (256 B runs), 16 KB tiles
  using F128 = Cfg<float, 128, 16, 16, 1, 16, 8>;    // 128 threads, 16 lines (128 B runs), 16 KB tiles
  // strided axis of N1*N2 points: cfg3's column axis 8192 = 64 x 128; 4096 and 16384 for the neighbouring sizes
  add(make_band<BandCfg<F64, F128, MODE_STRIDED, false, G, NS>>());
}
}  // namespace b200fft
