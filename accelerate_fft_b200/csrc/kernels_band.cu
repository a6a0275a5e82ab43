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

void register_band(void (*add)(const BandEntry&)) {
  using F64 = Cfg<float, 64, 16, 32, 1, 16, 4>;      // 128 threads, 32 lines (256 B runs), 16 KB tiles
  using F128 = Cfg<float, 128, 16, 16, 1, 16, 8>;    // 128 threads, 16 lines (128 B runs), 16 KB tiles
  // strided axis of N1*N2 points (6 groups, 7 landing stages; 4 groups / 9 stages measured equal within 3 %):
  add(make_band<BandCfg<F64, F128, MODE_STRIDED, false, 6, 7>>());     // 8192 = 64 x 128: cfg3's column axis
  // (4096 = 64 x 64 measured equal to the two unfused passes -- 141 vs 138 us on 4096^2 -- and is left to them)
  add(make_band<BandCfg<F128, F128, MODE_STRIDED, false, 6, 7>>());    // 16384 = 128 x 128 columns
  // large contiguous 1D transforms, N = Nout x 16384 in TWO HBM round trips (cfg4: 2^28 = 2^14 x 2^14):
  //   pass 1 = the Nout-point strided axis + the outer four-step twiddle, pass 2 = 16384-point rows, transposed store
  add(make_band<BandCfg<F64, F128, MODE_STRIDED, true, 6, 7>>());
  add(make_band<BandCfg<F128, F128, MODE_STRIDED, true, 6, 7>>());
  add(make_band<BandCfg<F128, F128, MODE_ROWS, false, 6, 6>>());   // (row-layout exchange buffers: 6 landing stages fit)
}
}  // namespace b200fft
