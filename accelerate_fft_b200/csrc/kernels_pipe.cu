// Persistent software-pipelined column kernels (pipe_kernel.cuh), plain (CS = 1) and cluster (CS > 1).
//                      PipeCfg<Cfg<T, N1, E, TL, minb, R0, R1, R2, R3>, CS>   N = N1 * CS
#include "kernel_inst.cuh"
#include "pipe_kernel.cuh"
namespace b200fft {

template <class P, bool TW4>
KernelEntry make_pipe_entry() {
  using K = typename P::K;
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = P::N; e.N1 = K::N; e.CS = P::CS; e.E = K::E; e.TL = K::TL;
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.flavor = P::ROWS ? FL_PIPEROW : FL_PIPE;
  e.tw4 = TW4;
  e.threads = P::THREADS;
  e.smem = P::SMEM;
  e.G = P::G; e.NS = P::NQ;
  e.minb = 1;
  e.func = reinterpret_cast<const void*>(&fft_pipe_cols_kernel<P, TW4>);
  return e;
}

#define REG_PIPE(CS, ...) add(make_pipe_entry<PipeCfg<Cfg<__VA_ARGS__>, CS>, false>())
#define REG_PIPE_ROWS(...) add(make_pipe_entry<PipeCfg<Cfg<__VA_ARGS__>, 1, 4, true>, false>())
#define REG_PIPE_TW(CS, ...) add(make_pipe_entry<PipeCfg<Cfg<__VA_ARGS__>, CS>, true>())   // + four-step twiddle at the store

void register_pipe(void (*add)(const KernelEntry&)) {
  // c64: 1024-point CTA share, 8 columns (64 B runs), 2 x 256 threads x 128 registers, 64 KB landing + 2 x 66 KB exchange
  REG_PIPE(1, float, 1024, 32, 8, 1, 32, 32);
  // (landing buffer in 2 / 8 parts instead of 4: 88.1 % / 74.5 % against 90.2 % on [64][1024][1024])
  REG_PIPE_TW(1, float, 1024, 32, 8, 1, 32, 32);
  REG_PIPE(2, float, 1024, 32, 8, 1, 32, 32);
#ifdef B200FFT_EXPERIMENTAL   // opt-in, measured slower than the default plans (they ride on the cluster column kernels' planner path)
  REG_PIPE(4, float, 1024, 32, 8, 1, 32, 32);
  REG_PIPE(8, float, 1024, 32, 8, 1, 32, 32);            // cfg3's column axis
  REG_PIPE(16, float, 1024, 32, 8, 1, 32, 32);
  // contiguous rows: one 64 KB line per tile
  REG_PIPE_ROWS(float, 8192, 32, 1, 1, 32, 16, 16);
  REG_PIPE_ROWS(double, 4096, 16, 1, 1, 16, 16, 16);
#endif
  // c64 N=512: 16 columns (128 B runs)
  REG_PIPE(1, float, 512, 32, 16, 1, 32, 16);
  REG_PIPE_TW(1, float, 512, 32, 16, 1, 32, 16);
  // c128 N=1024 in one CTA: 4 columns (64 B runs), three stages -- 78 % of HBM peak against 54 % for the 128 KB lock-step tile
  // (registered before the 512 x 2 cluster so that it is the default for N=1024)
  REG_PIPE(1, double, 1024, 16, 4, 1, 16, 16, 4);
  REG_PIPE_TW(1, double, 1024, 16, 4, 1, 16, 16, 4);
  // c128: 512-point CTA share, 8 columns (128 B runs)
  REG_PIPE(1, double, 512, 16, 8, 1, 16, 16, 2);
  REG_PIPE_TW(1, double, 512, 16, 8, 1, 16, 16, 2);
  REG_PIPE(2, double, 512, 16, 8, 1, 16, 16, 2);
#ifdef B200FFT_EXPERIMENTAL
  REG_PIPE(4, double, 512, 16, 8, 1, 16, 16, 2);
  REG_PIPE(8, double, 512, 16, 8, 1, 16, 16, 2);
  REG_PIPE(16, double, 512, 16, 8, 1, 16, 16, 2);
#endif
}
}  // namespace b200fft
