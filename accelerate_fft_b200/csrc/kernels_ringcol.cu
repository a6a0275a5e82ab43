// Persistent single-buffer TMA-fed column kernels (ringcol_kernel.cuh) for the column tiles that fill an SM's shared memory.
#include "kernel_inst.cuh"
#include "ringcol_kernel.cuh"
namespace b200fft {

template <class R>
static KernelEntry make_ringcol_entry() {
  using K = typename R::K;
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = K::N; e.E = K::E; e.TL = K::TL;
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.flavor = FL_RINGCOL;
  e.tw4 = R::TW4;
  e.threads = R::THREADS;
  e.smem = R::SMEM;
  e.G = 1; e.NS = 1;
  e.N1 = R::BOX_ROWS;          // rows per TMA box
  e.minb = 1;
  e.func = reinterpret_cast<const void*>(&fft_ringcol_kernel<R>);
  return e;
}

template <class R>
static KernelEntry make_ringtrans_entry() {
  using K = typename R::K;
  KernelEntry e{};
  e.is_double = sizeof(typename K::real) == 8;
  e.N = K::N; e.E = K::E; e.TL = K::TL;
  e.S = K::S;
  for (int i = 0; i < 4; i++) e.rad[i] = K::rad[i];
  e.tw_len = K::TW_LEN;
  e.flavor = FL_RINGTRANS;
  e.threads = R::THREADS;
  e.smem = R::SMEM;
  e.G = 1; e.NS = 1;
  e.minb = 1;
  e.func = reinterpret_cast<const void*>(&fft_ringtrans_kernel<R>);
  return e;
}

void register_ringcol(void (*add)(const KernelEntry&)) {
  // the last pass of a big contiguous 1D transform (cfg4: 512-point rows, transposed store), 32 lines per tile: 256 B runs
  add(make_ringtrans_entry<RingTransCfg<Cfg<float, 512, 32, 32, 1, 32, 16>>>());
  add(make_ringcol_entry<RingColCfg<Cfg<float, 1024, 32, 16, 1, 32, 32>>>());       // 512 thr x 128 regs, 128 B runs, 128 KB: cfg5's z axis
  // the first pass of a big contiguous 1D transform (cfg4: 512-point columns at a 4 MB row stride + the four-step twiddle), 32 columns
  // wide: 256 B runs where the lock-step kernel's 64 KB tile has 128 B runs
  add(make_ringcol_entry<RingColCfg<Cfg<float, 512, 32, 32, 1, 32, 16>, true>>());
  add(make_ringcol_entry<RingColCfg<Cfg<double, 1024, 16, 8, 1, 16, 8, 8>>>());     // c128: 8 columns (128 B runs), 128 KB; the radices of the lock-step kernel
}
}  // namespace b200fft
