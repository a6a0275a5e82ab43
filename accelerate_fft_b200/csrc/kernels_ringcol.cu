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
  e.threads = R::THREADS;
  e.smem = R::SMEM;
  e.G = 1; e.NS = 1;
  e.N1 = R::BOX_ROWS;          // rows per TMA box
  e.minb = 1;
  e.func = reinterpret_cast<const void*>(&fft_ringcol_kernel<R>);
  return e;
}

void register_ringcol(void (*add)(const KernelEntry&)) {
  add(make_ringcol_entry<RingColCfg<Cfg<float, 1024, 32, 16, 1, 32, 32>>>());       // 512 thr x 128 regs, 128 B runs, 128 KB: cfg5's z axis
  add(make_ringcol_entry<RingColCfg<Cfg<double, 1024, 16, 8, 1, 16, 8, 8>>>());     // c128: 8 columns (128 B runs), 128 KB; the radices of the lock-step kernel
}
}  // namespace b200fft
