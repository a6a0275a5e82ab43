// c128 row kernels, N = 2048 .. 8192
#include "kernel_inst.cuh"
namespace b200fft {
void register_f64_large(void (*add)(const KernelEntry&)) {
  // measured on B200 (profiles/variants_r01.txt): E=16 beats E=8 at every c128 size >= 1024
  REG_ROW(double, 2048, 16, 1, 0, 16, 16, 8);         // v0: 128 thr x 128 regs            96.8 %
  REG_ROW(double, 2048, 8, 1, 0, 8, 8, 8, 4);         // v1                                 85.6 %
  REG_ROW(double, 4096, 16, 1, 2, 16, 16, 16);        // v0: 256 thr x 128 regs, 2 CTA/SM  82.2 %
  REG_ROW(double, 4096, 8, 1, 2, 8, 8, 8, 8);         // v1: 512 thr x 64 regs, 2 CTA/SM   76.2 %
  REG_ROW(double, 4096, 8, 1, 1, 8, 8, 8, 8);         // v2: 512 thr x 114 regs, 1 CTA/SM  56.3 %
  // (tried: minb=3 -> 80 regs, 776 B spills: 43 % -- removed)
  REG_ROW(double, 8192, 16, 1, 0, 16, 16, 16, 2);
}
}  // namespace b200fft
