// c64 row kernels, N = 2048 .. 16384
#include "kernel_inst.cuh"
namespace b200fft {
void register_f32_large(void (*add)(const KernelEntry&)) {
  REG_ROW(float, 2048, 16, 1, 0, 16, 16, 8);
  REG_ROW(float, 4096, 16, 1, 0, 16, 16, 16);
  // (radix orders 32x32x8, 16x16x32, 16x32x16, 8x32x32 measured: 83.0 / 84.0 / 82.6 / 69.2 % against 85.4 %)
  REG_ROW(float, 8192, 32, 1, 0, 32, 16, 16);         // v0: 256 thr x 128 regs -> 2 CTA/SM
  REG_ROW(float, 8192, 16, 1, 2, 16, 16, 8, 4);       // v1: 512 thr x 64 regs -> 2 CTA/SM (66 % vs 80 %)
  REG_ROW(float, 16384, 32, 1, 0, 32, 32, 16);        // v0
  REG_ROW(float, 16384, 16, 1, 0, 16, 16, 16, 4);     // v1: 1024 thr x 64 regs
}
}  // namespace b200fft
