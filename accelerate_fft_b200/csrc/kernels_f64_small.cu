// c128 row kernels, N = 2 .. 1024
#include "kernel_inst.cuh"
namespace b200fft {
void register_f64_small(void (*add)(const KernelEntry&)) {
  REG_ROW(double, 2, 2, 128, 0, 2);
  REG_ROW(double, 4, 4, 128, 0, 4);
  REG_ROW(double, 8, 8, 128, 0, 8);
  REG_ROW(double, 16, 16, 64, 0, 16);                  // v0: one line per thread, rows staged through shared memory (77 % -> 107 %)
  REG_ROW(double, 16, 8, 64, 0, 8, 2);                 // v1
  REG_ROW(double, 32, 8, 32, 0, 8, 4);
  REG_ROW(double, 64, 8, 16, 0, 8, 8);
  REG_ROW(double, 128, 8, 8, 0, 8, 8, 2);
  REG_ROW(double, 256, 8, 4, 0, 8, 8, 4);
  REG_ROW(double, 512, 8, 2, 0, 8, 8, 8);
  REG_ROW(double, 1024, 16, 2, 0, 16, 8, 8);          // v0  102 %
  REG_ROW(double, 1024, 8, 2, 0, 8, 8, 8, 2);         // v1   85 %
}
}  // namespace b200fft
