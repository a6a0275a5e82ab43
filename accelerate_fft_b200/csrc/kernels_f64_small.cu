// c128 row kernels, N = 2 .. 1024
#include "kernel_inst.cuh"
namespace b200fft {
void register_f64_small(void (*add)(const KernelEntry&)) {
  REG_ROW(double, 2, 2, 128, 2);
  REG_ROW(double, 4, 4, 128, 4);
  REG_ROW(double, 8, 8, 128, 8);
  REG_ROW(double, 16, 8, 64, 8, 2);
  REG_ROW(double, 32, 8, 32, 8, 4);
  REG_ROW(double, 64, 8, 16, 8, 8);
  REG_ROW(double, 128, 8, 8, 8, 8, 2);
  REG_ROW(double, 256, 8, 4, 8, 8, 4);
  REG_ROW(double, 512, 8, 2, 8, 8, 8);
  REG_ROW(double, 1024, 8, 2, 8, 8, 8, 2);
}
}  // namespace b200fft
