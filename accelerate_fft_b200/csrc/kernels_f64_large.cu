// c128 row kernels, N = 2048 .. 8192
#include "kernel_inst.cuh"
namespace b200fft {
void register_f64_large(void (*add)(const KernelEntry&)) {
  REG_ROW(double, 2048, 8, 1, 8, 8, 8, 4);
  REG_ROW(double, 4096, 8, 1, 8, 8, 8, 8);
  REG_ROW(double, 8192, 16, 1, 16, 16, 16, 2);
}
}  // namespace b200fft
